#!/usr/bin/env bash
# Build the UNMODIFIED reference (fbreitwieser/krakenuniq) hot-path programs from the sources where they
# lie under $KUQ_REFERENCE_SRC (default /root/reference/src) into oracle/_ref/ (git-ignored, shipped by gpurun).
# TEST INFRASTRUCTURE ONLY: nothing under krakenuniq_b200/ may link or execute these files.
#
# No reference source is copied or edited.  Two command-line work-arounds replace the edits SURVEY §8(c) made
# to a private copy:
#   * `-include cstdint`        : uid_mapping.hpp:42 uses uint32_t without <cstdint> (g++ 13).
#   * `-DBXZSTR_CONFIG_HPP ...` : pre-defines the include guard of third_party/bxzstr/include/config.hpp so its
#                                 `#define BXZSTR_BZ2_SUPPORT 1` is skipped (no bzlib.h in this image); zlib stays on.
# Everything else is the flags of src/Makefile:14 (-O2 -std=c++11 -fopenmp -DNDEBUG).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${KUQ_REFERENCE_SRC:-/root/reference/src}"
OUT="$HERE/_ref"
if [ ! -d "$SRC" ]; then
  echo "build_ref.sh: reference sources not found at $SRC (expected on the GPU box: prebuilt oracle/_ref travels)" >&2
  exit 3
fi
mkdir -p "$OUT/obj"
CXX="${KUQ_CXX:-/usr/bin/g++}"   # not $CXX: this image exports CXX=/opt/gcc/bin/g++, which has no libgomp.spec
CXXFLAGS=(-Wall -Wextra -Wfatal-errors -pipe -O2 -std=c++11 -fopenmp -DNDEBUG -w
          -include cstdint -I"$SRC" -I"$SRC/gzstream"
          -DBXZSTR_CONFIG_HPP -DBXZSTR_Z_SUPPORT=1 -DBXZSTR_BZ2_SUPPORT=0 -DBXZSTR_LZMA_SUPPORT=0 -DBXZSTR_ZSTD_SUPPORT=0)
cc() { # cc <src> <obj> [extra flags]
  local s="$1" o="$2"; shift 2
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ]; then "$CXX" "${CXXFLAGS[@]}" "$@" -c "$s" -o "$o"; fi
}
for f in krakendb quickfile krakenutil seqreader uid_mapping hyperloglogplus; do
  cc "$SRC/$f.cpp" "$OUT/obj/$f.o" &
done
cc "$SRC/gzstream/gzstream.C" "$OUT/obj/gzstream.o" &
# position-independent copies for the KAT shim library
for f in krakendb quickfile krakenutil hyperloglogplus; do
  cc "$SRC/$f.cpp" "$OUT/obj/$f.pic.o" -fPIC &
done
wait
O="$OUT/obj"
"$CXX" "${CXXFLAGS[@]}" -o "$OUT/classify" "$SRC/classify.cpp" $O/krakendb.o $O/quickfile.o $O/krakenutil.o $O/seqreader.o $O/uid_mapping.o $O/gzstream.o $O/hyperloglogplus.o -lz &
"$CXX" "${CXXFLAGS[@]}" -DEXACT_COUNTING -o "$OUT/classifyExact" "$SRC/classify.cpp" $O/krakendb.o $O/quickfile.o $O/krakenutil.o $O/seqreader.o $O/uid_mapping.o $O/gzstream.o $O/hyperloglogplus.o -lz &
"$CXX" "${CXXFLAGS[@]}" -o "$OUT/db_sort" "$SRC/db_sort.cpp" $O/krakendb.o $O/quickfile.o -lz &
"$CXX" "${CXXFLAGS[@]}" -o "$OUT/set_lcas" "$SRC/set_lcas.cpp" $O/krakendb.o $O/quickfile.o $O/krakenutil.o $O/seqreader.o $O/uid_mapping.o -lz &
# KAT shim: OUR thin extern "C" wrapper (oracle/ref_shim.cpp) around the reference's own classes
"$CXX" "${CXXFLAGS[@]}" -fPIC -shared -o "$OUT/libkuref.so" "$HERE/ref_shim.cpp" $O/krakendb.pic.o $O/quickfile.pic.o $O/krakenutil.pic.o $O/hyperloglogplus.pic.o -lz &
wait
echo "oracle/_ref built: $(ls "$OUT" | tr '\n' ' ')"
