// Scalar minimizer arithmetic for code that handles one stored key per thread (db_sort).  The read path computes the
// same quantity warp-cooperatively in k_scan; this per-key form is host-compilable so that a CPU test can hold it
// against the oracle (tests/test_cabi_cpu.py::test_scalar_minimizer_matches_oracle) — it is not a CPU path of the
// product: nothing in libkuq.so calls it on the host.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define KUQ_HD __host__ __device__ __forceinline__
#else
#define KUQ_HD inline
#endif

namespace kuq {

KUQ_HD uint64_t revcomp_n(uint64_t x, uint32_t n) {                            // krakendb.cpp:218-225
  x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
  x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
  x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
  x = (x >> 32) | (x << 32);
  return (~x) >> (64 - 2 * n);
}

// KrakenDB::bin_key(kmer, nt), krakendb.cpp:200-215, KRAKIX2 flavour, on the key AS STORED (db_sort does not
// canonicalise the k-mer, db_sort.cpp:101)
KUQ_HD uint32_t bin_key_of(uint64_t kmer, uint32_t k, uint32_t nt) {
  const uint64_t mask = (1ull << (2 * nt)) - 1;
  const uint64_t xor_mask = 0xe37e28c4271b5a2dull & mask;                     // INDEX2_XOR_MASK, :45
  uint64_t best = ~0ull;
  for (uint32_t i = 0; i + nt <= k; i++) {
    const uint64_t m = kmer & mask, rc = revcomp_n(m, nt);
    const uint64_t t = xor_mask ^ (m < rc ? m : rc);
    best = t < best ? t : best;
    kmer >>= 2;
  }
  return (uint32_t)best;
}

}  // namespace kuq
