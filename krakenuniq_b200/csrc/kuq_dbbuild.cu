// db_sort on the GPU — SURVEY.md §8 row f4 (db_sort.cpp:41-116, make_index krakendb.cpp:118-148).
// ROUND 1 STATUS: compiled for sm_100a, NOT yet run on hardware (written after the round's GPU budget was spent);
// its tests are gated (tests/test_zz_unvalidated_gpu.py).  The CPU statement of the same tool, pinned byte for byte
// against the reference executable, is oracle/kuq_oracle.c:kuqo_db_sort.
//
// The reference bins the records by bin_key(key) and qsorts every bin by key.  Keys are distinct, so the result is
// THE order by (bin, key): one stable LSD radix sort by key followed by one by bin.  The two sorts are CUB's
// DeviceRadixSort (library code; database build is not the hot path north_star names), the kernels around them —
// minimizer of every stored key, bin histogram, record gather — are ours.  Records never move during the sorts:
// only (key, index) and (bin, index) pairs do, and one gather at the end writes the 12-byte records.
#include "kuq_kernels.cuh"
#include "kuq_minimizer.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace kuq {
namespace {

// one thread per record: key, its bin, its position; the bin histogram on the way (make_index :126-134)
__global__ void k_dbs_keys(const uint8_t *pairs, uint64_t n, uint32_t key_len, uint32_t k, uint32_t nt,
                           unsigned long long *keys, uint32_t *bins, unsigned long long *index,
                           unsigned long long *bin_counts) {
  const uint32_t pair_sz = key_len + 4;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint8_t *r = pairs + i * pair_sz;
    uint64_t key = 0;
    for (uint32_t b = 0; b < key_len; b++) key |= (uint64_t)r[b] << (8 * b);  // memcpy(&kmer, pair, key_len), :99-100
    const uint32_t bin = bin_key_of(key, k, nt);
    keys[i] = key;
    bins[i] = bin;
    index[i] = i;
    atomicAdd(bin_counts + bin + 1, 1ull);                                     // offsets[b + 1] after the scan
  }
}

__global__ void k_dbs_gather_bins(const uint32_t *bins, const unsigned long long *index, uint64_t n, uint32_t *out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = bins[index[i]];
}

// output record i = input record index[i]; -z zeroes the value (db_sort.cpp:105-106)
__global__ void k_dbs_emit(const uint8_t *in, const unsigned long long *index, uint64_t n, uint32_t key_len,
                           int zero_vals, uint8_t *out) {
  const uint32_t pair_sz = key_len + 4;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint8_t *src = in + index[i] * pair_sz;
    uint8_t *dst = out + i * pair_sz;
    for (uint32_t b = 0; b < key_len; b++) dst[b] = src[b];
    for (uint32_t b = 0; b < 4; b++) dst[key_len + b] = zero_vals ? (uint8_t)0 : src[key_len + b];
  }
}

int seterr(char *err, size_t cap, const char *fmt, ...) {
  if (err && cap) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, cap, fmt, ap);
    va_end(ap);
  }
  return -1;
}

struct Buffers {                    // frees on scope exit whatever was allocated
  void *p[16];
  int n = 0;
  template <typename T>
  cudaError_t get(T **out, uint64_t count) {
    cudaError_t e = cudaMalloc((void **)out, count * sizeof(T) + 16);
    if (e == cudaSuccess) p[n++] = *out;
    return e;
  }
  ~Buffers() { for (int i = 0; i < n; i++) cudaFree(p[i]); }
};

#define DB_CU(call)                                                                                      \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess) return seterr(err, err_cap, "db_sort: %s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

}  // namespace

// Host images in, host images out (kdb_out: header + key_ct * pair size bytes — what db_sort.cpp:70-72 writes;
// idx_out: 8 + 8 * (4^nt + 1) bytes).  Returns 0 or -1 (+ err).
int dbsort_device(const uint8_t *jdb, uint64_t jdb_bytes, uint32_t nt, int zero_vals, uint8_t *kdb_out, uint8_t *idx_out,
                  char *err, size_t err_cap) {
  if (!jdb || !kdb_out || !idx_out) return seterr(err, err_cap, "db_sort: NULL image");
  if (jdb_bytes < 56 || memcmp(jdb, "JFLISTDN", 8) != 0) return seterr(err, err_cap, "db_sort: not a Jellyfish-style database (krakendb.cpp:67)");
  if (nt < 1 || nt > 15) return seterr(err, err_cap, "db_sort: -n must be 1..15");
  uint64_t key_bits, val_len, key_ct;
  memcpy(&key_bits, jdb + 8, 8);
  memcpy(&val_len, jdb + 16, 8);
  memcpy(&key_ct, jdb + 48, 8);
  if (val_len != 4) return seterr(err, err_cap, "db_sort: can only handle 4 byte DB values (krakendb.cpp:73-74)");
  if (key_bits < 2 * nt || key_bits > 64) return seterr(err, err_cap, "db_sort: key_bits %llu out of range", (unsigned long long)key_bits);
  const uint32_t k = (uint32_t)(key_bits / 2), key_len = (uint32_t)(key_bits / 8 + !!(key_bits % 8));
  const uint64_t pair_sz = key_len + 4, header = 72 + 2 * (4 + 8 * key_bits);
  if (jdb_bytes < header + key_ct * pair_sz) return seterr(err, err_cap, "db_sort: image shorter than its header says");
  const uint64_t entries = 1ull << (2 * nt);
  const uint64_t n = key_ct;

  Buffers buf;
  uint8_t *d_in = nullptr, *d_out = nullptr;
  unsigned long long *d_keys = nullptr, *d_keys2 = nullptr, *d_idx = nullptr, *d_idx2 = nullptr, *d_off = nullptr;
  uint32_t *d_bins = nullptr, *d_bins2 = nullptr, *d_bins3 = nullptr;
  DB_CU(buf.get(&d_in, n * pair_sz));
  DB_CU(buf.get(&d_out, n * pair_sz));
  DB_CU(buf.get(&d_keys, n));
  DB_CU(buf.get(&d_keys2, n));
  DB_CU(buf.get(&d_idx, n));
  DB_CU(buf.get(&d_idx2, n));
  DB_CU(buf.get(&d_bins, n));
  DB_CU(buf.get(&d_bins2, n));
  DB_CU(buf.get(&d_bins3, n));
  DB_CU(buf.get(&d_off, entries + 1));
  cudaStream_t st = 0;
  DB_CU(cudaMemcpyAsync(d_in, jdb + header, n * pair_sz, cudaMemcpyHostToDevice, st));
  DB_CU(cudaMemsetAsync(d_off, 0, (entries + 1) * 8, st));
  const int grid = 148 * 8;
  if (n) k_dbs_keys<<<grid, 256, 0, st>>>(d_in, n, key_len, k, nt, d_keys, d_bins, d_idx, d_off);

  // offsets = running sum of the bin sizes (make_index :136-139)
  size_t tmp_bytes = 0, need = 0;
  DB_CU(cub::DeviceScan::InclusiveSum((void *)nullptr, need, d_off, d_off, (int64_t)(entries + 1), st));
  tmp_bytes = need;
  if (n) {
    DB_CU(cub::DeviceRadixSort::SortPairs((void *)nullptr, need, d_keys, d_keys2, d_idx, d_idx2, (int64_t)n, 0, (int)key_bits, st));
    if (need > tmp_bytes) tmp_bytes = need;
    DB_CU(cub::DeviceRadixSort::SortPairs((void *)nullptr, need, d_bins2, d_bins3, d_idx2, d_idx, (int64_t)n, 0, (int)(2 * nt), st));
    if (need > tmp_bytes) tmp_bytes = need;
  }
  uint8_t *d_tmp = nullptr;
  DB_CU(buf.get(&d_tmp, tmp_bytes));
  need = tmp_bytes;
  DB_CU(cub::DeviceScan::InclusiveSum(d_tmp, need, d_off, d_off, (int64_t)(entries + 1), st));
  if (n) {
    // by key (pair_cmp, db_sort.cpp:118-128) ...
    need = tmp_bytes;
    DB_CU(cub::DeviceRadixSort::SortPairs(d_tmp, need, d_keys, d_keys2, d_idx, d_idx2, (int64_t)n, 0, (int)key_bits, st));
    // ... then, stably, by bin (:96-107): the order inside a bin stays the key order
    k_dbs_gather_bins<<<grid, 256, 0, st>>>(d_bins, d_idx2, n, d_bins2);
    need = tmp_bytes;
    DB_CU(cub::DeviceRadixSort::SortPairs(d_tmp, need, d_bins2, d_bins3, d_idx2, d_idx, (int64_t)n, 0, (int)(2 * nt), st));
    k_dbs_emit<<<grid, 256, 0, st>>>(d_in, d_idx, n, key_len, zero_vals, d_out);
  }
  DB_CU(cudaGetLastError());
  memcpy(kdb_out, jdb, header);                                                // db_sort.cpp:56-58,71
  DB_CU(cudaMemcpyAsync(kdb_out + header, d_out, n * pair_sz, cudaMemcpyDeviceToHost, st));
  memcpy(idx_out, "KRAKIX2", 7);                                               // krakendb.cpp:144-147
  idx_out[7] = (uint8_t)nt;
  DB_CU(cudaMemcpyAsync(idx_out + 8, d_off, (entries + 1) * 8, cudaMemcpyDeviceToHost, st));
  DB_CU(cudaStreamSynchronize(st));
  return 0;
}

}  // namespace kuq
