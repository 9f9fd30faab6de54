// kuq_clades.cu — device side of kuq_clade_counts_tree: the sketches of ALL clades of the taxonomy in a few launches
// (the report of the reference rolls every taxon's ReadCounts into each ancestor, TaxReport, taxdb.hpp:928-982, with
// ReadCounts::operator+= = HLL merge, readcounts.hpp:76-88 / hyperloglogplus.cpp:596-621).
//
//  dense clades  (some member's sketch is dense → register-wise max of the members' registers, :604-621):
//      one CTA per clade folds its members' 4 KB register arrays; register histograms of all clades in one launch.
//  sparse clades (no dense member → union of the members' code sets, :600-603): the number of distinct codes of
//      every subtree at once.  Keys (taxon, code) of the sparse tier are sorted by (code, preorder(taxon)); the
//      taxa of a subtree are a preorder interval, so the occurrences of one code inside a subtree are a run of that
//      sorted sequence and   distinct(subtree u) = keys in u − adjacent equal-code pairs inside u,
//      and a pair lies inside u exactly when its taxa's lowest common ancestor does.  The kernel books every
//      adjacent pair at its LCA (per encoded rank, because the estimator wants the rank histogram of the union);
//      the host sums the books up the tree.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/cub.cuh>

#include "kuq_kernels.cuh"

namespace kuq {
namespace {

constexpr uint32_t NONE32 = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t enc_rank(uint32_t e) {          // getEncodedRank(enc, 25, 12), hyperloglogplus.cpp:152-161
  if (e & 1) return 13 + ((e >> 1) & 0x3F);
  const uint32_t r = e << 12;
  return (r ? (uint32_t)__clz(r) : 20u) + 1;
}

// out[c][0..4096) = max over the members of clade c; 256 threads x 16 bytes
__global__ void __launch_bounds__(256) k_clade_max_batch(const uint8_t *__restrict__ regs, const uint32_t *__restrict__ offs,
                                                         const uint32_t *__restrict__ members, uint8_t *__restrict__ out) {
  const uint32_t c = blockIdx.x;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (uint32_t i = offs[c]; i < offs[c + 1]; i++) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(regs + (size_t)members[i] * HLL_M) + threadIdx.x);
    acc.x = __vmaxu4(acc.x, v.x); acc.y = __vmaxu4(acc.y, v.y); acc.z = __vmaxu4(acc.z, v.z); acc.w = __vmaxu4(acc.w, v.w);
  }
  reinterpret_cast<uint4 *>(out + (size_t)c * HLL_M)[threadIdx.x] = acc;
}

// keys of the taxa that are still sparse and belong to the tree → (code << 32 | preorder of the taxon)
__global__ void __launch_bounds__(256) k_clade_sort_keys(const unsigned long long *__restrict__ slots, uint64_t cap,
                                                         const uint8_t *__restrict__ dense_flag, const uint32_t *__restrict__ pre_of_taxon,
                                                         unsigned long long *__restrict__ out, uint64_t out_cap,
                                                         unsigned long long *cursor) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t n_round = (cap + stride - 1) / stride * stride;          // whole warps stay in the loop (ballot below)
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    unsigned long long sk = 0;
    bool have = false;
    if (i < cap) {
      const unsigned long long key = slots[i];
      if (key) {
        const uint32_t taxon = (uint32_t)(key >> 32) - 1;
        if (!dense_flag[taxon]) {
          const uint32_t pre = pre_of_taxon[taxon];
          if (pre != NONE32) { sk = ((unsigned long long)(uint32_t)key << 32) | pre; have = true; }
        }
      }
    }
    const uint32_t vote = __ballot_sync(0xFFFFFFFFu, have);
    if (vote) {
      const uint32_t lane = threadIdx.x & 31;
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(cursor, (unsigned long long)__popc(vote));
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (have) {
        const unsigned long long at = base + __popc(vote & ((1u << lane) - 1));
        if (at < out_cap) out[at] = sk;
      }
    }
  }
}

// adjacent equal codes: book the pair at the lowest common ancestor of its two taxa
__global__ void __launch_bounds__(256) k_clade_dups(const unsigned long long *__restrict__ sorted, uint64_t n,
                                                    const uint32_t *__restrict__ node_of_pre, const uint32_t *__restrict__ parent_c,
                                                    const uint32_t *__restrict__ depth_c, const int32_t *__restrict__ sid_of_node,
                                                    uint32_t *dup) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += stride) {
    const unsigned long long a = sorted[i - 1], b = sorted[i];
    if ((a >> 32) != (b >> 32)) continue;
    uint32_t x = node_of_pre[(uint32_t)a], y = node_of_pre[(uint32_t)b];
    uint32_t dx = depth_c[x], dy = depth_c[y];
    while (dx > dy && x != NONE32) { x = parent_c[x]; dx--; }
    while (dy > dx && y != NONE32) { y = parent_c[y]; dy--; }
    while (x != y && x != NONE32 && y != NONE32) { x = parent_c[x]; y = parent_c[y]; }
    if (x == NONE32 || y == NONE32 || x != y) continue;               // different trees: no clade holds both
    const int32_t s = sid_of_node[x];
    if (s >= 0) atomicAdd(dup + (size_t)s * 64 + min(enc_rank((uint32_t)(a >> 32)), 63u), 1u);
  }
}

}  // namespace

void launch_clade_max_batch(const uint8_t *regs, const uint32_t *d_offs, const uint32_t *d_members, uint32_t n_clades,
                            uint8_t *d_out, cudaStream_t stream) {
  if (n_clades) k_clade_max_batch<<<n_clades, 256, 0, stream>>>(regs, d_offs, d_members, d_out);
}

// 0 = ok, 1 = not enough device memory (the caller falls back to one union per clade), 2 = CUDA error
int sparse_clade_dups(const unsigned long long *slots, uint64_t cap, const uint8_t *dense_flag, uint64_t n_keys_upper,
                      const uint32_t *d_pre_of_taxon, const uint32_t *d_node_of_pre, const uint32_t *d_parent_c,
                      const uint32_t *d_depth_c, const int32_t *d_sid_of_node, uint32_t *d_dup, int n_sm, cudaStream_t stream) {
  if (n_keys_upper < 2 || cap == 0) return 0;
  unsigned long long *buf[2] = {nullptr, nullptr}, *cursor = nullptr;
  void *temp = nullptr;
  size_t temp_bytes = 0;
  int rc = 0;
  auto cleanup = [&]() { cudaFree(buf[0]); cudaFree(buf[1]); cudaFree(cursor); cudaFree(temp); };
  cub::DoubleBuffer<unsigned long long> keys(nullptr, nullptr);
  if (cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, keys, (int64_t)n_keys_upper, 0, 64, stream) != cudaSuccess) {
    (void)cudaGetLastError();
    return 2;
  }
  if (cudaMalloc((void **)&buf[0], n_keys_upper * 8) != cudaSuccess || cudaMalloc((void **)&buf[1], n_keys_upper * 8) != cudaSuccess ||
      cudaMalloc((void **)&cursor, 8) != cudaSuccess || cudaMalloc(&temp, temp_bytes ? temp_bytes : 1) != cudaSuccess) {
    (void)cudaGetLastError();
    cleanup();
    return 1;
  }
  cudaMemsetAsync(cursor, 0, 8, stream);
  k_clade_sort_keys<<<n_sm * 8, 256, 0, stream>>>(slots, cap, dense_flag, d_pre_of_taxon, buf[0], n_keys_upper, cursor);
  unsigned long long n = 0;
  cudaMemcpyAsync(&n, cursor, 8, cudaMemcpyDeviceToHost, stream);
  if (cudaStreamSynchronize(stream) != cudaSuccess) { (void)cudaGetLastError(); cleanup(); return 2; }
  if (n > n_keys_upper) { cleanup(); return 2; }                       // more keys than the per-taxon counters promised
  if (n >= 2) {
    keys = cub::DoubleBuffer<unsigned long long>(buf[0], buf[1]);
    if (cub::DeviceRadixSort::SortKeys(temp, temp_bytes, keys, (int64_t)n, 0, 64, stream) != cudaSuccess) rc = 2;
    if (!rc) k_clade_dups<<<n_sm * 8, 256, 0, stream>>>(keys.Current(), n, d_node_of_pre, d_parent_c, d_depth_c, d_sid_of_node, d_dup);
    if (cudaStreamSynchronize(stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = 2;
  }
  cleanup();
  return rc;
}

}  // namespace kuq
