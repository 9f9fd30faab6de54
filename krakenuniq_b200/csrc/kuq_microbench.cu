// kuq_microbench.cu — the second roofline denominator of SURVEY.md §8(d): how many random 32-byte sectors per second
// this GPU's HBM serves.  k_lookup's bin probes are exactly that access pattern (a 12-byte record inside a random
// 32-byte sector of a multi-GB array), so its ceiling is not the copy bandwidth of MEASURED_PEAKS.json but this
// number.  Same launch shape as k_lookup: 256 threads per CTA, 8 CTAs per SM; every thread issues `ILP`
// independent loads per round (no dependent chain, i.e. the memory system's best case for this granularity).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/kuq.h"

namespace {

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 31; x *= 0x7fb5d329728ea185ull;
  x ^= x >> 27; x *= 0x81dadef4bc2dd44dull;
  x ^= x >> 33;
  return x;
}

// every thread reads `rounds * ILP` sectors at hashed positions; GRAN = bytes touched per access (8 or 32)
template <int ILP, int GRAN>
__global__ void __launch_bounds__(256, 8) k_random_gather(const uint8_t *__restrict__ buf, uint64_t n_sectors,
                                                          uint32_t rounds, uint64_t seed, unsigned long long *sink) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t n_threads = (uint64_t)gridDim.x * blockDim.x;
  uint64_t acc = 0;
  for (uint32_t r = 0; r < rounds; r++) {
    uint64_t v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) {
      const uint64_t h = mix(seed + (uint64_t)(r * ILP + j) * n_threads + tid);
      // 128-bit multiply-shift maps the hash onto [0, n_sectors) without a modulo
      const uint64_t s = __umul64hi(h, n_sectors);
      if (GRAN == 32) {
        const ulonglong4 *p = reinterpret_cast<const ulonglong4 *>(buf + s * 32);
        const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2 *>(p));
        const ulonglong2 b = __ldg(reinterpret_cast<const ulonglong2 *>(p) + 1);
        v[j] = a.x ^ a.y ^ b.x ^ b.y;
      } else {
        v[j] = __ldg(reinterpret_cast<const uint64_t *>(buf + s * 32 + ((h >> 5) & 24)));
      }
    }
#pragma unroll
    for (int j = 0; j < ILP; j++) acc ^= v[j];
  }
  if (acc == 0x1234567ull) atomicAdd(sink, 1ull);   // keeps the loads alive
}

}  // namespace

extern "C" int kuq_random_gather_peak(int device, uint64_t buffer_bytes, uint64_t n_sectors_to_read, uint32_t bytes_per_access,
                                      double *gsectors_per_s, double *gbytes_per_s, double *kernel_ms) {
  if (bytes_per_access != 8 && bytes_per_access != 32) return KUQ_E_INVALID_ARG;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) { (void)cudaGetLastError(); return KUQ_E_NO_DEVICE; }
  if (cudaSetDevice(device) != cudaSuccess) return KUQ_E_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return KUQ_E_CUDA;
  buffer_bytes &= ~31ull;
  if (buffer_bytes < (1ull << 20)) return KUQ_E_INVALID_ARG;
  uint8_t *buf = nullptr;
  unsigned long long *sink = nullptr;
  if (cudaMalloc((void **)&buf, buffer_bytes) != cudaSuccess) { (void)cudaGetLastError(); return KUQ_E_NOMEM; }
  cudaMalloc((void **)&sink, 8);
  cudaMemset(buf, 0x5a, buffer_bytes);
  cudaMemset(sink, 0, 8);
  constexpr int ILP = 4;
  const int grid = prop.multiProcessorCount * 8;
  const uint64_t n_threads = (uint64_t)grid * 256;
  uint32_t rounds = (uint32_t)((n_sectors_to_read + n_threads * ILP - 1) / (n_threads * ILP));
  if (rounds == 0) rounds = 1;
  const uint64_t n_read = (uint64_t)rounds * ILP * n_threads;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {            // rep 0 = warm-up (TLB, clocks)
    cudaEventRecord(e0);
    if (bytes_per_access == 32)
      k_random_gather<ILP, 32><<<grid, 256>>>(buf, buffer_bytes / 32, rounds, 0x9E3779B97F4A7C15ull * (rep + 1), sink);
    else
      k_random_gather<ILP, 8><<<grid, 256>>>(buf, buffer_bytes / 32, rounds, 0x9E3779B97F4A7C15ull * (rep + 1), sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const cudaError_t err = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  cudaFree(sink);
  if (err != cudaSuccess) return KUQ_E_CUDA;
  if (gsectors_per_s) *gsectors_per_s = (double)n_read / (best * 1e-3) / 1e9;
  if (gbytes_per_s) *gbytes_per_s = (double)n_read * 32.0 / (best * 1e-3) / 1e9;
  if (kernel_ms) *kernel_ms = best;
  return KUQ_OK;
}
