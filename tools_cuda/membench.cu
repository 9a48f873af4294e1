// HBM ceilings of this B200 for read / write mixes (the roofline denominator of the write-heavy 1x1 GEMMs):
//   plain 16-byte loads / stores, and TMA bulk stores (cp.async.bulk.global.shared::cta) from shared memory.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench membench.cu ; ./membench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// R reads and W writes of 16 bytes per thread-iteration (different streams of memory)
template <int R, int W>
__global__ void __launch_bounds__(256) mix_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint4 acc = make_uint4(1, 2, 3, 4);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const uint4 v = __ldcs(src + (size_t)r * n + i);
      acc.x ^= v.x; acc.y += v.y; acc.z ^= v.z; acc.w += v.w;
    }
#pragma unroll
    for (int w = 0; w < W; ++w) { acc.x += w; __stcs(dst + (size_t)w * n + i, acc); }
  }
}

// TMA bulk stores: every CTA streams `chunks` 16 KB chunks out of one shared-memory buffer, up to 8 groups in flight
__global__ void __launch_bounds__(128) bulk_store_kernel(uint8_t* __restrict__ dst, size_t chunks_total) {
  extern __shared__ __align__(128) uint8_t sm[];
  for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
    for (size_t c = blockIdx.x; c < chunks_total; c += gridDim.x) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * 16384), "r"(s), "r"(16384) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <typename F>
float time_ms(F f, int iters) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f(); f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / iters;
}

int main() {
  const size_t unit = (size_t)512 << 20;     // 512 MiB per stream (> 4x the 126 MB L2)
  const size_t n = unit / 16;
  uint4 *src, *dst;
  CK(cudaMalloc(&src, 4 * unit));
  CK(cudaMalloc(&dst, 4 * unit));
  CK(cudaMemset(src, 1, 4 * unit));
  CK(cudaMemset(dst, 0, 4 * unit));
  const int grid = 148 * 8;
#define RUN(R, W) { float ms = time_ms([&] { mix_kernel<R, W><<<grid, 256>>>(src, dst, n); }, 10); \
    printf("read %d : write %d   %8.1f us  %7.1f GB/s total  (%.1f read, %.1f write)\n", R, W, ms * 1e3, (R + W) * unit / ms / 1e6, R * unit / ms / 1e6, W * unit / ms / 1e6); }
  RUN(1, 0) RUN(2, 0) RUN(4, 0) RUN(0, 1) RUN(0, 2) RUN(1, 1) RUN(2, 1) RUN(3, 1) RUN(1, 2) RUN(1, 4) RUN(2, 2) RUN(4, 1)
  CK(cudaGetLastError());
  CK(cudaFuncSetAttribute(bulk_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  for (int ctas = 148; ctas <= 148 * 4; ctas *= 2) {
    float ms = time_ms([&] { bulk_store_kernel<<<ctas, 128, 16384>>>((uint8_t*)dst, 4 * unit / 16384); }, 10);
    printf("TMA bulk store, %4d CTAs, 8 x 16 KB in flight each: %8.1f us  %7.1f GB/s\n", ctas, ms * 1e3, 4.0 * unit / ms / 1e6);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
