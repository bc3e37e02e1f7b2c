// Microbenchmark 2: TMA bulk-store stream (8 warps/SM, 23 KB chunks) mixed with small reads per chunk:
// mode 0: no reads; mode 1: 13 x 128 B from 13 separate arrays (SoA id layout);
// mode 2: one contiguous 1664 B read (tile-blocked id layout); mode 3: mode 1 + 13 gathers per lane.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256, 1) k(double* out, size_t total_chunks, int chunk_doubles, int mode,
                                            const int* ids, size_t ldc, const double* x, size_t nx, int* sink) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double* b = smem + (size_t)warp * chunk_doubles;
  int acc = 0;
  const size_t stride = (size_t)gridDim.x * nw;
  if (mode == 6 || mode == 7) {
    // bursty reads: every G chunks the warp reads the ids of its next G chunks in one go
    // (mode 6: G = 16, mode 7: G = 64), then streams G chunks without touching DRAM for reads
    const int G = mode == 6 ? 16 : 64;
    size_t c = (size_t)blockIdx.x * nw + warp;
    while (c < total_chunks) {
      for (int k = 0; k < G; ++k) {
        const size_t cc = c + (size_t)k * stride;
        if (cc < total_chunks) for (int a = 0; a < 13; ++a) acc += __ldg(ids + (cc * 13 + a) * 32 + lane);
      }
      for (int k = 0; k < G && c < total_chunks; ++k, c += stride) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        for (int i = lane; i < chunk_doubles / 2; i += 32) reinterpret_cast<double2*>(b)[i] = make_double2(1.0, (double)acc);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * chunk_doubles), "r"(s32(b)), "r"(chunk_doubles * 8) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (acc == 123456789) *sink = acc;
    return;
  }
  if (mode == 4 || mode == 5) {
    size_t c = (size_t)blockIdx.x * nw + warp;
    int g[13], gn[13]; double xv[13];
    for (int a = 0; a < 13; ++a) g[a] = c < total_chunks ? __ldg(ids + a * ldc + c * 32 + lane) : 0;
    for (int a = 0; a < 13; ++a) gn[a] = c + stride < total_chunks ? __ldg(ids + a * ldc + (c + stride) * 32 + lane) : 0;
    if (mode == 5) for (int a = 0; a < 13; ++a) xv[a] = __ldg(x + ((size_t)(unsigned)g[a] % nx));
    for (; c < total_chunks; c += stride) {
      int g2[13]; double xn[13];
      if (mode == 5) for (int a = 0; a < 13; ++a) xn[a] = __ldg(x + ((size_t)(unsigned)gn[a] % nx));
      for (int a = 0; a < 13; ++a) g2[a] = c + 2 * stride < total_chunks ? __ldg(ids + a * ldc + (c + 2 * stride) * 32 + lane) : 0;
      for (int a = 0; a < 13; ++a) acc += g[a];
      if (mode == 5) { double v = 0; for (int a = 0; a < 13; ++a) v += xv[a]; acc += (int)v; }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      for (int i = lane; i < chunk_doubles / 2; i += 32) reinterpret_cast<double2*>(b)[i] = make_double2(1.0, (double)acc);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * chunk_doubles), "r"(s32(b)), "r"(chunk_doubles * 8) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      for (int a = 0; a < 13; ++a) { g[a] = gn[a]; gn[a] = g2[a]; if (mode == 5) xv[a] = xn[a]; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (acc == 123456789) *sink = acc;
    return;
  }
  for (size_t c = (size_t)blockIdx.x * nw + warp; c < total_chunks; c += (size_t)gridDim.x * nw) {
    int g[13];
    if (mode == 1 || mode == 3) { for (int a = 0; a < 13; ++a) g[a] = __ldg(ids + a * ldc + c * 32 + lane); }
    if (mode == 2) { for (int a = 0; a < 13; ++a) g[a] = __ldg(ids + (c * 13 + a) * 32 + lane); }
    double v = 0;
    if (mode == 3) { for (int a = 0; a < 13; ++a) v += __ldg(x + ((size_t)(unsigned)g[a] % nx)); }
    if (mode) { for (int a = 0; a < 13; ++a) acc += g[a]; acc += (int)v; }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    for (int i = lane; i < chunk_doubles / 2; i += 32) reinterpret_cast<double2*>(b)[i] = make_double2(1.0, (double)acc);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * chunk_doubles), "r"(s32(b)), "r"(chunk_doubles * 8) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (acc == 123456789) *sink = acc;
}
int main() {
  size_t bytes = (size_t)6 << 30; int cb = 23296;
  size_t chunks = bytes / cb;
  double* out; cudaMalloc(&out, bytes);
  size_t ldc = chunks * 32;
  int* ids; cudaMalloc(&ids, ldc * 13 * 4);
  int* h = (int*)malloc(ldc * 13 * 4);
  for (size_t a = 0; a < 13; ++a) for (size_t c = 0; c < ldc; ++c) h[a * ldc + c] = (int)((c * 4 + a * 2049) % 20000000);
  cudaMemcpy(ids, h, ldc * 13 * 4, cudaMemcpyHostToDevice);
  size_t nx = 21000000; double* x; cudaMalloc(&x, nx * 8); cudaMemset(x, 0, nx * 8);
  int* sink; cudaMalloc(&sink, 4);
  size_t smem = 8 * (size_t)cb;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int mode = 0; mode < 8; ++mode) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 2; ++i) k<<<148, 256, smem>>>(out, chunks, cb / 8, mode, ids, ldc, x, nx, sink);
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) k<<<148, 256, smem>>>(out, chunks, cb / 8, mode, ids, ldc, x, nx, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    printf("mode %d: %.3f ms  %.0f GB/s written  (%s)\n", mode, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
