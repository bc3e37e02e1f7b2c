// Microbenchmark: shared->global TMA bulk-store throughput on one GPU as a function of
// chunk size, bulk stores in flight per warp (ring depth) and warps per CTA (1 CTA per SM).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/tma_store_bench.cu -o tma_store_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int DEPTH>
__global__ void __launch_bounds__(512, 1) k(double* out, size_t total_chunks, int chunk_doubles, int fill, int dst_off, int src_off) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double* buf = smem + (size_t)warp * DEPTH * (chunk_doubles + 16) + src_off;
  out += dst_off;
  int slot = 0;
  for (size_t c = (size_t)blockIdx.x * nw + warp; c < total_chunks; c += (size_t)gridDim.x * nw) {
    double* b = buf + slot * (chunk_doubles + 16);
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
    __syncwarp();
    if (fill) for (int i = lane; i < chunk_doubles / 2; i += 32) reinterpret_cast<double2*>(b)[i] = make_double2(1.0, 2.0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * chunk_doubles), "r"(s32(b)), "r"(chunk_doubles * 8) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    slot = (slot + 1) % DEPTH;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
template <int DEPTH>
float run(double* out, size_t bytes, int warps, int chunk_bytes, int fill, int dst_off = 0, int src_off = 0) {
  int cd = chunk_bytes / 8;
  size_t chunks = bytes / chunk_bytes;
  size_t smem = (size_t)warps * DEPTH * (chunk_bytes + 128);
  chunks -= 1;
  if (smem > 227 * 1024) return -1;
  cudaFuncSetAttribute(k<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 2; ++i) k<DEPTH><<<148, warps * 32, smem>>>(out, chunks, cd, fill, dst_off, src_off);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) k<DEPTH><<<148, warps * 32, smem>>>(out, chunks, cd, fill, dst_off, src_off);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  if (cudaGetLastError() != cudaSuccess) return -2;
  return ms / 5;
}
int main() {
  size_t bytes = (size_t)6 << 30;
  double* out; cudaMalloc(&out, bytes);
  cudaMemset(out, 0, bytes);
  printf("latency study (depth 1, fill 0): warps chunkB ms GB/s us_per_chunk_per_warp\n");
  for (int w : {1, 2, 4, 8})
    for (int cb : {1024, 4096, 11648, 23296}) {
      float ms = run<1>(out, bytes / 4, w, cb, 0, 0, 0);
      double chunks_per_warp = (double)(bytes / 4 / cb) / (148.0 * w);
      printf("%3d %6d %7.3f %7.0f %7.3f\n", w, cb, ms, bytes / 4 / ms / 1e6, ms * 1e3 / chunks_per_warp);
    }
  return 0;
}
