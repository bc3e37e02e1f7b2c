// Microbenchmark 3: can the read/write mixing penalty be removed by batching?
// Per batch: kernel A touches the batch's input ids (pure read phase, L2 persisting window), kernel B runs the
// TMA store stream and re-reads the ids (now L2 hits).  Compare with the mixed single-kernel baseline (mode 1).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k_touch(const int* ids, size_t n, int* sink) {
  int acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc += __ldg(ids + i);
  if (acc == 123456789) *sink = acc;
}
__global__ void __launch_bounds__(256, 1) k_emit(double* out, size_t c_begin, size_t c_end, int chunk_doubles, const int* ids, int* sink) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double* b = smem + (size_t)warp * chunk_doubles;
  int acc = 0;
  for (size_t c = c_begin + (size_t)blockIdx.x * nw + warp; c < c_end; c += (size_t)gridDim.x * nw) {
    for (int a = 0; a < 13; ++a) acc += __ldg(ids + (c * 13 + a) * 32 + lane);
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
  size_t nid = chunks * 13 * 32;
  int* ids; cudaMalloc(&ids, nid * 4); cudaMemset(ids, 0, nid * 4);
  int* sink; cudaMalloc(&sink, 4);
  size_t smem = 8 * (size_t)cb;
  cudaFuncSetAttribute(k_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaStream_t st; cudaStreamCreate(&st);
  for (int nb : {1, 8, 16, 32, 64}) {
    for (int persist = 0; persist < 2; ++persist) {
      if (nb == 1 && persist) continue;
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist ? (48 << 20) : 0);
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      float best = 1e9;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a, st);
        for (int k = 0; k < nb; ++k) {
          size_t c0 = chunks * k / nb, c1 = chunks * (k + 1) / nb;
          if (nb > 1) {
            if (persist) {
              cudaStreamAttrValue attr = {};
              attr.accessPolicyWindow.base_ptr = (void*)(ids + c0 * 13 * 32);
              attr.accessPolicyWindow.num_bytes = (c1 - c0) * 13 * 32 * 4;
              attr.accessPolicyWindow.hitRatio = 1.0f;
              attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
              attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
              cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
            }
            k_touch<<<148 * 4, 256, 0, st>>>(ids + c0 * 13 * 32, (c1 - c0) * 13 * 32, sink);
          }
          k_emit<<<148, 256, smem, st>>>(out, c0, c1, cb / 8, ids, sink);
        }
        cudaEventRecord(b, st); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
      }
      printf("batches %3d persist %d: %.3f ms  %.0f GB/s written  (%s)\n", nb, persist, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
      if (persist) { cudaCtxResetPersistingL2Cache(); cudaStreamAttrValue attr = {}; attr.accessPolicyWindow.num_bytes = 0; cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr); }
    }
  }
  return 0;
}
