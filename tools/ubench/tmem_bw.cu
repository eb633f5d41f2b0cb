// Microbenchmark: tcgen05.ld throughput per SM (how many bytes/clk can the epilogue/softmax warps pull out of TMEM?)
#include <cstdio>
#include <cuda_runtime.h>
#include "../../robustsq_whisper_b200/csrc/tc_ptx.cuh"
namespace tsw { void set_error(const char*, ...) {} int sm_count() { return 148; } EncodeTiledFn get_encode_fn() { return nullptr; } }
using namespace tsw;

__global__ void __launch_bounds__(512, 1) tmem_read_kernel(int iters, int nwarps_active, float* out, long long* cycles) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps_active) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int c = 0; c < 512; c += 32) {
        float v[32];
        tmem_ld32(base + c, v);
        acc += v[0] + v[31];
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}

int main() {
  float* out; long long* cyc; long long h[148];
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int nw : {1, 4, 8, 16}) {
    const int iters = 200;
    tmem_read_kernel<<<148, 512>>>(iters, nw, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
    const double bytes = (double)nw * iters * 16 * 32 * 32 * 4;  // per SM
    printf("warps=%2d: %lld cycles, %.1f B/clk/SM (%s)\n", nw, h[0], bytes / h[0], cudaGetErrorString(e));
  }
  return 0;
}
