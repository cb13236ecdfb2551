// Microbenchmark (B200): issue rate of MUFU.EX2 in its fp32, packed f16x2 and packed bf16x2 forms — does the packed
// form deliver two exponentials per MUFU slot?  (The attention softmax is bound by 16 fp32 exps / clk / SM.)
// Each thread runs 8 independent dependency chains so that latency is hidden; 4 / 8 / 16 warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ float ex2_f32(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* out, uint32_t* sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t v[8];
  for (int i = 0; i < 8; ++i) v[i] = MODE == 0 ? __float_as_uint(-0.001f * (lane + i)) : 0x80008000u + lane + i;  // tiny negatives
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) v[i] = __float_as_uint(ex2_f32(__uint_as_float(v[i])) - 1.0f);
      else if (MODE == 1) v[i] = ex2_f16x2(v[i]) ^ 0x3c003c00u;   // 1.0 -> 0: keeps the argument small
      else v[i] = ex2_bf16x2(v[i]) ^ 0x3f803f80u;
    }
  }
  const long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
  uint32_t s = 0;
  for (int i = 0; i < 8; ++i) s ^= v[i];
  if (s == 0x12345678u) sink[0] = s;
}

template <int MODE>
void run(const char* name, int warps, long long* out, uint32_t* sink) {
  const int iters = 20000;
  long long h[16];
  for (int rep = 0; rep < 2; ++rep) {
    k<MODE><<<148, warps * 32>>>(iters, out, sink);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mxc = 0;
  for (int w = 0; w < warps; ++w) mxc = h[w] > mxc ? h[w] : mxc;
  const double instr = (double)iters * 8 * warps;           // warp-level MUFU instructions per SM
  const double per_lane = instr * 32 / mxc;                 // thread-level MUFU ops / clk / SM
  printf("%-22s %2d warps: %6.2f MUFU lane-ops/clk/SM = %6.2f exps/clk/SM\n", name, warps, per_lane, per_lane * (MODE == 0 ? 1 : 2));
}

int main() {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, 148 * 16 * 8); cudaMalloc(&sink, 4);
  for (int w : {4, 8, 16}) run<0>("ex2.approx.ftz.f32", w, out, sink);
  for (int w : {4, 8, 16}) run<1>("ex2.approx.f16x2", w, out, sink);
  for (int w : {4, 8, 16}) run<2>("ex2.approx.ftz.bf16x2", w, out, sink);
  return 0;
}
