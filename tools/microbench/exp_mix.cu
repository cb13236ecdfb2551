// Microbenchmark (B200): throughput of the attention softmax inner loop (exp_chunk / chunk_max of softmax_math.cuh)
// with 1, 2, 4 warps per SM sub-partition, with and without the polynomial exp2 share.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../simple-tad_b200/csrc/softmax_math.cuh"
using namespace stad;

template <bool kPoly, bool kMax>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* out, float* sink, float c, float neg_m) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t sv[2][32];
  for (int q = 0; q < 2; ++q)
    for (int i = 0; i < 32; ++i) sv[q][i] = __float_as_uint(-0.01f * (lane + i + q));
  float a0 = 0, a1 = 0, b0 = 0, b1 = 0, mx = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t pk0[16], pk1[16];
    if (kMax) mx += fmaxf(chunk_max(sv[0]), chunk_max(sv[1]));
    exp_chunk<kPoly>(sv[0], c, neg_m, a0, a1, pk0);
    exp_chunk<kPoly>(sv[1], c, neg_m, b0, b1, pk1);
    // feed back so nothing is hoisted: next scores depend on this iteration's packed output
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      sv[0][2 * i] ^= pk0[i] & 0x00010001u;
      sv[1][2 * i + 1] ^= pk1[i] & 0x00010001u;
    }
  }
  const long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
  if (a0 + a1 + b0 + b1 + mx == 123.456f) sink[0] = a0;
}

template <bool kPoly, bool kMax>
void run(const char* name, int warps, long long* out, float* sink) {
  const int iters = 4000;
  long long h[16];
  for (int rep = 0; rep < 2; ++rep) {
    k<kPoly, kMax><<<148, warps * 32>>>(iters, out, sink, 0.18f, 0.5f);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mxc = 0;
  for (int w = 0; w < warps; ++w) mxc = h[w] > mxc ? h[w] : mxc;
  const double per_chunk = (double)mxc / (iters * 2.0);
  const double elems = (double)iters * 64 * 32 * warps;
  printf("%-28s %2d warps: %7.1f clk per 32-col chunk per warp; %6.2f elements/clk/SM (MUFU-only bound 16)\n", name, warps,
         per_chunk, elems / mxc);
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 148 * 16 * 8); cudaMalloc(&sink, 4);
  for (int w : {4, 8, 16}) run<false, false>("exp only, all MUFU", w, out, sink);
  for (int w : {4, 8, 16}) run<true, false>("exp only, poly share", w, out, sink);
  for (int w : {4, 8, 16}) run<false, true>("max + exp, all MUFU", w, out, sink);
  for (int w : {4, 8, 16}) run<true, true>("max + exp, poly share", w, out, sink);
  return 0;
}
