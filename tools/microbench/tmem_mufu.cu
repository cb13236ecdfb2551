// Microbenchmark (B200): TMEM load/store throughput per warp / per SM and MUFU.EX2 throughput, alone and overlapped.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_mufu tmem_mufu.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../simple-tad_b200/csrc/ptx.cuh"
using namespace stad;

// mode bit0: LDTM loop on warps [0, nld)   bit1: MUFU loop on warps [nld, nld+nmu)   bit2: STTM instead of LDTM
__global__ void __launch_bounds__(512, 1) k(int mode, int nld, int nmu, int iters, int cols_per_ld, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc<512>(&slot); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  long long t0 = clock64();
  float acc = 0.f;
  if (warp < nld && (mode & 1)) {
    uint32_t v[32];
    if (mode & 4) {
      for (int i = 0; i < 32; ++i) v[i] = lane + i;
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_st32(base + ((it & 1) * 128 + c * 32), v);
        tmem_st_wait();
      }
    } else {
      for (int it = 0; it < iters; ++it) {
        if (cols_per_ld == 128) {  // four x32 loads in flight, one wait (what the attention kernel does)
          uint32_t w1[32], w2[32], w3[32];
          tmem_ld32(base + (it & 1) * 128, v);
          tmem_ld32(base + (it & 1) * 128 + 32, w1);
          tmem_ld32(base + (it & 1) * 128 + 64, w2);
          tmem_ld32(base + (it & 1) * 128 + 96, w3);
          tmem_ld_wait();
          acc += __uint_as_float(v[0] ^ w1[5] ^ w2[9] ^ w3[31]);
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (cols_per_ld == 32) tmem_ld32(base + ((it & 1) * 128 + c * 32), v);
            else { uint32_t (&h)[16] = *reinterpret_cast<uint32_t(*)[16]>(v); tmem_ld16(base + ((it & 1) * 128 + c * 32), h); }
            tmem_ld_wait();
            acc += __uint_as_float(v[0] ^ v[7]);
          }
        }
      }
    }
  } else if (warp >= nld && warp < nld + nmu && (mode & 2)) {
    float x[16];
    for (int i = 0; i < 16; ++i) x[i] = -0.001f * (lane + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fast_exp2(x[i]) - 1.0f;
    }
    for (int i = 0; i < 16; ++i) acc += x[i];
  }
  long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(slot); }
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 148 * 16 * 8); cudaMalloc(&sink, 4);
  long long h[16];
  struct Cfg { const char* name; int mode, nld, nmu, cols; } cfgs[] = {
    {"LDTM x32, 1 warp", 1, 1, 0, 32}, {"LDTM x32, 4 warps (1/SMSP)", 1, 4, 0, 32}, {"LDTM x32, 8 warps (2/SMSP)", 1, 8, 0, 32},
    {"LDTM x16, 4 warps", 1, 4, 0, 16}, {"LDTM x16, 8 warps", 1, 8, 0, 16},
    {"LDTM 4x32 then wait, 1 warp", 1, 1, 0, 128}, {"LDTM 4x32 then wait, 4 warps", 1, 4, 0, 128}, {"LDTM 4x32 then wait, 8 warps", 1, 8, 0, 128},
    {"LDTM 4x32+wait 8 warps + MUFU 8", 3, 8, 8, 128},
    {"STTM x32, 4 warps", 5, 4, 0, 32}, {"STTM x32, 8 warps", 5, 8, 0, 32},
    {"MUFU, 4 warps", 2, 0, 4, 32}, {"MUFU, 8 warps", 2, 0, 8, 32},
    {"LDTM 4 warps + MUFU 4 warps", 3, 4, 4, 32}, {"LDTM 8 warps + MUFU 8 warps", 3, 8, 8, 32},
  };
  const int iters = 2000;
  for (auto& c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) {
      k<<<148, 512>>>(c.mode, c.nld, c.nmu, iters, c.cols, out, sink);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    }
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    const int nw = c.nld + c.nmu;
    printf("%-34s:", c.name);
    double ld_clk = 0, mu_clk = 0;
    for (int w = 0; w < nw; ++w) { if (w < c.nld) ld_clk = h[w] > ld_clk ? h[w] : ld_clk; else mu_clk = h[w] > mu_clk ? h[w] : mu_clk; }
    if (c.nld) {
      double bytes_per_warp = (double)iters * 4 * 32 * (c.cols == 128 ? 32 : c.cols) * 4;
      printf("  tmem: %.1f clk per %d-col op/warp, %.1f B/clk/warp, %.1f B/clk/SM", ld_clk / (iters * 4.0), c.cols,
             bytes_per_warp / ld_clk, bytes_per_warp * c.nld / ld_clk);
    }
    if (c.nmu) {
      double ex_per_warp = (double)iters * 8 * 16 * 32;
      printf("  mufu: %.2f ex2/clk/warp, %.2f ex2/clk/SM", ex_per_warp / mu_clk, ex_per_warp * c.nmu / mu_clk);
    }
    printf("\n");
  }
  return 0;
}
