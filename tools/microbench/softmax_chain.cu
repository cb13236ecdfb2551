// Microbenchmark (B200): the per-tile chain of one attention softmax warp WITHOUT the tensor core and without barriers —
//   tcgen05.ld of a score row (NCH x 32 fp32 columns) -> row max -> exp2(s*c - m) + row sum + bf16 pack per chunk ->
//   tcgen05.st of the packed P row -> wait
// for W softmax warps per SM sub-partition.  It bounds what the softmax side of attention_kernel can sustain for a
// given (warps per sub-partition, keys per tile) design: today 2 x 128; candidates 3 x 96 (three query tiles per CTA,
// P written over S) and 4 x 64.  Output: clocks per tile per warp and score elements / clk / SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_chain softmax_chain.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../simple-tad_b200/csrc/softmax_math.cuh"
using namespace stad;

// exp2 + pack of a chunk WITHOUT the row-sum adds (what the loop costs if the row sum came out of the tensor core: a
// ones column appended to V makes P·[V | 1] deliver it in a 65th accumulator column)
STAD_DEVICE void exp_chunk_nosum(const uint32_t (&s)[32], float c, float neg_m, uint32_t (&pk)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float x0, x1, e0, e1;
    fma2(x0, x1, __uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]), c, c, neg_m, neg_m);
    if (kPolyPeriod > 0 && (i % (kPolyPeriod > 0 ? kPolyPeriod : 1)) == (kPolyPeriod - 1)) {
      exp2_poly2(e0, e1, x0, x1);
    } else {
      e0 = ex2(x0);
      e1 = ex2(x1);
    }
    pk[i] = pack_bf16(e0, e1);
  }
}

// exp2 + row sum + pack of a 16-column half chunk (a half-row design: two threads per query row, 48 keys each)
STAD_DEVICE void exp_half_chunk(const uint32_t (&s)[16], float c, float neg_m, float& acc0, float& acc1, uint32_t (&pk)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float x0, x1, e0, e1;
    fma2(x0, x1, __uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]), c, c, neg_m, neg_m);
    if (kPolyPeriod > 0 && (i % (kPolyPeriod > 0 ? kPolyPeriod : 1)) == (kPolyPeriod - 1)) {
      exp2_poly2(e0, e1, x0, x1);
    } else {
      e0 = ex2(x0);
      e1 = ex2(x1);
    }
    add2(acc0, acc1, acc0, acc1, e0, e1);
    pk[i] = pack_bf16(e0, e1);
  }
}
STAD_DEVICE void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// MODE bit 0: no row-sum adds; bit 1: no row max (a reference maximum is assumed known, e.g. from the previous tile);
// bit 2: one extra 16-column half chunk per tile (keys per tile = 32 NCH + 16)
template <int NCH, int THREADS, int MODE>
__global__ void __launch_bounds__(THREADS, 1) k(int iters, long long* out, float* sink, float c) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc<512>(&slot); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  // warps w, w+4, w+8, ... share lane quarter w & 3; each gets its own column range (S at col0, P over the same columns)
  const int group = warp >> 2;
  const uint32_t col0 = static_cast<uint32_t>(group * (NCH * 32 + ((MODE & 4) ? 16 : 0))) & 511u;
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + col0;
  {  // something finite to read
    uint32_t z[32];
    for (int i = 0; i < 32; ++i) z[i] = __float_as_uint(-0.01f * (lane + i));
    for (int ch = 0; ch < NCH; ++ch) tmem_st32(base + ch * 32, z);
    tmem_st_wait();
  }
  float a0 = 0.f, a1 = 0.f, m_run = -1e30f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t s[NCH][32];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) tmem_ld32(base + ch * 32, s[ch]);
    uint32_t sh[16];
    if constexpr (MODE & 4) tmem_ld16(base + NCH * 32, sh);
    tmem_ld_wait();
    if constexpr (!(MODE & 2)) {
      float mx = chunk_max(s[0]);
#pragma unroll
      for (int ch = 1; ch < NCH; ++ch) mx = fmaxf(mx, chunk_max(s[ch]));
      m_run = fmaxf(m_run, mx * c);
    } else {
      m_run = fmaxf(m_run, __uint_as_float(s[0][0]) * c);  // a known reference: no pass over the row
    }
    const float neg_m = -m_run;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      uint32_t pk[16];
      if constexpr (MODE & 1) exp_chunk_nosum(s[ch], c, neg_m, pk);
      else exp_chunk<true>(s[ch], c, neg_m, a0, a1, pk);
      tmem_st16(base + ch * 16, pk);  // P (bf16 pairs) over the first half of the chunk's own S columns
    }
    if constexpr (MODE & 4) {
      uint32_t pk8[8];
      exp_half_chunk(sh, c, neg_m, a0, a1, pk8);
      tmem_st8(base + NCH * 16, pk8);
    }
    tmem_st_wait();
  }
  const long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 32 + warp] = t1 - t0;
  if (a0 + a1 + m_run == 123.456f) sink[0] = a0;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(slot); }
}

template <int NCH, int WPS, int MODE = 0>
void run(long long* out, float* sink) {
  constexpr int THREADS = WPS * 4 * 32;
  const int iters = 3000;
  long long h[32];
  for (int rep = 0; rep < 2; ++rep) {
    k<NCH, THREADS, MODE><<<148, THREADS>>>(iters, out, sink, 0.18f);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(cudaGetLastError())); return; }
  }
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mxc = 0;
  for (int w = 0; w < WPS * 4; ++w) mxc = h[w] > mxc ? h[w] : mxc;
  const double per_tile = static_cast<double>(mxc) / iters;
  const int keys = NCH * 32 + ((MODE & 4) ? 16 : 0);
  const double elems = static_cast<double>(iters) * keys * 32 * WPS * 4;
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, k<NCH, THREADS, MODE>);
  printf("%d warps/sub-partition x %3d keys/tile%s%s: %7.1f clk per tile per warp, %6.2f elements/clk/SM, "
         "%6.1f clk per 256 x 128 scores  (%d regs/thread)\n",
         WPS, keys, (MODE & 1) ? ", no row sum" : "", (MODE & 2) ? ", no row max" : "", per_tile, elems / mxc,
         256.0 * 128.0 / (elems / mxc), fa.numRegs);
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 148 * 32 * 8); cudaMalloc(&sink, 4);
  run<4, 1>(out, sink);
  run<4, 2>(out, sink);   // today's kernel: 2 x 128
  run<3, 2>(out, sink);
  run<3, 3>(out, sink);   // candidate: three query tiles per CTA, 96-key tiles
  run<2, 3>(out, sink);
  run<2, 4>(out, sink);   // four query tiles, 64-key tiles
  run<4, 3>(out, sink);   // 3 x 128 (TMEM would not allow it; register-pressure reference)
  // which instruction group the chain is bound by
  run<4, 2, 1>(out, sink);
  run<4, 2, 2>(out, sink);
  run<4, 2, 3>(out, sink);
  run<3, 2, 1>(out, sink);
  run<3, 2, 3>(out, sink);
  // round 2: lazy reference (no row max in the steady state): today's 2 x 96 and the candidates with more warps per
  // sub-partition (half-row: two threads per query row, 48 keys each; four streams of 64-key tiles; 3 x 64; 4 x 32)
  run<3, 2, 2>(out, sink);
  run<1, 4, 2 | 4>(out, sink);
  run<2, 4, 2>(out, sink);
  run<2, 3, 2>(out, sink);
  run<1, 4, 2>(out, sink);
  run<1, 3, 2 | 4>(out, sink);
  run<3, 3, 2>(out, sink);
  return 0;
}
