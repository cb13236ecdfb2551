// Microbenchmark (B200): cost of issuing tcgen05.mma from one thread, for the instruction shapes the GEMM / attention
// kernels use, under three source patterns:
//   0: `if (lane == 0) { loop }`            (per-lane values: ptxas emits ELECT + R2UR.BROADCAST + BRA.U.ANY per MMA)
//   1: whole warp runs the loop, `if (elect_one()) mma`   (values warp-uniform by construction)
//   2: like 1, operands hoisted through __shfl_sync(.., 0) first
#include <cstdio>
#include <cuda_runtime.h>
#include "../../simple-tad_b200/csrc/ptx.cuh"
using namespace stad;

template <int PATTERN, int N, bool TS>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) { tmem_alloc<512>(&slot); tmem_relinquish(); }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tmem_base = slot;
  constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, TS ? 1 : 0);
  if (warp == 1) {
    if (PATTERN == 2) tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t da = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t db = make_smem_desc_sw128(smem_u32(smem + 16384), TS ? 0 : 16, 1024);
    long long t0 = 0, t1 = 0;
    if (PATTERN == 0) {
      if (lane == 0) {
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (TS) umma_ts(tmem_base + 256, tmem_base + kk * 8, db + (kk * 2048 >> 4), idesc, 1);
            else umma_ss(tmem_base, da + 2 * kk, db + 2 * kk, idesc, 1);
          }
        }
        t1 = clock64();
        umma_commit(&bar);
      }
    } else {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (TS) umma_ts(tmem_base + 256, tmem_base + kk * 8, db + (kk * 2048 >> 4), idesc, 1);
            else umma_ss(tmem_base, da + 2 * kk, db + 2 * kk, idesc, 1);
          }
        }
        __syncwarp();
      }
      t1 = clock64();
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
    }
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (lane == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(slot); }
}

template <int PATTERN, int N, bool TS>
void run(const char* name, long long* out) {
  const int iters = 2000;
  cudaFuncSetAttribute(k<PATTERN, N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  long long h[2];
  for (int rep = 0; rep < 2; ++rep) { k<PATTERN, N, TS><<<148, 128, 64 * 1024>>>(iters, out); if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s: error %s\n", name, cudaGetErrorString(cudaGetLastError())); return; } }
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-40s pattern %d: issue %.1f clk/MMA, complete %.1f clk/MMA (tensor-pipe floor %d)\n", name, PATTERN,
         (double)h[0] / (iters * 4.0), (double)h[1] / (iters * 4.0), 128 * N / 256);
}

int main() {
  long long* out; cudaMalloc(&out, 148 * 16);
  run<0, 256, false>("SS M128 N256 K16", out); run<1, 256, false>("SS M128 N256 K16", out); run<2, 256, false>("SS M128 N256 K16", out);
  run<0, 128, false>("SS M128 N128 K16", out); run<1, 128, false>("SS M128 N128 K16", out); run<2, 128, false>("SS M128 N128 K16", out);
  run<0, 64, false>("SS M128 N64 K16", out);  run<1, 64, false>("SS M128 N64 K16", out);
  run<0, 64, true>("TS M128 N64 K16 (A=TMEM, B MN-major)", out); run<1, 64, true>("TS M128 N64 K16 (A=TMEM, B MN-major)", out); run<2, 64, true>("TS M128 N64 K16 (A=TMEM, B MN-major)", out);
  return 0;
}
