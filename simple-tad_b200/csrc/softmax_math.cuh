// Register-level softmax arithmetic of the attention kernel (sm_100a): packed fp32 pairs (FFMA2 / FADD2), three-input
// max (FMNMX3), MUFU.EX2 and a polynomial exp2 that runs on the FMA pipe.  Shared with tools/microbench.
#pragma once
#include "ptx.cuh"

namespace stad {

// One pair of every kPolyPeriod pairs takes its exp2 on the FMA pipe instead of the MUFU (0 = never).
#ifndef STAD_ATT_POLY_PERIOD
#define STAD_ATT_POLY_PERIOD 4
#endif
constexpr int kPolyPeriod = STAD_ATT_POLY_PERIOD;

STAD_DEVICE float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
STAD_DEVICE float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for a pair, entirely on the FMA/ALU pipes: x = n + f, n = round(x), f in [-0.5, 0.5];
// 2^f by a degree-3 minimax polynomial (max relative error 7.5e-5, far below bf16 resolution of P), n added to the
// exponent field.  x is clamped to [-125, 127] so the exponent never wraps (the attention kernel's lazy reference max
// allows positive arguments; 2^127 trips its row-sum guard).
STAD_DEVICE void exp2_poly2(float& e0, float& e1, float x0, float x1) {
  constexpr float kMagic = 12582912.f;  // 1.5 * 2^23: x + kMagic holds round(x) in its low mantissa bits
  x0 = fminf(fmaxf(x0, -125.f), 127.f);
  x1 = fminf(fmaxf(x1, -125.f), 127.f);
  float r0, r1, n0, n1, f0, f1, p0, p1;
  add2(r0, r1, x0, x1, kMagic, kMagic);
  add2(n0, n1, r0, r1, -kMagic, -kMagic);
  fma2(f0, f1, n0, n1, -1.f, -1.f, x0, x1);
  fma2(p0, p1, f0, f1, 0.05517164245247841f, 0.05517164245247841f, 0.2426111400127411f, 0.2426111400127411f);
  fma2(p0, p1, p0, p1, f0, f1, 0.6932609677314758f, 0.6932609677314758f);
  fma2(p0, p1, p0, p1, f0, f1, 0.9999280571937561f, 0.9999280571937561f);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23));
}

// exp2(s * c - m) of one 32-column chunk of a score row -> 16 packed bf16 pairs; adds the fp32 row sum into acc.
template <bool kPoly>
STAD_DEVICE void exp_chunk(const uint32_t (&s)[32], float c, float neg_m, float& acc0, float& acc1,
                           uint32_t (&pk)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float x0, x1, e0, e1;
    fma2(x0, x1, __uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]), c, c, neg_m, neg_m);
    if (kPoly && kPolyPeriod > 0 && (i % (kPolyPeriod > 0 ? kPolyPeriod : 1)) == (kPolyPeriod - 1)) {
      exp2_poly2(e0, e1, x0, x1);
    } else {
      e0 = ex2(x0);
      e1 = ex2(x1);
    }
    add2(acc0, acc1, acc0, acc1, e0, e1);
    pk[i] = pack_bf16(e0, e1);
  }
}

// Row max of one 32-column chunk: four independent FMNMX3 chains (a single chain of 16 dependent three-input maxima is
// latency-bound: ~5 clk per step, and the row max sits on the critical path of every tile).
STAD_DEVICE float chunk_max(const uint32_t (&s)[32]) {
  float m[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    m[k] = max3(__uint_as_float(s[8 * k]), __uint_as_float(s[8 * k + 1]), __uint_as_float(s[8 * k + 2]));
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    m[k] = max3(m[k], __uint_as_float(s[8 * k + 3]), __uint_as_float(s[8 * k + 4]));
    m[k] = max3(m[k], __uint_as_float(s[8 * k + 5]), __uint_as_float(s[8 * k + 6]));
  }
  return fmaxf(max3(m[0], m[1], __uint_as_float(s[7])), max3(max3(m[2], m[3], __uint_as_float(s[15])),
                                                             __uint_as_float(s[23]), __uint_as_float(s[31])));
}

}  // namespace stad
