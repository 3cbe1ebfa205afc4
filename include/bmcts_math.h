/*
 * bmcts_math.h -- elementary FP32 functions shared by the CUDA kernels
 * (planner-mcts_b200/csrc) and the CPU oracle (oracle/).
 *
 * Why this header exists: BASELINE.json's north_star demands bit-exact
 * collision / violation flags and terminal indices between the GPU path and the
 * CPU restatement.  Flags are threshold tests on FP32 geometry, so a 1-ulp
 * difference between glibc's sinf and CUDA's libdevice sinf (or an FMA that one
 * compiler contracts and the other does not) flips them.  Everything
 * transcendental is therefore implemented HERE, once, from +, *, fma, IEEE
 * division / sqrt, rint and floor only -- operations that round identically on
 * x86-64 SSE and on sm_100a.  The translation units that include this header
 * must be compiled without implicit contraction (gcc: -ffp-contract=off,
 * nvcc: --fmad=false); FMAs appear only where bm_fma() is written out.
 *
 * Only ELEMENTARY math lives here.  The step / evaluation / reward formulas are
 * written separately in the kernel and in the oracle, so the parity tests
 * compare two implementations and not one source compiled twice.
 *
 * Polynomials: Cephes single-precision minimax coefficients (sinf, cosf,
 * atanf, tanf; S. Moshier, public domain), arguments reduced by a three-term
 * Cody-Waite split of pi/2 evaluated with fma.
 */
#ifndef BMCTS_MATH_H_
#define BMCTS_MATH_H_

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define BM_HD __host__ __device__ __forceinline__
#else
#define BM_HD static inline
#endif

#define BM_PI_F 3.1415927410e+00f
#define BM_PIO2_HI 1.5707963705e+00f  /* 0x3fc90fdb */
#define BM_PIO2_MID -4.3711388287e-08f /* 0xb33bbd2e */
#define BM_PIO2_LO -1.7151245100e-15f  /* 0xa6f72ced */
#define BM_TWO_PI_HI 6.2831854820e+00f
#define BM_TWO_PI_MID -1.7484555315e-07f
#define BM_TWO_OVER_PI 6.3661974669e-01f
#define BM_INV_TWO_PI 1.5915493667e-01f
#define BM_PIO4 7.8539818525e-01f

/* single-rounding multiply-add on both sides */
BM_HD float bm_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return __builtin_fmaf(a, b, c);
#endif
}

/* IEEE-754 correctly rounded division and square root (independent of
 * -prec-div / -prec-sqrt / --use_fast_math on the device side) */
BM_HD float bm_div(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}

BM_HD float bm_sqrt(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}

/* min / max with the semantics of the PTX min.f32 / max.f32 instructions (one FMNMX on
 * sm_100a): a NaN operand yields the other operand, and -0 orders below +0.  The host
 * side spells the same rule out (x86 minss/maxss differ on NaN and signed zeros). */
BM_HD float bm_min(float a, float b) {
#if defined(__CUDA_ARCH__)
  return fminf(a, b);
#else
  if (a != a) return b;
  if (b != b) return a;
  if (a == b) return signbit(a) ? a : b;
  return (a < b) ? a : b;
#endif
}
BM_HD float bm_max(float a, float b) {
#if defined(__CUDA_ARCH__)
  return fmaxf(a, b);
#else
  if (a != a) return b;
  if (b != b) return a;
  if (a == b) return signbit(a) ? b : a;
  return (a > b) ? a : b;
#endif
}
BM_HD float bm_clamp(float x, float lo, float hi) { return bm_min(bm_max(x, lo), hi); }
BM_HD float bm_abs(float a) { return fabsf(a); }

/* wrap an angle to [-pi, pi) */
BM_HD float bm_norm_angle(float a) {
  float k = floorf((a + BM_PI_F) * BM_INV_TWO_PI);
  float r = bm_fma(-k, BM_TWO_PI_HI, a);
  r = bm_fma(-k, BM_TWO_PI_MID, r);
  return r;
}

/* sin and cos of x (|x| up to a few hundred radians), ~1-2 ulp */
BM_HD void bm_sincos(float x, float* s_out, float* c_out) {
  float kf = rintf(x * BM_TWO_OVER_PI);
  float r = bm_fma(-kf, BM_PIO2_HI, x);
  r = bm_fma(-kf, BM_PIO2_MID, r);
  r = bm_fma(-kf, BM_PIO2_LO, r);
  int q = ((int)kf) & 3;
  float z = r * r;
  /* sin on [-pi/4, pi/4] */
  float ps = bm_fma(-1.9515295891e-4f, z, 8.3321608736e-3f);
  ps = bm_fma(ps, z, -1.6666654611e-1f);
  float sr = bm_fma(ps * z, r, r);
  /* cos on [-pi/4, pi/4] */
  float pc = bm_fma(2.443315711809948e-5f, z, -1.388731625493765e-3f);
  pc = bm_fma(pc, z, 4.166664568298827e-2f);
  float cr = bm_fma(pc * z, z, bm_fma(-0.5f, z, 1.0f));
  float s, c;
  if (q & 1) {
    s = cr;
    c = -sr;
  } else {
    s = sr;
    c = cr;
  }
  if (q & 2) {
    s = -s;
    c = -c;
  }
  *s_out = s;
  *c_out = c;
}

/* tan(x) for |x| <= pi/4 (steering angles; delta_max is validated <= pi/4) */
BM_HD float bm_tan_small(float x) {
  float z = x * x;
  float p = bm_fma(9.38540185543e-3f, z, 3.11992232697e-3f);
  p = bm_fma(p, z, 2.44301354525e-2f);
  p = bm_fma(p, z, 5.34112807005e-2f);
  p = bm_fma(p, z, 1.33387994085e-1f);
  p = bm_fma(p, z, 3.33331568548e-1f);
  return bm_fma(p * z, x, x);
}

/* atan(x), all x, ~2 ulp */
BM_HD float bm_atan(float x) {
  float ax = bm_abs(x);
  float y0, t;
  if (ax > 2.414213562373095f) { /* tan(3pi/8) */
    y0 = BM_PIO2_HI;
    t = -bm_div(1.0f, ax);
  } else if (ax > 0.4142135623730950f) { /* tan(pi/8) */
    y0 = BM_PIO4;
    t = bm_div(ax - 1.0f, ax + 1.0f);
  } else {
    y0 = 0.0f;
    t = ax;
  }
  float z = t * t;
  float p = bm_fma(8.05374449538e-2f, z, -1.38776856032e-1f);
  p = bm_fma(p, z, 1.99777106478e-1f);
  p = bm_fma(p, z, -3.33329491539e-1f);
  float r = y0 + bm_fma(p * z, t, t);
  return (x < 0.0f) ? -r : r;
}

/* atan2(y, x) in (-pi, pi], ~2 ulp.  The octant reduction of bm_atan is applied to the pair
 * (|y|, |x|) directly, so there is exactly ONE division whatever the range (a reduction on the
 * rounded quotient needs a second one for steep angles, which diverges badly on the GPU):
 *   |y| <= tan(pi/8) |x|   : t =  |y| / |x|,              y0 = 0
 *   |y| <= tan(3pi/8) |x|  : t = (|y| - |x|) / (|y| + |x|), y0 = pi/4
 *   otherwise              : t = -|x| / |y|,              y0 = pi/2
 * A zero numerator (an agent exactly on its lane centre line, or |y| == |x|) skips the
 * division: the device's IEEE division would take its slow path for it. */
BM_HD float bm_atan2(float y, float x) {
  /* x == 0 needs no case of its own: it lands in the third range (num = -0, no division,
   * +-pi/2) or, with y == 0 as well, in the first (num = 0, result +0) */
  const float ay = bm_abs(y), ax = bm_abs(x);
  const int big = ay > 2.414213562373095f * ax;
  const int mid = ay > 0.4142135623730950f * ax;
  const float num = big ? -ax : (mid ? ay - ax : ay);
  const float den = big ? ay : (mid ? ay + ax : ax);
  const float y0 = big ? BM_PIO2_HI : (mid ? BM_PIO4 : 0.0f);
  const float t = (num == 0.0f) ? 0.0f : bm_div(num, den);
  const float z = t * t;
  float p = bm_fma(8.05374449538e-2f, z, -1.38776856032e-1f);
  p = bm_fma(p, z, 1.99777106478e-1f);
  p = bm_fma(p, z, -3.33329491539e-1f);
  const float r = y0 + bm_fma(p * z, t, t); /* atan(|y| / |x|) in [0, pi/2] */
  float a = ((y < 0.0f) != (x < 0.0f)) ? -r : r;
  if (x < 0.0f) a = (y >= 0.0f) ? (a + BM_PI_F) : (a - BM_PI_F);
  return a;
}

/* x^n for a small non-negative integer n (IDM "Exponent", default 4) */
BM_HD float bm_powi(float x, int n) {
  /* left-to-right products ((x*x)*x)*x ...: the unrolled cases round exactly like the loop */
  switch (n) {
    case 0: return 1.0f;
    case 1: return x;
    case 2: return x * x;
    case 3: return (x * x) * x;
    case 4: return ((x * x) * x) * x;
    default: break;
  }
  float p = ((x * x) * x) * x;
  for (int i = 4; i < n; ++i) p = p * x;
  return p;
}

/* uniform in [0, 1) from 32 random bits: 24-bit mantissa, exact conversion */
BM_HD float bm_u01(uint32_t bits) { return (float)(bits >> 8) * 5.9604644775390625e-8f; }

/* unbiased-enough integer in [0, n) by multiply-shift (n <= 2^16) */
BM_HD uint32_t bm_bounded(uint32_t bits, uint32_t n) {
  return (uint32_t)(((uint64_t)bits * (uint64_t)n) >> 32);
}

#endif /* BMCTS_MATH_H_ */
