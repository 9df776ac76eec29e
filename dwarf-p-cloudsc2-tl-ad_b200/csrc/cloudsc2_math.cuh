// cloudsc2_math.cuh -- branch-free FP64 elementary functions for the CLOUDSC2 kernels.
//
// Why: ncu on the first NL kernel (profiles/r1_nl_baseline.md) showed only 37 % of the executed
// warp instructions on the FP64 pipe; the rest was overhead of the general-purpose libdevice
// routines -- 64-bit immediates for polynomial coefficients re-materialised at every call (UMOV
// 12 %), special-case tests and slow-path branches of exp()/division (BRA/BSSY/BSYNC/FSETP/FSEL
// 14 %) -- which also fence the scheduler's view and leave the dependent DFMA chains exposed
// ("wait" stalls).  The functions below assume what the physics guarantees (finite arguments,
// denominators well inside the normal range) and are straight-line code:
//   csc2_rcp(x)   : MUFU.RCP64H seed + Newton refinement, no denormal / inf fix-up branch
//   csc2_exp(x)   : Cody-Waite reduction + degree-11 minimax polynomial in two interleaved chains
//   csc2_expn(x)  : same with the argument clamped below at -700 (exp(-700) = 1e-304 stands in
//                   for underflow; used where the argument is -(something unbounded))
//   csc2_sqrt(x)  : MUFU.RSQ64H seed + coupled Newton (Goldschmidt) iteration, x = 0 -> 0
//   csc2_tanh_p1  : tanh(a) + 1 = 2 E / (E + 1), E = exp(2a)
//   csc2_tanh_p1_sech2 : the same plus 1 / cosh(a)^2 = 4 E / (E + 1)^2 (TL / AD)
// Accuracy (tests/test_gpu_math.py, against numpy on the ranges used): rcp <= 1 ulp, exp <= 2 ulp,
// sqrt <= 1 ulp -- the same order as the libm-vs-libdevice differences the parity tolerances
// already allow for.
#pragma once
#include <cuda_runtime.h>

// polynomial coefficients live in the constant bank so that DFMA takes them as c[bank][off]
// operands (no UMOV pairs).  exp(r) = 1 + r + r^2 P(r) on |r| <= ln2/2, P of degree 9 from a
// Remez exchange on the relative error of exp (tools/gen_exp_coeffs.py 11): 1.1e-17 with the
// coefficients rounded to double.
static __constant__ double csc2_expc[10] = {
    5.00000000000001110e-01, 1.66666666666664132e-01, 4.16666666665301555e-02,
    8.33333333349446127e-03, 1.38888889436301140e-03, 1.98412695065786879e-04,
    2.48014930989065504e-05, 2.75575863738401650e-06, 2.76302483792114415e-07,
    2.50000616028356664e-08};

// 1/x: the MUFU.RCP64H seed works on the high word of x (relative error ~2^-20 .. 2^-23); two
// quadratic Newton steps take it below 2^-70, the last FMA rounds to <= 1 ulp.
__device__ __forceinline__ double csc2_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));   // MUFU.RCP64H
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}
__device__ __forceinline__ double csc2_div(double a, double b) { return a * csc2_rcp(b); }

// exp for arguments known to lie in [-700, 700]
__device__ __forceinline__ double csc2_exp(double x) {
  // k = nearest integer to x*log2(e) via the 1.5*2^52 shift; its low word is k as int32
  const double shift = 6755399441055744.0;
  const double t = fma(x, 1.4426950408889634, shift);
  const int k = __double2loint(t);
  const double kd = t - shift;
  // r = x - k*ln2 in two pieces (ln2_hi has 21 trailing zero bits: k*ln2_hi is exact)
  double r = fma(kd, -6.93147180369123816490e-01, x);
  r = fma(kd, -1.90821492927058770002e-10, r);
  const double r2 = r * r;
  // P(r) = E(r2) + r O(r2) in two interleaved Horner chains
  double pe = csc2_expc[8];
  double po = csc2_expc[9];
  pe = fma(pe, r2, csc2_expc[6]);
  po = fma(po, r2, csc2_expc[7]);
  pe = fma(pe, r2, csc2_expc[4]);
  po = fma(po, r2, csc2_expc[5]);
  pe = fma(pe, r2, csc2_expc[2]);
  po = fma(po, r2, csc2_expc[3]);
  pe = fma(pe, r2, csc2_expc[0]);
  po = fma(po, r2, csc2_expc[1]);
  const double q = fma(po, r, pe);
  const double p = fma(q, r2, r) + 1.0;
  // scale by 2^k: add k to the exponent field (|k| <= 1010 keeps the result normal)
  const int hi = __double2hiint(p) + (k << 20);
  return __hiloint2double(hi, __double2loint(p));
}
// exp for arguments <= 0 of unbounded magnitude
__device__ __forceinline__ double csc2_expn(double x) { return csc2_exp(fmax(x, -700.0)); }

// sqrt(x), x >= 0 (normal or zero): one coupled Goldschmidt step on the MUFU.RSQ64H seed and a
// final residual correction
__device__ __forceinline__ double csc2_sqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));  // MUFU.RSQ64H
  double g = x * y;            // ~ sqrt(x)
  double h = 0.5 * y;          // ~ 1 / (2 sqrt(x))
  const double r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  const double d = fma(-g, g, x);
  g = fma(d, h, g);
  return x > 0.0 ? g : 0.0;
}

// tanh(a) + 1, and sech(a)^2, from one exponential
__device__ __forceinline__ double csc2_tanh_p1(double a) {
  const double e = csc2_exp(2.0 * a);
  return 2.0 * e * csc2_rcp(e + 1.0);
}
__device__ __forceinline__ void csc2_tanh_p1_sech2(double a, double &tanh_p1, double &sech2) {
  const double e = csc2_exp(2.0 * a);
  const double r = csc2_rcp(e + 1.0);
  const double er = e * r;
  tanh_p1 = 2.0 * er;
  sech2 = 4.0 * er * r;
}
