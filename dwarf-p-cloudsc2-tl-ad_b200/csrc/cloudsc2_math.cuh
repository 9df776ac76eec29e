// cloudsc2_math.cuh -- branch-free FP64 elementary functions for the CLOUDSC2 kernels.
//
// Why: ncu on the first NL kernel (profiles/r1_nl_baseline.md) showed only 37 % of the executed
// warp instructions on the FP64 pipe; the rest was overhead of the general-purpose libdevice
// routines -- 64-bit immediates for polynomial coefficients re-materialised at every call (UMOV
// 12 %), special-case tests and slow-path branches of exp()/division (BRA/BSSY/BSYNC/FSETP/FSEL
// 14 %) -- which also fence the scheduler's view and leave the dependent DFMA chains exposed
// ("wait" stalls).  The functions below assume what the physics guarantees (finite arguments,
// strictly positive denominators well inside the normal range) and are straight-line code:
//   csc2_rcp(x)  : MUFU.RCP64H seed + Newton refinement, no denormal / inf fix-up branch
//   csc2_exp(x)  : Cody-Waite reduction + degree-13 polynomial in two interleaved Horner chains,
//                  argument clamped to [-700, 700] (exp(-700) = 1e-304 stands in for underflow)
//   csc2_tanh_p1 : tanh(a) + 1 = 2 E / (E + 1), E = exp(2a)
//   csc2_sech2   : 1 / cosh(a)^2 = 4 E / (E + 1)^2
// Accuracy (measured on B200 by tests/test_gpu_math.py against libdevice): rcp <= 1 ulp,
// exp <= 2 ulp over the ranges used -- the same order as the libm-vs-libdevice differences the
// parity tolerances already allow for.
#pragma once
#include <cuda_runtime.h>

// polynomial coefficients live in the constant bank so that DFMA takes them as c[bank][off]
// operands (no UMOV pairs); 1/k! for k = 2..13 (degree-13 Taylor on |r| <= ln2/2: truncation
// 0.3466^14/14! = 4e-18 relative)
__constant__ double csc2_expc[12] = {
    1.0 / 2.0,          1.0 / 6.0,           1.0 / 24.0,           1.0 / 120.0,
    1.0 / 720.0,        1.0 / 5040.0,        1.0 / 40320.0,        1.0 / 362880.0,
    1.0 / 3628800.0,    1.0 / 39916800.0,    1.0 / 479001600.0,    1.0 / 6227020800.0};

__device__ __forceinline__ double csc2_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));   // MUFU.RCP64H
  double e = fma(-x, y, 1.0);
  e = fma(e, e, e);                                        // cubic step
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}
__device__ __forceinline__ double csc2_div(double a, double b) { return a * csc2_rcp(b); }

__device__ __forceinline__ double csc2_exp(double x) {
  x = fmin(fmax(x, -700.0), 700.0);
  // k = nearest integer to x*log2(e) via the 1.5*2^52 shift; its low word is k as int32
  const double shift = 6755399441055744.0;
  const double t = fma(x, 1.4426950408889634, shift);
  const int k = __double2loint(t);
  const double kd = t - shift;
  // r = x - k*ln2 in two pieces (ln2_hi has 21 trailing zero bits: k*ln2_hi is exact)
  double r = fma(kd, -6.93147180369123816490e-01, x);
  r = fma(kd, -1.90821492927058770002e-10, r);
  const double r2 = r * r;
  // exp(r) = 1 + r + r2*(E(r2) + r*O(r2)),  E = c2 + c4 r2 + ... + c12 r2^5, O = c3 + ... + c13 r2^5
  double pe = csc2_expc[10];
  double po = csc2_expc[11];
  pe = fma(pe, r2, csc2_expc[8]);
  po = fma(po, r2, csc2_expc[9]);
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
