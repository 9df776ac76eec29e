// cloudsc2_math.cuh -- branch-free FP64 elementary functions for the CLOUDSC2 kernels.
//
// Why: ncu on the first NL kernel (profiles/r1_nl_baseline.md) showed only 37 % of the executed
// warp instructions on the FP64 pipe; the rest was overhead of the general-purpose libdevice
// routines -- 64-bit immediates for polynomial coefficients re-materialised at every call (UMOV
// 12 %), special-case tests and slow-path branches of exp()/division (BRA/BSSY/BSYNC/FSETP/FSEL
// 14 %) -- which also fence the scheduler's view and leave the dependent DFMA chains exposed
// ("wait" stalls).  The functions below assume what the physics guarantees (finite arguments,
// denominators well inside the normal range) and are straight-line code:
//   csc2_rcp(x)   : MUFU.RCP64H seed + one third-order step, no denormal / inf fix-up branch
//   csc2_exp(x)   : Cody-Waite reduction to |r| <= ln2/32 with a 16-entry shared-memory table of
//                   2^(j/16) + degree-6 minimax polynomial
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

// exp(x) = 2^k * 2^(j/16) * exp(r), n = 16 k + j = nearest integer to 16 x / ln2, |r| <= ln2/32.
//  * 2^(j/16) comes from a 16-entry table in SHARED memory: 16 doubles cover the 32 banks exactly
//    once, so any pattern of per-lane indices is conflict-free (equal indices broadcast);
//  * exp(r) = 1 + r + r^2 P(r), P of degree 4 from a Remez exchange on the relative error of exp
//    (tools/gen_exp_coeffs.py 6 16): 1.2e-17;
//  * 11 FP64-pipe instructions instead of 16 for the table-free degree-11 form -- the kernels are
//    FP64-issue bound and spend a third of their FP64 work in exp.
// Polynomial coefficients live in the constant bank so that DFMA takes them as c[bank][off]
// operands (no 64-bit immediates / UMOV pairs).
static __constant__ double csc2_expc[5] = {
    4.99999999999996059e-01, 1.66666666644170708e-01, 4.16666666932605373e-02,
    8.33347194847556573e-03, 1.38885940074739067e-03};
static __constant__ double csc2_exptab_c[16] = {
    1.00000000000000000e+00, 1.04427378242741375e+00, 1.09050773266525769e+00,
    1.13878863475669156e+00, 1.18920711500272103e+00, 1.24185781207348400e+00,
    1.29683955465100964e+00, 1.35425554693689265e+00, 1.41421356237309515e+00,
    1.47682614593949935e+00, 1.54221082540794074e+00, 1.61049033194925428e+00,
    1.68179283050742900e+00, 1.75625216037329945e+00, 1.83400808640934243e+00,
    1.91520656139714740e+00};
__shared__ double csc2_exptab[16];

// Every kernel that evaluates csc2_exp calls this once, with ALL threads of the CTA, before any
// thread may exit.
__device__ __forceinline__ void csc2_math_init() {
  if (threadIdx.x < 16) csc2_exptab[threadIdx.x] = csc2_exptab_c[threadIdx.x];
  __syncthreads();
}

// 1/x: the MUFU.RCP64H seed works on the high word of x (relative error e ~ 2^-20 .. 2^-23); one
// third-order step y (1 + e + e^2) leaves e^3 < 2^-60, the final FMA rounds to <= 1 ulp.
__device__ __forceinline__ double csc2_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));   // MUFU.RCP64H
  double e = fma(-x, y, 1.0);
  e = fma(e, e, e);
  return fma(y, e, y);
}
__device__ __forceinline__ double csc2_div(double a, double b) { return a * csc2_rcp(b); }

// exp for arguments known to lie in [-700, 700]
__device__ __forceinline__ double csc2_exp(double x) {
  // n = nearest integer to x*16/ln2 via the 1.5*2^52 shift; its low word is n as int32
  const double shift = 6755399441055744.0;
  const double t = fma(x, 2.30831206542234142e+01, shift);
  const int n = __double2loint(t);
  const double kd = t - shift;
  const double tab = csc2_exptab[n & 15];
  // r = x - n*ln2/16 in two pieces (the high piece has 21 trailing zero bits: n*hi is exact)
  double r = fma(kd, -4.33216987730702385306e-02, x);
  r = fma(kd, -1.19263433079411731251e-11, r);
  const double r2 = r * r;
  const double pa = fma(csc2_expc[1], r, csc2_expc[0]);
  double pb = fma(csc2_expc[3], r, csc2_expc[2]);
  pb = fma(csc2_expc[4], r2, pb);
  const double q = fma(pb, r2, pa);
  const double p = fma(q, r2, r);               // exp(r) - 1
  const double v = fma(tab, p, tab);            // in [0.97, 1.96]
  // scale by 2^k: add k to the exponent field (|k| <= 1011 keeps the result normal)
  const int hi = __double2hiint(v) + ((n >> 4) << 20);
  return __hiloint2double(hi, __double2loint(v));
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

// Keep an expensive value unconditionally computed: without this the compiler turns
// `cond ? f(x) : c` back into a (divergent) branch around f, which splits the level's basic block
// and stops the scheduler from interleaving f with the surrounding chains.
__device__ __forceinline__ double csc2_pin(double x) {
  asm volatile("" : "+d"(x));
  return x;
}

// Comparisons of a double with a POSITIVE constant on the integer pipe: for c > 0 and any
// non-NaN x (negative, zero of either sign, positive) the IEEE order of x and c equals the order
// of their bit patterns read as signed 64-bit integers.  Moves the compare off the FP64 pipe
// (DSETP costs a 2-cycle FP64 slot), which is what bounds these kernels (tools/probes/fp64_probe).
__device__ __forceinline__ bool csc2_lt_pos(double x, double c) {   // x <  c, c > 0
  return __double_as_longlong(x) < __double_as_longlong(c);
}
__device__ __forceinline__ bool csc2_gt_pos(double x, double c) {   // x >  c, c > 0
  return __double_as_longlong(x) > __double_as_longlong(c);
}
__device__ __forceinline__ bool csc2_ge_pos(double x, double c) {   // x >= c, c > 0
  return __double_as_longlong(x) >= __double_as_longlong(c);
}
__device__ __forceinline__ double csc2_min_pos(double x, double c) { return csc2_lt_pos(x, c) ? x : c; }
__device__ __forceinline__ double csc2_max_pos(double x, double c) { return csc2_gt_pos(x, c) ? x : c; }
