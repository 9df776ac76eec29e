// cloudsc2_common.cuh -- shared device-side definitions of the B200 CLOUDSC2 kernels.
//
// Design: one thread = one column; the NPROMA-contiguous blocked layout of the
// reference is kept, so a warp reads 32 consecutive JL of one level (one or two 128-B lines per
// field per level).  All physics constants travel as a by-value kernel parameter (constant
// bank), including the per-level CETA / ZSCALM vectors, so they cost no load instructions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cloudsc2_b200.h"
#include "cloudsc2_math.cuh"

#define CSC2_KLEV_MAX 256   // KLEV of the dwarf is 137

// Constants of one run, passed by value to every kernel (lives in the constant bank).
struct KConst {
  // YOMCST / YOETHF / YRECLDP / YREPHLI (cloudsc2.F90:104-111)
  double rg, rd, rcpd, retv, rlvtt, rlstt, rlmlt, rtt;
  double r2es, r3les, r3ies, r4les, r4ies, r5les, r5ies, r5alvcp, r5alscp;
  double ralvdcp, ralsdcp, rtwat, rtice, rtwat_rtice_r, rvtmp2;
  double rclcrit, rkconv, rlmin, rlptrc;
  // derived per run (cloudsc2.F90:235-244), computed once on the host
  double ptsphy, zckcodtl, zckcodti, zckcodtla, zckcodtia, zcons2, zcons3, zmeltp2, zqtmst;
  double rlcrit_inv;       // 1 / (2*RCLCRIT)   (ZLCRIT, cloudsc2.F90:508,525)
  double rcpd_inv;         // 1 / RCPD
  double rlmlt_inv;        // 1 / RLMLT
  double zcons2_inv;       // PTSPHY * RG = 1 / ZCONS2
  double zcor_cap;         // 1 / (1 - RETV*ZQMAX): ZCOR when ZESDP is capped (cloudsc2.F90:356,372)
  int lregcl;              // YRNCL%LREGCL
  int klev;
  int kwin0, kwin1;        // bounding range of levels with 0.1 < CETA < 0.4 (tropopause window)
};

// Per-level constants of one run, in __constant__ memory (one copy per translation unit, uploaded
// by csc2_upload_levels at cloudsc2_gpu_init):
//   [0] YRECLD%CETA   [1] ZSCALM = ZSCAL*MAX(CETA-0.2,ZEPS1)**0.2 (cloudsc2.F90:266)
//   [2] SQRT(MAX(1-CETA,0)), factor of the lowest ZCRH2 segment (:398)
// They were part of the by-value KConst kernel parameter at first; that made the parameter block
// 4.9 KB, and on this stack (sm_100a, driver 580, CUDA 12.9) parameters beyond byte 4096 were
// read WRONG by a few CTAs of large grids (found by tests/test_gpu_next_rows.py at 655 360
// columns: wrong SQRT(1-CETA) at levels >= 111).  All kernel parameter blocks are now < 1 KB
// (static_assert below).
static __constant__ double csc2_lev[3][CSC2_KLEV_MAX];
#define CSC2_CETA(jk) csc2_lev[0][jk]
#define CSC2_ZSCALM(jk) csc2_lev[1][jk]
#define CSC2_SQ1MCETA(jk) csc2_lev[2][jk]
static_assert(sizeof(KConst) <= 512, "keep the kernel parameter block small");

// host side, one instance per translation unit that contains kernels (see cloudsc2_launch.h)
static inline cudaError_t csc2_upload_levels_impl(const double *ceta, const double *zscalm,
                                                  const double *sq1mceta, int klev, cudaStream_t s) {
  cudaError_t e = cudaMemcpyToSymbolAsync(csc2_lev, ceta, klev * sizeof(double), 0, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess)
    e = cudaMemcpyToSymbolAsync(csc2_lev, zscalm, klev * sizeof(double), CSC2_KLEV_MAX * sizeof(double),
                                cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess)
    e = cudaMemcpyToSymbolAsync(csc2_lev, sq1mceta, klev * sizeof(double), 2 * CSC2_KLEV_MAX * sizeof(double),
                                cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  return e;
}

// Addressing of the blocked arrays: element (jl, jk, ibl) of a field with block stride bs is
// p[ibl*bs + jk*nproma + jl].  The same kernels serve the reference's host layout
// (PCLV: bs = 5*nproma*klev, B_CML/B_LOC: bs = 8*nproma*klev) and compact device layouts.
struct Geom {
  int nproma, klev, ngptot, nblocks;
};

// Trajectory inputs of CLOUDSC2 (cloudsc2.F90:10-18), device pointers + block strides.
struct TrajIn {
  const double *paph, *pap, *pq, *pt, *pl, *pi, *plude, *plu, *pmfu, *pmfd;
  const double *gt, *gq, *gl, *gi, *psupsat;
  const double *pqs;            // optional: PQS given (CLOUDSC2-call semantics); NULL = fused SATUR
  long long bs_cld, bs_cml;     // block strides of pl/pi and of gt/gq/gl/gi (doubles)
};
// Trajectory outputs.
struct TrajOut {
  double *tent, *tenq, *tenl, *teni, *pclc, *pfplsl, *pfplsn, *pfhpsl, *pfhpsn, *pcovptot;
  double *loc_last;             // TENDENCY_LOC%CLD(:,:,NCLV) zeroed by the driver (driver_mod.F90:88) or NULL
  long long bs_loc;             // block stride of tent/tenq/tenl/teni/loc_last
};
// The 16 + 10 increment arrays of CLOUDSC2TL / CLOUDSC2AD; all plain (NPROMA,KLEV[+1],NBLOCKS).
struct IncIn {
  double *paph, *pap, *pq, *pqs, *pt, *pl, *pi, *plude, *plu, *pmfu, *pmfd;
  double *gt, *gq, *gl, *gi, *psupsat;
};
struct IncOut {
  double *tent, *tenq, *tenl, *teni, *pclc, *pfplsl, *pfplsn, *pfhpsl, *pfhpsn, *pcovptot;
};

#define CSC2_ZQMAX 0.5
#define CSC2_ZEPS2 1.e-10

__device__ __forceinline__ double dmin_(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double dmax_(double a, double b) { return a > b ? a : b; }

// SATUR, LDPHYLIN branch (satur.F90:106-123) incl. FOEALFA (fcttre.func.h:73-75).  Straight-line:
// both Tetens exponentials are always evaluated (their weights may be 0) so that the compiler
// can interleave the two polynomial chains.
__device__ __forceinline__ double satur_point(const KConst &c, double t, double pap_inv) {
  const double x = (csc2_max_pos(csc2_min_pos(t, c.rtwat), c.rtice) - c.rtice) * c.rtwat_rtice_r;
  const double alfa = csc2_min_pos(x * x, 1.0);
  const double tm = t - c.rtt;
  const double el = csc2_exp(c.r3les * tm * csc2_rcp(t - c.r4les));
  const double ei = csc2_exp(c.r3ies * tm * csc2_rcp(t - c.r4ies));
  const double foeew = c.r2es * fma(alfa, el - ei, ei);      // ALFA*EL+(1-ALFA)*EI
  const double qs = csc2_min_pos(foeew * pap_inv, CSC2_ZQMAX);
  return qs * csc2_rcp(1.0 - c.retv * qs);
}

// Critical relative humidity profile (cloudsc2.F90:384-399).  zrh2 / zdeta1 depend only on the
// column's ZTRPAUS and are hoisted out of the level loop by the callers.
struct CritRH {
  double zeta3, zrh2, zdeta1, zrsq_deta1;
};
__device__ __forceinline__ CritRH make_critrh(double ztrpaus) {
  CritRH r;
  r.zeta3 = ztrpaus;
  double q = (ztrpaus - 0.25) / 0.15;
  r.zrh2 = 0.35 + 0.14 * (q * q) + 0.04 * dmin_(ztrpaus - 0.25, 0.0) / 0.15;
  r.zdeta1 = 0.09 + 0.16 * (0.4 - ztrpaus) / 0.3;
  r.zrsq_deta1 = 1.0 / sqrt(r.zdeta1);
  return r;
}
// sq1mceta = SQRT(1-CETA) of the level: SQRT((1-CETA)/ZDETA1) = SQRT(1-CETA) / SQRT(ZDETA1)
__device__ __forceinline__ double crit_rh(const CritRH &r, double ceta, double sq1mceta) {
  const double zdeta2 = 0.3;
  // the four segments of cloudsc2.F90:388-399, selected without branches
  const double lin = 1.0 + (r.zrh2 - 1.0) * ((ceta - r.zeta3) * (1.0 / zdeta2));
  const double low = 1.0 + (r.zrh2 - 1.0) * (sq1mceta * r.zrsq_deta1);
  double v = low;
  if (ceta < 1.0 - r.zdeta1) v = r.zrh2;
  if (ceta < r.zeta3 + zdeta2) v = lin;
  if (ceta < r.zeta3) v = 1.0;
  return v;
}

// Tropopause pre-pass (cloudsc2.F90:315-326): ZTRPAUS = CETA of the lowest level JK < KLEV in the
// window 0.1 < CETA < 0.4 whose first-guess T exceeds that of the level below; default 0.1.
__device__ __forceinline__ double tropopause_eta(const KConst &c, const double *__restrict__ pt,
                                                 const double *__restrict__ gt, size_t o_pt,
                                                 size_t o_gt, int nproma) {
  double ztrpaus = 0.1;
  if (c.kwin1 < c.kwin0) return ztrpaus;
  double t_hi = pt[o_pt + (size_t)c.kwin0 * nproma] + c.ptsphy * gt[o_gt + (size_t)c.kwin0 * nproma];
  for (int jk = c.kwin0; jk <= c.kwin1; ++jk) {       // kwin1 <= klev-2
    double t_lo = pt[o_pt + (size_t)(jk + 1) * nproma] + c.ptsphy * gt[o_gt + (size_t)(jk + 1) * nproma];
    double e = CSC2_CETA(jk);
    if (e > 0.1 && e < 0.4 && t_hi > t_lo) ztrpaus = e;
    t_hi = t_lo;
  }
  return ztrpaus;
}
