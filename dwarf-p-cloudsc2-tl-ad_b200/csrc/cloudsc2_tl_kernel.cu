// cloudsc2_tl_kernel.cu -- tangent-linear kernel and the kernels of the Taylor test.
//  k_cloudsc2_tl      : SATUR + CLOUDSC2TL for every column (cloudsc2tl.F90), increments either read
//                       from arrays or generated on load as pert_scale * x (the drivers' 1 % rule,
//                       cloudsc_driver_tl_mod.F90:156-171), optional in-kernel column sums / norms.
//  k_taylor_nl        : the 10 perturbed nonlinear sweeps of the Taylor test (:197-230), one grid.y
//                       slice per lambda, perturbation applied on load, emitting only
//                       sum_levels(F - F5) per column and field.
//  k_taylor_finalize  : ERROR_NORM (:21-31) per block and lambda, max over blocks (:247-252).
#include <cstdlib>

#include "cloudsc2_tl.cuh"
#include "cloudsc2_stage.cuh"
#include "cloudsc2_launch.h"

namespace {

__device__ __forceinline__ double ldin(const double *p) { return __ldg(p); }
__device__ __forceinline__ void stout(double *p, double v) { __stcs(p, v); }

constexpr int NT = CSC2_TL_THREADS;
// fields staged per level: 15 trajectory inputs, PQS (optional), 16 increments
constexpr int TL_NF = 32;

// increments from arrays (all plain (NPROMA,KLEV[+1],NBLOCKS)) -> fields 16..31
template <int NT>
__device__ __forceinline__ void stage_incr(double *d, const IncIn &di, const ColOffsets &o, int jk,
                                           int klev, int nproma) {
  const size_t l = (size_t)jk * nproma;
  csc2_cp_async8(d + 16 * NT, di.paph + o.oh + l + nproma);
  csc2_cp_async8(d + 17 * NT, di.pap + o.o1 + l);
  csc2_cp_async8(d + 18 * NT, di.pt + o.o1 + l);
  csc2_cp_async8(d + 19 * NT, di.pq + o.o1 + l);
  csc2_cp_async8(d + 20 * NT, di.pl + o.o1 + l);
  csc2_cp_async8(d + 21 * NT, di.pi + o.o1 + l);
  csc2_cp_async8(d + 22 * NT, di.plude + o.o1 + l);
  if (jk < klev - 1) csc2_cp_async8(d + 23 * NT, di.plu + o.o1 + l + nproma);
  csc2_cp_async8(d + 24 * NT, di.pmfu + o.o1 + l);
  csc2_cp_async8(d + 25 * NT, di.pmfd + o.o1 + l);
  csc2_cp_async8(d + 26 * NT, di.gt + o.o1 + l);
  csc2_cp_async8(d + 27 * NT, di.gq + o.o1 + l);
  csc2_cp_async8(d + 28 * NT, di.gl + o.o1 + l);
  csc2_cp_async8(d + 29 * NT, di.gi + o.o1 + l);
  csc2_cp_async8(d + 30 * NT, di.psupsat + o.o1 + l);
  csc2_cp_async8(d + 31 * NT, di.pqs + o.o1 + l);
}
// direct (unstaged) loads of one level, used by the Taylor-test kernel
__device__ __forceinline__ LevIn load_level(const TrajIn &in, const ColOffsets &o, int jk, int klev,
                                            int nproma) {
  LevIn x;
  const size_t l = (size_t)jk * nproma;
  x.paph1 = ldin(in.paph + o.oh + l + nproma);
  x.pap = ldin(in.pap + o.o1 + l);
  x.pt = ldin(in.pt + o.o1 + l);
  x.pq = ldin(in.pq + o.o1 + l);
  x.pl = ldin(in.pl + o.ocld + l);
  x.pi = ldin(in.pi + o.ocld + l);
  x.plude = ldin(in.plude + o.o1 + l);
  x.plu1 = (jk < klev - 1) ? ldin(in.plu + o.o1 + l + nproma) : 0.0;
  x.pmfu = ldin(in.pmfu + o.o1 + l);
  x.pmfd = ldin(in.pmfd + o.o1 + l);
  x.gt = ldin(in.gt + o.ocml + l);
  x.gq = ldin(in.gq + o.ocml + l);
  x.gl = ldin(in.gl + o.ocml + l);
  x.gi = ldin(in.gi + o.ocml + l);
  x.psupsat = ldin(in.psupsat + o.o1 + l);
  return x;
}

__device__ __forceinline__ LevIn scale_level(const LevIn &x, double f, bool zero_sup) {
  LevIn d;
  d.paph1 = x.paph1 * f; d.pap = x.pap * f; d.pt = x.pt * f; d.pq = x.pq * f; d.pl = x.pl * f;
  d.pi = x.pi * f; d.plude = x.plude * f; d.plu1 = x.plu1 * f; d.pmfu = x.pmfu * f;
  d.pmfd = x.pmfd * f; d.gt = x.gt * f; d.gq = x.gq * f; d.gl = x.gl * f; d.gi = x.gi * f;
  d.psupsat = zero_sup ? 0.0 : x.psupsat * f;
  return d;
}

// NT threads per CTA, MINB CTAs per SM -> register cap 65536 / (MINB * NT): 128 x 2 = 8 warps/SM at 255
// registers is the default; 64 x 5 = 10 warps at 200 registers and 64 x 6 = 12 warps at 168 are tuning
// variants.
template <bool ONFLY, int STAGES, bool RV, bool LREG, int MINB, int NT>
__global__ void __launch_bounds__(NT) __maxnreg__((65536 / (MINB * NT)) > 255 ? 255 : (65536 / (MINB * NT)) / 8 * 8)
k_cloudsc2_tl(const __grid_constant__ KConst c, const Geom g, const TrajIn in, const TrajOut out,
              const IncIn din, const IncOut dout, const TLOpts opt) {
  extern __shared__ double ring_all[];
  double *ring = ring_all + threadIdx.x;
  csc2_math_init();
  const int gcol = blockIdx.x * blockDim.x + threadIdx.x;
  const int ibl = gcol / g.nproma;
  if (ibl >= g.nblocks || gcol >= g.ngptot) return;
  const int jl = gcol - ibl * g.nproma;
  const int klev = g.klev, nproma = g.nproma;
  const ColOffsets o = csc2_col_offsets(ibl, jl, nproma, klev, in.bs_cld, in.bs_cml, out.bs_loc);
  const bool zero_sup = opt.zero_psupsat_pert != 0;
  const bool wr = dout.tent != nullptr;
  constexpr int SLOT = TL_NF * NT;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < klev) {
      csc2_stage_traj<NT, false>(ring + s * SLOT, in, o, s, klev, nproma);
      if (!ONFLY) stage_incr<NT>(ring + s * SLOT, din, o, s, klev, nproma);
    }
    csc2_cp_async_commit();
  }

  // ZTRPAUS from the trajectory only (cloudsc2tl.F90:431-442)
  const CritRH crh = make_critrh(tropopause_eta(c, in.pt, in.gt, o.o1, o.ocml, nproma));

  Carry st5; CarryTL st;
  st5.paph0 = ldin(in.paph + o.oh); st5.rfl = 0.0; st5.sfl = 0.0;
  st.paph0 = ONFLY ? st5.paph0 * opt.pert_scale : ldin(din.paph + o.oh);
  st.rfl = 0.0; st.sfl = 0.0;
  // top rows (cloudsc2tl.F90:419-422, :1108-1115)
  stout(out.pfplsl + o.oh, 0.0); stout(out.pfplsn + o.oh, 0.0);
  stout(out.pfhpsl + o.oh, -0.0 * c.rlvtt); stout(out.pfhpsn + o.oh, -0.0 * c.rlstt);
  if (wr) {
    stout(dout.pfplsl + o.oh, 0.0); stout(dout.pfplsn + o.oh, 0.0);
    stout(dout.pfhpsl + o.oh, -0.0 * c.rlvtt); stout(dout.pfhpsn + o.oh, -0.0 * c.rlstt);
  }
  double s_t = 0, s_q = 0, s_l = 0, s_i = 0, s_c = 0, s_fl = 0, s_fn = 0, s_hl = 0, s_hn = 0;
  double q_t = 0, q_q = 0, q_l = 0, q_i = 0, q_c = 0, q_fl = 0, q_fn = 0, q_hl = 0, q_hn = 0;

  int slot = 0, pslot = STAGES - 1;
  for (int jk = 0; jk < klev; ++jk) {
    const int pf = jk + STAGES - 1;
    if (pf < klev) {
      csc2_stage_traj<NT, false>(ring + pslot * SLOT, in, o, pf, klev, nproma);
      if (!ONFLY) stage_incr<NT>(ring + pslot * SLOT, din, o, pf, klev, nproma);
    }
    csc2_cp_async_commit();
    csc2_cp_async_wait<STAGES - 1>();
    const double *d = ring + slot * SLOT;
    const LevIn cur = csc2_read_level<NT>(d, jk, klev);
    const double pqs5 = in.pqs ? d[15 * NT] : satur_point(c, cur.pt, csc2_rcp(cur.pap));
    LevIn dx; double dpqs;
    if (ONFLY) {
      dx = scale_level(cur, opt.pert_scale, zero_sup);
      dpqs = pqs5 * opt.pert_scale;
    } else {
      dx = csc2_read_level<NT>(d + 16 * NT, jk, klev);
      dpqs = d[31 * NT];
    }
    LevOut y5, dy;
    tl_level<RV, LREG>(c, crh, jk, cur, pqs5, dx, dpqs, st5, st, y5, dy);

    const size_t l = (size_t)jk * nproma;
    // trajectory outputs, re-emitted like the reference (cloudsc2tl.F90:1079-1091)
    stout(out.tent + o.oloc + l, y5.tent); stout(out.tenq + o.oloc + l, y5.tenq);
    stout(out.tenl + o.oloc + l, y5.tenl); stout(out.teni + o.oloc + l, y5.teni);
    stout(out.pclc + o.o1 + l, y5.pclc);   stout(out.pcovptot + o.o1 + l, 0.0);
    stout(out.pfplsl + o.oh + l + nproma, y5.rfln); stout(out.pfplsn + o.oh + l + nproma, y5.sfln);
    stout(out.pfhpsl + o.oh + l + nproma, -y5.rfln * c.rlvtt);
    stout(out.pfhpsn + o.oh + l + nproma, -y5.sfln * c.rlstt);
    const double hl = -dy.rfln * c.rlvtt, hn = -dy.sfln * c.rlstt;
    if (wr) {
      stout(dout.tent + o.o1 + l, dy.tent); stout(dout.tenq + o.o1 + l, dy.tenq);
      stout(dout.tenl + o.o1 + l, dy.tenl); stout(dout.teni + o.o1 + l, dy.teni);
      stout(dout.pclc + o.o1 + l, dy.pclc); stout(dout.pcovptot + o.o1 + l, 0.0);
      stout(dout.pfplsl + o.oh + l + nproma, dy.rfln); stout(dout.pfplsn + o.oh + l + nproma, dy.sfln);
      stout(dout.pfhpsl + o.oh + l + nproma, hl); stout(dout.pfhpsn + o.oh + l + nproma, hn);
    }
    s_t += dy.tent; s_q += dy.tenq; s_l += dy.tenl; s_i += dy.teni; s_c += dy.pclc;
    s_fl += dy.rfln; s_fn += dy.sfln; s_hl += hl; s_hn += hn;
    q_t += dy.tent * dy.tent; q_q += dy.tenq * dy.tenq; q_l += dy.tenl * dy.tenl;
    q_i += dy.teni * dy.teni; q_c += dy.pclc * dy.pclc; q_fl += dy.rfln * dy.rfln;
    q_fn += dy.sfln * dy.sfln; q_hl += hl * hl; q_hn += hn * hn;
    slot = (slot + 1 == STAGES) ? 0 : slot + 1;
    pslot = (pslot + 1 == STAGES) ? 0 : pslot + 1;
  }
  if (opt.colsum) {
    double *s = opt.colsum + gcol;
    const long long n = opt.ncol_pad;
    s[0] = s_t; s[n] = s_q; s[2 * n] = s_l; s[3 * n] = s_i; s[4 * n] = s_c; s[5 * n] = s_fl;
    s[6 * n] = s_fn; s[7 * n] = s_hl; s[8 * n] = s_hn; s[9 * n] = 0.0;   // PCOVPTOT' == 0
  }
  if (opt.colsq)   // ZNORM1, summed in the reference's field order (cloudsc_driver_ad_mod.F90:184-195)
    opt.colsq[gcol] = q_t + q_q + q_l + q_i + q_c + q_fl + q_fn + q_hl + q_hn + 0.0;
}

struct Lambdas {
  double v[10];
};

// x5 = x + lambda * (x * 0.01)   (cloudsc_driver_tl_mod.F90:156-171, 200-215)
__device__ __forceinline__ double pert(double x, double lam) { return x + lam * (x * 0.01); }

// Ring fields of the Taylor kernel: the 15 (+1) trajectory inputs, then the nine baseline outputs
// F5 the differences are formed against.
constexpr int TY_NF = CSC2_NTRAJ + 9;
__device__ __forceinline__ void stage_base(double *d, const TrajOut &b, const ColOffsets &o, int jk,
                                           int nproma) {
  const size_t l = (size_t)jk * nproma;
  csc2_cp_async8(d + 16 * NT, b.tent + o.oloc + l);
  csc2_cp_async8(d + 17 * NT, b.tenq + o.oloc + l);
  csc2_cp_async8(d + 18 * NT, b.tenl + o.oloc + l);
  csc2_cp_async8(d + 19 * NT, b.teni + o.oloc + l);
  csc2_cp_async8(d + 20 * NT, b.pclc + o.o1 + l);
  csc2_cp_async8(d + 21 * NT, b.pfplsl + o.oh + l + nproma);
  csc2_cp_async8(d + 22 * NT, b.pfplsn + o.oh + l + nproma);
  csc2_cp_async8(d + 23 * NT, b.pfhpsl + o.oh + l + nproma);
  csc2_cp_async8(d + 24 * NT, b.pfhpsn + o.oh + l + nproma);
}

// CTA order: lambda is the FASTEST index (blockIdx.x = column_cta * 10 + ilam), so the ten sweeps over
// the same 128 columns are resident together and nine of them find the inputs and the baseline
// outputs in L2 -- DRAM sees them once, not ten times.  Levels are staged one ahead in the same
// shared-memory ring as the NL kernel.
// (168 registers = 3 CTAs/SM: at 128 the kernel spills and the ten sweeps take 8.7 instead of 8.3 ms)
template <bool HAS_PQS, bool RV>
__global__ void __maxnreg__(168)
k_taylor_nl(const __grid_constant__ KConst c, const Geom g, const TrajIn in, const TrajOut base,
            const __grid_constant__ Lambdas lams, double *__restrict__ diffsum, const long long ncol_pad) {
  extern __shared__ double ring_all[];
  double *ring = ring_all + threadIdx.x;
  csc2_math_init();
  const int cta = blockIdx.x / 10;
  const int ilam = blockIdx.x - cta * 10;
  const int gcol = cta * NT + threadIdx.x;
  const int ibl = gcol / g.nproma;
  if (ibl >= g.nblocks || gcol >= g.ngptot) return;
  const int jl = gcol - ibl * g.nproma;
  const int klev = g.klev, nproma = g.nproma;
  const double lam = lams.v[ilam];
  const ColOffsets o = csc2_col_offsets(ibl, jl, nproma, klev, in.bs_cld, in.bs_cml, base.bs_loc);
  constexpr int SLOT = TY_NF * NT;

  csc2_stage_traj<NT, false, HAS_PQS ? 1 : 0>(ring, in, o, 0, klev, nproma);
  stage_base(ring, base, o, 0, nproma);
  csc2_cp_async_commit();

  // tropopause level of the PERTURBED state (the perturbed run is a plain CLOUDSC2 call)
  double ztrpaus = 0.1;
  if (c.kwin1 >= c.kwin0) {
    auto tfg = [&](int jk) {
      const size_t l = (size_t)jk * nproma;
      return pert(ldin(in.pt + o.o1 + l), lam) + c.ptsphy * pert(ldin(in.gt + o.ocml + l), lam);
    };
    double t_hi = tfg(c.kwin0);
    for (int jk = c.kwin0; jk <= c.kwin1; ++jk) {
      const double t_lo = tfg(jk + 1);
      const double e = CSC2_CETA(jk);
      if (e > 0.1 && e < 0.4 && t_hi > t_lo) ztrpaus = e;
      t_hi = t_lo;
    }
  }
  const CritRH crh = make_critrh(ztrpaus);

  Carry st;
  st.paph0 = pert(ldin(in.paph + o.oh), lam);
  st.rfl = 0.0; st.sfl = 0.0;
  double d_t = 0, d_q = 0, d_l = 0, d_i = 0, d_c = 0, d_fl = 0, d_fn = 0, d_hl = 0, d_hn = 0;
  int slot = 0;
  for (int jk = 0; jk < klev; ++jk) {
    if (jk + 1 < klev) {
      csc2_stage_traj<NT, false, HAS_PQS ? 1 : 0>(ring + (slot ^ 1) * SLOT, in, o, jk + 1, klev, nproma);
      stage_base(ring + (slot ^ 1) * SLOT, base, o, jk + 1, nproma);
    }
    csc2_cp_async_commit();
    csc2_cp_async_wait<1>();
    const double *r = ring + slot * SLOT;
    LevIn x = csc2_read_level<NT>(r, jk, klev);
    // PQS5 = ZQSAT + lambda*(0.01*ZQSAT) with ZQSAT = SATUR of the UNPERTURBED state (:135,:204)
    const double qs = HAS_PQS ? r[15 * NT] : satur_point(c, x.pt, csc2_rcp(x.pap));
    const double pqs5 = pert(qs, lam);
    x.paph1 = pert(x.paph1, lam); x.pap = pert(x.pap, lam); x.pt = pert(x.pt, lam);
    x.pq = pert(x.pq, lam); x.pl = pert(x.pl, lam); x.pi = pert(x.pi, lam);
    x.plude = pert(x.plude, lam); x.plu1 = pert(x.plu1, lam); x.pmfu = pert(x.pmfu, lam);
    x.pmfd = pert(x.pmfd, lam); x.gt = pert(x.gt, lam); x.gq = pert(x.gq, lam);
    x.gl = pert(x.gl, lam); x.gi = pert(x.gi, lam); x.psupsat = pert(x.psupsat, lam);
    LevOut y;
    nl_level<RV>(c, crh, jk, x, pqs5, st, y);
    d_t += r[16 * NT] - y.tent;
    d_q += r[17 * NT] - y.tenq;
    d_l += r[18 * NT] - y.tenl;
    d_i += r[19 * NT] - y.teni;
    d_c += r[20 * NT] - y.pclc;
    d_fl += r[21 * NT] - y.rfln;
    d_fn += r[22 * NT] - y.sfln;
    d_hl += r[23 * NT] - (-y.rfln * c.rlvtt);
    d_hn += r[24 * NT] - (-y.sfln * c.rlstt);
    slot ^= 1;
  }
  double *d = diffsum + (size_t)ilam * 10 * ncol_pad + gcol;
  const long long n = ncol_pad;
  d[0] = d_t; d[n] = d_q; d[2 * n] = d_l; d[3 * n] = d_i; d[4 * n] = d_c; d[5 * n] = d_fl;
  d[6 * n] = d_fn; d[7 * n] = d_hl; d[8 * n] = d_hn; d[9 * n] = 0.0;
}

__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v) {
  // valid for non-negative doubles: the IEEE bit pattern is monotone in the value
  atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

// One WARP per (block, lambda): lanes stride over the block's columns, the two sums of ERROR_NORM
// (SUM(PTL), SUM(PNL-PNL5), cloudsc_driver_tl_mod.F90:21-31) are reduced with warp shuffles in a
// fixed order (deterministic), lane 0 folds the ten fields and publishes the block's ratio.
__global__ void __launch_bounds__(128)
k_taylor_finalize(const Geom g, const Lambdas lams, const double *__restrict__ tlsum,
                  const double *__restrict__ diffsum, const long long ncol_pad,
                  double *__restrict__ ratios_blk, double *__restrict__ znormg,
                  int *__restrict__ degenerate) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= g.nblocks * 10) return;               // whole warps leave together
  const int ibl = w / 10, ilam = w - ibl * 10;
  const double lam = lams.v[ilam];
  const int c0 = ibl * g.nproma;
  int icend = g.ngptot - c0;
  if (icend > g.nproma) icend = g.nproma;
  double znorm = 0.0, zcount = 0.0;
  for (int f = 0; f < 10; ++f) {   // T,Q,L,I,PA,PFPLSL,PFPLSN,PFHPSL,PFHPSN,PCOVPTOT (:235-244)
    double s_tl = 0.0, s_d = 0.0;
    const double *a = tlsum + (size_t)f * ncol_pad + c0;
    const double *b = diffsum + ((size_t)ilam * 10 + f) * ncol_pad + c0;
    for (int jl = lane; jl < icend; jl += 32) { s_tl += a[jl]; s_d += b[jl]; }
    for (int off = 16; off > 0; off >>= 1) {
      s_tl += __shfl_xor_sync(0xffffffffu, s_tl, off);
      s_d += __shfl_xor_sync(0xffffffffu, s_d, off);
    }
    const double den = s_tl * lam;
    if (fabs(den) > 2.220446049250313e-16) {   // EPSILON(ZLAMBDA)
      zcount += 1.0;
      znorm += fabs(s_d / den);
    }
  }
  if (lane != 0) return;
  if (znorm == 0.0 || zcount == 0.0) {
    atomicAdd(degenerate, 1);
    ratios_blk[(size_t)ibl * 10 + ilam] = nan("");
  } else {
    const double r = znorm / zcount;
    ratios_blk[(size_t)ibl * 10 + ilam] = r;
    atomic_max_nonneg(&znormg[ilam], r);
  }
}

}  // namespace

template <bool ONFLY, int STAGES, bool RV, bool LREG, int MINB = 2, int TNT = CSC2_TL_THREADS>
static cudaError_t launch_tl_k(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
                               const IncIn &din, const IncOut &dout, const TLOpts &opt, int /*grid*/,
                               cudaStream_t s) {
  const long long ncol = (long long)g.nblocks * g.nproma;
  const int grid = (int)((ncol + TNT - 1) / TNT);
  const size_t smem = (size_t)STAGES * TL_NF * TNT * sizeof(double);
  auto kern = k_cloudsc2_tl<ONFLY, STAGES, RV, LREG, MINB, TNT>;
  static CSC2_SMEM_FLAGS smem_ok_on_device{0};
  if (cudaError_t e0 = csc2_allow_smem(kern, smem, smem_ok_on_device)) return e0;
  kern<<<grid, TNT, smem, s>>>(c, g, in, out, din, dout, opt);
  return cudaGetLastError();
}
// dispatch on the two run-time switches that are compile-time in the kernel
template <bool ONFLY, int STAGES>
static cudaError_t launch_tl_variant(const KConst &c, const Geom &g, const TrajIn &in,
                                     const TrajOut &out, const IncIn &din, const IncOut &dout,
                                     const TLOpts &opt, int grid, cudaStream_t s) {
  const bool rv = c.rvtmp2 != 0.0, lreg = c.lregcl != 0;
  if (rv) {
    // RVTMP2 != 0 never happens in this dwarf: one shape only
    if (lreg) return launch_tl_k<ONFLY, 2, true, true>(c, g, in, out, din, dout, opt, grid, s);
    return launch_tl_k<ONFLY, 2, true, false>(c, g, in, out, din, dout, opt, grid, s);
  }
#ifdef CSC2_EXPERIMENTS   // occupancy variants measured in DESIGN.md 3.3 (CSC2_TL_MINB); the product has one shape
  static const int minb = [] { const char *e = getenv("CSC2_TL_MINB"); return e ? atoi(e) : 2; }();
  if (minb == 93)  // 96-thread CTAs, 3 per SM = 9 warps at 224 registers (no spills), 147 kB of ring per SM
    return lreg ? launch_tl_k<ONFLY, STAGES, false, true, 3, 96>(c, g, in, out, din, dout, opt, grid, s)
                : launch_tl_k<ONFLY, STAGES, false, false, 3, 96>(c, g, in, out, din, dout, opt, grid, s);
  if (minb == 9)   // 32-thread CTAs, 9 per SM = 9 warps at 224 registers
    return lreg ? launch_tl_k<ONFLY, STAGES, false, true, 9, 32>(c, g, in, out, din, dout, opt, grid, s)
                : launch_tl_k<ONFLY, STAGES, false, false, 9, 32>(c, g, in, out, din, dout, opt, grid, s);
  if (minb == 8)   // 32-thread CTAs, 8 per SM = 8 warps at 255 registers (finer tail)
    return lreg ? launch_tl_k<ONFLY, STAGES, false, true, 8, 32>(c, g, in, out, din, dout, opt, grid, s)
                : launch_tl_k<ONFLY, STAGES, false, false, 8, 32>(c, g, in, out, din, dout, opt, grid, s);
  if (minb == 5 || minb == 6 || minb == 4) {   // tuning knobs: 64-thread CTAs, 4/5/6 per SM = 8/10/12 warps
    if (minb == 4) return lreg ? launch_tl_k<ONFLY, STAGES, false, true, 4, 64>(c, g, in, out, din, dout, opt, grid, s)
                               : launch_tl_k<ONFLY, STAGES, false, false, 4, 64>(c, g, in, out, din, dout, opt, grid, s);
    if (minb == 5) return lreg ? launch_tl_k<ONFLY, STAGES, false, true, 5, 64>(c, g, in, out, din, dout, opt, grid, s)
                               : launch_tl_k<ONFLY, STAGES, false, false, 5, 64>(c, g, in, out, din, dout, opt, grid, s);
    return lreg ? launch_tl_k<ONFLY, STAGES, false, true, 6, 64>(c, g, in, out, din, dout, opt, grid, s)
                : launch_tl_k<ONFLY, STAGES, false, false, 6, 64>(c, g, in, out, din, dout, opt, grid, s);
  }
  if (minb == 3) {   // tuning knob: 3 CTAs/SM at 168 registers
    if (lreg) return launch_tl_k<ONFLY, STAGES, false, true, 3>(c, g, in, out, din, dout, opt, grid, s);
    return launch_tl_k<ONFLY, STAGES, false, false, 3>(c, g, in, out, din, dout, opt, grid, s);
  }
#endif
  if (lreg) return launch_tl_k<ONFLY, STAGES, false, true>(c, g, in, out, din, dout, opt, grid, s);
  return launch_tl_k<ONFLY, STAGES, false, false>(c, g, in, out, din, dout, opt, grid, s);
}


cudaError_t csc2_launch_tl(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
                           const IncIn &din, const IncOut &dout, const TLOpts &opt, cudaStream_t s) {
  const long long ncol = (long long)g.nblocks * g.nproma;
  const int grid = (int)((ncol + CSC2_TL_THREADS - 1) / CSC2_TL_THREADS);
  if (opt.pert_scale != 0.0) return launch_tl_variant<true, 2>(c, g, in, out, din, dout, opt, grid, s);
#ifdef CSC2_EXPERIMENTS
  static const int stages = [] { const char *e = getenv("CSC2_TL_STAGES"); return e && atoi(e) == 3 ? 3 : 2; }();
  if (stages == 3) return launch_tl_variant<false, 3>(c, g, in, out, din, dout, opt, grid, s);
#endif
  return launch_tl_variant<false, 2>(c, g, in, out, din, dout, opt, grid, s);
}

static Lambdas make_lambdas() {
  Lambdas l;
  for (int i = 1; i <= 10; ++i) l.v[i - 1] = pow(10.0, -(double)i);   // ZLAMBDA=10**(-ILAM) (:198)
  return l;
}

template <bool HAS_PQS, bool RV>
static cudaError_t launch_taylor_nl_k(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &base,
                                      double *diffsum, long long ncol_pad, cudaStream_t s) {
  const long long ncol = (long long)g.nblocks * g.nproma;
  const unsigned ncta = (unsigned)((ncol + NT - 1) / NT);
  const size_t smem = (size_t)2 * TY_NF * NT * sizeof(double);
  auto kern = k_taylor_nl<HAS_PQS, RV>;
  static CSC2_SMEM_FLAGS smem_ok_on_device{0};
  if (cudaError_t e0 = csc2_allow_smem(kern, smem, smem_ok_on_device)) return e0;
  kern<<<ncta * 10u, NT, smem, s>>>(c, g, in, base, make_lambdas(), diffsum, ncol_pad);
  return cudaGetLastError();
}

cudaError_t csc2_launch_taylor_nl(const KConst &c, const Geom &g, const TrajIn &in,
                                  const TrajOut &base, double *diffsum, long long ncol_pad,
                                  cudaStream_t s) {
  if (c.rvtmp2 != 0.0)
    return in.pqs ? launch_taylor_nl_k<true, true>(c, g, in, base, diffsum, ncol_pad, s)
                  : launch_taylor_nl_k<false, true>(c, g, in, base, diffsum, ncol_pad, s);
  return in.pqs ? launch_taylor_nl_k<true, false>(c, g, in, base, diffsum, ncol_pad, s)
                : launch_taylor_nl_k<false, false>(c, g, in, base, diffsum, ncol_pad, s);
}

cudaError_t csc2_launch_taylor_finalize(const Geom &g, const double *tlsum, const double *diffsum,
                                        long long ncol_pad, double *ratios_blk, double *znormg,
                                        int *degenerate, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(znormg, 0, 16 * sizeof(double), s);   // znormg[10] + flag
  if (e != cudaSuccess) return e;
  const long long n = (long long)g.nblocks * 10 * 32;      // one warp per (block, lambda)
  k_taylor_finalize<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g, make_lambdas(), tlsum, diffsum, ncol_pad,
                                                    ratios_blk, znormg, degenerate);
  return cudaGetLastError();
}

cudaError_t csc2_upload_levels_tl(const double *ceta, const double *zscalm, const double *sq1mceta,
                                  int klev, cudaStream_t s) {
  return csc2_upload_levels_impl(ceta, zscalm, sq1mceta, klev, s);
}
