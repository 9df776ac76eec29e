// cloudsc2_ad_kernel.cu -- the adjoint kernel (CLOUDSC2AD, reference
// src/cloudsc2_ad/cloudsc2ad.F90:10-1746) and the reduction of the adjoint (dot-product) test
// (cloudsc_driver_ad_mod.F90:184-267).
//
// One thread per column, two sweeps in one launch:
//   forward  JK = 1..KLEV : fused SATUR + nonlinear level (identical device function to the NL
//            kernel), writing the trajectory outputs like the reference (:842-864) and
//            check-pointing ONLY the rain/snow flux entering each level (2 doubles per level,
//            instead of the reference's 116 stored (KLON,KLEV) arrays);
//   reverse  JK = KLEV..1 : reload the level's inputs, recompute its local trajectory, run the
//            adjoint statements (ad_level), and finish the level's 16 input adjoints at once --
//            either accumulated into the caller's arrays (X = X + ..., PSUPSAT assigned, :1733)
//            or contracted on the fly with dx = dot_scale * x for the adjoint test (ZNORM2).
// Output adjoints are consumed AND zeroed as the reference does (:917-919, :955-966, :1678-1690).
#include "cloudsc2_ad.cuh"
#include "cloudsc2_launch.h"

namespace {

__device__ __forceinline__ double ldin(const double *p) { return __ldg(p); }
__device__ __forceinline__ void stout(double *p, double v) { __stcs(p, v); }

struct ColOffsets {
  size_t o1, oh, ocld, ocml, oloc;
};

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

// read-and-zero of an output adjoint (plain loads: the arrays are read-write in this kernel)
__device__ __forceinline__ double take(double *p) {
  const double v = *p;
  *p = 0.0;
  return v;
}
__device__ __forceinline__ void acc(double *p, double v) { *p += v; }

template <bool RV, bool DOT>
__global__ void __launch_bounds__(CSC2_AD_THREADS)
k_cloudsc2_ad(const __grid_constant__ KConst c, const Geom g, const TrajIn in, const TrajOut out,
              const IncIn din, const IncOut dout, const ADOpts opt) {
  csc2_math_init();
  const int gcol = blockIdx.x * blockDim.x + threadIdx.x;
  const int ibl = gcol / g.nproma;
  if (ibl >= g.nblocks || gcol >= g.ngptot) return;
  const int jl = gcol - ibl * g.nproma;
  const int klev = g.klev, nproma = g.nproma;
  const size_t n2 = (size_t)nproma * klev;
  ColOffsets o;
  o.o1 = (size_t)ibl * n2 + jl;
  o.oh = (size_t)ibl * (n2 + nproma) + jl;
  o.ocld = (size_t)ibl * in.bs_cld + jl;
  o.ocml = (size_t)ibl * in.bs_cml + jl;
  o.oloc = (size_t)ibl * out.bs_loc + jl;
  double *ck_r = opt.ckpt + gcol;                                   // [klev][ncol_pad] rain
  double *ck_s = opt.ckpt + (size_t)klev * opt.ncol_pad + gcol;     // [klev][ncol_pad] snow
  const size_t cks = (size_t)opt.ncol_pad;

  const CritRH crh = make_critrh(tropopause_eta(c, in.pt, in.gt, o.o1, o.ocml, nproma));

  // ------------------------------ forward (trajectory) sweep --------------------------------
  {
    Carry st;
    st.paph0 = ldin(in.paph + o.oh);
    st.rfl = 0.0;
    st.sfl = 0.0;
    if (opt.write_traj) {
      stout(out.pfplsl + o.oh, 0.0);
      stout(out.pfplsn + o.oh, 0.0);
      stout(out.pfhpsl + o.oh, -0.0 * c.rlvtt);
      stout(out.pfhpsn + o.oh, -0.0 * c.rlstt);
    }
    LevIn cur = load_level(in, o, 0, klev, nproma);
    for (int jk = 0; jk < klev; ++jk) {
      LevIn nxt = cur;
      if (jk + 1 < klev) nxt = load_level(in, o, jk + 1, klev, nproma);
      ck_r[(size_t)jk * cks] = st.rfl;
      ck_s[(size_t)jk * cks] = st.sfl;
      const double pqs = in.pqs ? ldin(in.pqs + o.o1 + (size_t)jk * nproma)
                                : satur_point(c, cur.pt, 1.0 / cur.pap);
      LevOut y;
      nl_level(c, crh, jk, cur, pqs, st, y);
      if (opt.write_traj) {
        const size_t l = (size_t)jk * nproma;
        stout(out.tent + o.oloc + l, y.tent);
        stout(out.tenq + o.oloc + l, y.tenq);
        stout(out.tenl + o.oloc + l, y.tenl);
        stout(out.teni + o.oloc + l, y.teni);
        stout(out.pclc + o.o1 + l, y.pclc);
        stout(out.pcovptot + o.o1 + l, 0.0);
        stout(out.pfplsl + o.oh + l + nproma, y.rfln);
        stout(out.pfplsn + o.oh + l + nproma, y.sfln);
        stout(out.pfhpsl + o.oh + l + nproma, -y.rfln * c.rlvtt);
        stout(out.pfhpsn + o.oh + l + nproma, -y.sfln * c.rlstt);
      }
      cur = nxt;
    }
  }

  // ---------------------------------- reverse sweep ---------------------------------------------
  CarryAD ca;
  ca.rfl = 0.0;
  ca.sfl = 0.0;
  double paph_pending = 0.0;     // contribution of level JK+1 to PAPHP1(JK+1), not yet written
  double dot = 0.0;
  const double ds = opt.dot_scale;
  const bool zero_sup = opt.zero_psupsat_pert != 0;
  for (int jk = klev - 1; jk >= 0; --jk) {
    const size_t l = (size_t)jk * nproma;
    const LevIn x5 = load_level(in, o, jk, klev, nproma);
    const double paph0 = ldin(in.paph + o.oh + l);
    const double pqs5 = in.pqs ? ldin(in.pqs + o.o1 + l) : satur_point(c, x5.pt, 1.0 / x5.pap);
    const double rfl5 = ck_r[(size_t)jk * cks];
    const double sfl5 = ck_s[(size_t)jk * cks];
    LevAdjIn ya;
    ya.tent = take(dout.tent + o.o1 + l);
    ya.tenq = take(dout.tenq + o.o1 + l);
    ya.tenl = take(dout.tenl + o.o1 + l);
    ya.teni = take(dout.teni + o.o1 + l);
    ya.pclc = take(dout.pclc + o.o1 + l);
    dout.pcovptot[o.o1 + l] = 0.0;                                    // :1688-1690
    // enthalpy-flux adjoints folded into the flux adjoints (:914-921)
    ya.fl = take(dout.pfplsl + o.oh + l + nproma) - take(dout.pfhpsl + o.oh + l + nproma) * c.rlvtt;
    ya.fn = take(dout.pfplsn + o.oh + l + nproma) - take(dout.pfhpsn + o.oh + l + nproma) * c.rlstt;

    LevAdj a;
    ad_level<RV>(c, crh, jk, x5, pqs5, paph0, rfl5, sfl5, ya, ca, a);

    const double paph_hi = a.paph_hi + paph_pending;                  // total for PAPHP1(JK+1)
    paph_pending = a.paph_lo;
    if (DOT) {
      // <dx, dx*> with dx = ds * x (cloudsc_driver_ad_mod.F90:124-157, 240-256); ZSUPSAT0 = 0
      double s = x5.paph1 * paph_hi + x5.pap * a.pap + x5.pq * a.pq + pqs5 * a.pqs + x5.pt * a.pt +
                 x5.pl * a.pl + x5.pi * a.pi + x5.plude * a.plude + x5.plu1 * a.plu1 +
                 x5.pmfu * a.pmfu + x5.pmfd * a.pmfd + x5.gt * a.gt + x5.gq * a.gq +
                 x5.gl * a.gl + x5.gi * a.gi;
      if (!zero_sup) s += x5.psupsat * a.psupsat;
      dot += s;
      if (jk == 0) dot += paph0 * paph_pending;
    } else {
      acc(din.paph + o.oh + l + nproma, paph_hi);
      acc(din.pap + o.o1 + l, a.pap);
      acc(din.pq + o.o1 + l, a.pq);
      acc(din.pqs + o.o1 + l, a.pqs);
      acc(din.pt + o.o1 + l, a.pt);
      acc(din.pl + o.o1 + l, a.pl);
      acc(din.pi + o.o1 + l, a.pi);
      acc(din.plude + o.o1 + l, a.plude);
      if (jk < klev - 1) acc(din.plu + o.o1 + l + nproma, a.plu1);
      acc(din.pmfu + o.o1 + l, a.pmfu);
      acc(din.pmfd + o.o1 + l, a.pmfd);
      acc(din.gt + o.o1 + l, a.gt);
      acc(din.gq + o.o1 + l, a.gq);
      acc(din.gl + o.o1 + l, a.gl);
      acc(din.gi + o.o1 + l, a.gi);
      din.psupsat[o.o1 + l] = a.psupsat;                              // assignment, :1733
      if (jk == 0) acc(din.paph + o.oh, paph_pending);
    }
  }
  // top flux rows: consumed and zeroed (:917-919, :1677-1679)
  dout.pfplsl[o.oh] = 0.0;
  dout.pfplsn[o.oh] = 0.0;
  dout.pfhpsl[o.oh] = 0.0;
  dout.pfhpsn[o.oh] = 0.0;
  if (DOT) opt.coldot[gcol] = ds * dot;
}

__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

// ZNORM3 per column and its maximum (cloudsc_driver_ad_mod.F90:260-267): warp-shuffle max, one
// atomic per warp.
__global__ void k_ad_finalize(const Geom g, const double *__restrict__ n1, const double *__restrict__ n2,
                              double *__restrict__ norms_col, double *__restrict__ znormg) {
  const int gcol = blockIdx.x * blockDim.x + threadIdx.x;
  double n3 = 0.0;
  if (gcol < g.ngptot) {
    const double a = n1[gcol], b = n2[gcol];
    const double eps = 2.220446049250313e-16;                         // EPSILON(1._JPRB)
    n3 = (b == 0.0) ? fabs(a - b) / eps : fabs(a - b) / eps / b;
    norms_col[3 * (size_t)gcol] = a;
    norms_col[3 * (size_t)gcol + 1] = b;
    norms_col[3 * (size_t)gcol + 2] = n3;
    if (!(n3 == n3)) n3 = 1.0e300;                                    // NaN must fail the test
  }
  for (int off = 16; off > 0; off >>= 1) n3 = fmax(n3, __shfl_xor_sync(0xffffffffu, n3, off));
  if ((threadIdx.x & 31) == 0 && n3 > 0.0) atomic_max_nonneg(znormg, n3);
}

}  // namespace

cudaError_t csc2_launch_ad(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
                           const IncIn &din, const IncOut &dout, const ADOpts &opt, cudaStream_t s) {
  const long long ncol = (long long)g.nblocks * g.nproma;
  const int grid = (int)((ncol + CSC2_AD_THREADS - 1) / CSC2_AD_THREADS);
  const bool rv = c.rvtmp2 != 0.0;
  const bool dot = opt.dot_scale != 0.0;
  if (rv) {
    if (dot) k_cloudsc2_ad<true, true><<<grid, CSC2_AD_THREADS, 0, s>>>(c, g, in, out, din, dout, opt);
    else k_cloudsc2_ad<true, false><<<grid, CSC2_AD_THREADS, 0, s>>>(c, g, in, out, din, dout, opt);
  } else {
    if (dot) k_cloudsc2_ad<false, true><<<grid, CSC2_AD_THREADS, 0, s>>>(c, g, in, out, din, dout, opt);
    else k_cloudsc2_ad<false, false><<<grid, CSC2_AD_THREADS, 0, s>>>(c, g, in, out, din, dout, opt);
  }
  return cudaGetLastError();
}

cudaError_t csc2_launch_ad_finalize(const Geom &g, const double *n1, const double *n2,
                                    double *norms_col, double *znormg, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(znormg, 0, 16 * sizeof(double), s);
  if (e != cudaSuccess) return e;
  const long long ncol = (long long)g.nblocks * g.nproma;
  const int grid = (int)((ncol + 127) / 128);
  k_ad_finalize<<<grid, 128, 0, s>>>(g, n1, n2, norms_col, znormg);
  return cudaGetLastError();
}
