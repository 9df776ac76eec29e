// cloudsc2_ad_kernel.cu -- the adjoint kernel (CLOUDSC2AD, reference
// src/cloudsc2_ad/cloudsc2ad.F90:10-1746) and the reduction of the adjoint (dot-product) test
// (cloudsc_driver_ad_mod.F90:184-267).
//
// One thread per column, two sweeps in two launches:
//   forward  JK = 1..KLEV : the NL kernel itself (cloudsc2_nl_kernel.cu, csc2_launch_nl_traj: fused
//            SATUR + nonlinear level at the NL kernel's occupancy), writing the trajectory outputs
//            like the reference (:842-864); the ONLY state the reverse sweep takes from it is the
//            rain/snow flux entering each level = the outputs PFPLSL5 / PFPLSN5 themselves (2 doubles
//            per level, instead of the reference's 116 stored (KLON,KLEV) arrays);
//   reverse  JK = KLEV..1 : k_cloudsc2_ad below -- reload the level's inputs, recompute its local
//            trajectory, run the adjoint statements (ad_level), and finish the level's 16 input
//            adjoints at once -- either accumulated into the caller's arrays (X = X + ..., PSUPSAT
//            assigned, :1733) or contracted on the fly with dx = dot_scale * x for the adjoint
//            test (ZNORM2).
// Output adjoints are consumed AND zeroed as the reference does (:917-919, :955-966, :1678-1690).
#include <cstdlib>

#include "cloudsc2_ad.cuh"
#include "cloudsc2_stage.cuh"
#include "cloudsc2_launch.h"

namespace {

__device__ __forceinline__ double ldin(const double *p) { return __ldg(p); }
__device__ __forceinline__ void stout(double *p, double v) { __stcs(p, v); }

constexpr int NT = CSC2_AD_THREADS;
// fields staged per level: 15 trajectory inputs, PQS (optional), 2 check-points, 9 output adjoints
constexpr int AD_NF = 27;
constexpr int AD_STAGES = 2;   // 3 stages measured slower (3.11 vs 2.97 ms): the larger carve-out costs L1

// input-adjoint accumulation X = X + dX as a fire-and-forget reduction at the L2 (RED.ADD.F64):
// no load latency on the thread's critical path, same DRAM traffic as a read-modify-write
__device__ __forceinline__ void acc(double *p, double v) { atomicAdd(p, v); }

// MINB = CTAs per SM: the straight-line adjoint level wants the full 255 registers (2 CTAs/SM,
// 3.36 ms); at 168 registers (3 CTAs/SM) it spills ~450 B/thread and takes 3.79 ms.
template <bool RV, bool DOT, bool LREG, int MINB>
__global__ void __launch_bounds__(CSC2_AD_THREADS, MINB)
k_cloudsc2_ad(const __grid_constant__ KConst c, const Geom g, const TrajIn in, const TrajOut out,
              const IncIn din, const IncOut dout, const ADOpts opt) {
  extern __shared__ double ring_all[];
  double *ring = ring_all + threadIdx.x;
  csc2_math_init();
  const int gcol = blockIdx.x * blockDim.x + threadIdx.x;
  const int ibl = gcol / g.nproma;
  if (ibl >= g.nblocks || gcol >= g.ngptot) return;
  const int jl = gcol - ibl * g.nproma;
  const int klev = g.klev, nproma = g.nproma;
  const ColOffsets o = csc2_col_offsets(ibl, jl, nproma, klev, in.bs_cld, in.bs_cml, out.bs_loc);
  // Flux entering level JK (the only state the reverse sweep needs from the forward sweep): it IS
  // PFPLSL5(JK) / PFPLSN5(JK) of the trajectory outputs (cloudsc2ad.F90:847-848), written by the forward
  // sweep of this call or by the CLOUDSC2 / CLOUDSC2TL call the caller vouches for (have_traj) -- no
  // separate check-point array is written or read.
  constexpr int SLOT = AD_NF * NT;
  // level pitch of the trajectory flux arrays: NPROMA in the blocked layout (flux_pitch = 0).  Kept a run-time
  // value on purpose: folded into the shared level offset jk*NPROMA, ptxas spills a loop-carried value of the
  // 255-register adjoint level (24 B, an STL/LDL pair in the level loop) and the sweep is 3.5 % slower
  // (2.147 vs 2.073 ms, tools/ad_ab.sh); with its own multiply the spill is the two loop invariants (16 B).
  const size_t cks = opt.flux_pitch > 0 ? (size_t)opt.flux_pitch : (size_t)nproma;

  const CritRH crh = make_critrh(tropopause_eta(c, in.pt, in.gt, o.o1, o.ocml, nproma));

  // ---------------------------------- reverse sweep ---------------------------------------------
  // per level: trajectory inputs + the flux check-point + the 9 output adjoints, all staged one
  // level ahead; output adjoints are zeroed only after their staged copy has landed.
  auto stage_rev = [&](double *d, int jk) {
    const size_t l = (size_t)jk * nproma;
    csc2_stage_traj<NT, true>(d, in, o, jk, klev, nproma);
    csc2_cp_async8(d + 16 * NT, out.pfplsl + o.oh + (size_t)jk * cks);
    csc2_cp_async8(d + 17 * NT, out.pfplsn + o.oh + (size_t)jk * cks);
    csc2_cp_async8(d + 18 * NT, dout.tent + o.o1 + l);
    csc2_cp_async8(d + 19 * NT, dout.tenq + o.o1 + l);
    csc2_cp_async8(d + 20 * NT, dout.tenl + o.o1 + l);
    csc2_cp_async8(d + 21 * NT, dout.teni + o.o1 + l);
    csc2_cp_async8(d + 22 * NT, dout.pclc + o.o1 + l);
    csc2_cp_async8(d + 23 * NT, dout.pfplsl + o.oh + l + nproma);
    csc2_cp_async8(d + 24 * NT, dout.pfplsn + o.oh + l + nproma);
    csc2_cp_async8(d + 25 * NT, dout.pfhpsl + o.oh + l + nproma);
    csc2_cp_async8(d + 26 * NT, dout.pfhpsn + o.oh + l + nproma);
  };
#pragma unroll
  for (int s = 0; s < AD_STAGES - 1; ++s) {
    if (klev - 1 - s >= 0) stage_rev(ring + s * SLOT, klev - 1 - s);
    csc2_cp_async_commit();
  }

  CarryAD ca;
  ca.rfl = 0.0;
  ca.sfl = 0.0;
  double paph_pending = 0.0;     // contribution of level JK+1 to PAPHP1(JK+1), not yet written
  double paph_hi5 = ldin(in.paph + o.oh + (size_t)klev * nproma);   // PAPHP15(JK+1)
  double dot = 0.0;
  const double ds = opt.dot_scale;
  const bool zero_sup = opt.zero_psupsat_pert != 0;
  int slot = 0, pslot = AD_STAGES - 1;
  for (int jk = klev - 1; jk >= 0; --jk) {
    const size_t l = (size_t)jk * nproma;
    const int pf = jk - (AD_STAGES - 1);
    if (pf >= 0) stage_rev(ring + pslot * SLOT, pf);
    csc2_cp_async_commit();
    csc2_cp_async_wait<AD_STAGES - 1>();
    const double *d = ring + slot * SLOT;
    LevIn x5 = csc2_read_level<NT>(d, jk, klev);   // field 0 is PAPHP15(JK) in the reverse sweep
    const double paph0 = x5.paph1;
    x5.paph1 = paph_hi5;
    const double pqs5 = in.pqs ? d[15 * NT] : satur_point(c, x5.pt, csc2_rcp(x5.pap));
    const double rfl5 = d[16 * NT];
    const double sfl5 = d[17 * NT];
    LevAdjIn ya;
    ya.tent = d[18 * NT];
    ya.tenq = d[19 * NT];
    ya.tenl = d[20 * NT];
    ya.teni = d[21 * NT];
    ya.pclc = d[22 * NT];
    // enthalpy-flux adjoints folded into the flux adjoints (:914-921)
    ya.fl = d[23 * NT] - d[25 * NT] * c.rlvtt;
    ya.fn = d[24 * NT] - d[26 * NT] * c.rlstt;
    // consumed and zeroed (:955-966, :1678-1690)
    dout.tent[o.o1 + l] = 0.0; dout.tenq[o.o1 + l] = 0.0; dout.tenl[o.o1 + l] = 0.0;
    dout.teni[o.o1 + l] = 0.0; dout.pclc[o.o1 + l] = 0.0; dout.pcovptot[o.o1 + l] = 0.0;
    dout.pfplsl[o.oh + l + nproma] = 0.0; dout.pfplsn[o.oh + l + nproma] = 0.0;
    dout.pfhpsl[o.oh + l + nproma] = 0.0; dout.pfhpsn[o.oh + l + nproma] = 0.0;

    LevAdj a;
    ad_level<RV, LREG>(c, crh, jk, x5, pqs5, paph0, rfl5, sfl5, ya, ca, a);

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
    paph_hi5 = paph0;
    slot = (slot + 1 == AD_STAGES) ? 0 : slot + 1;
    pslot = (pslot + 1 == AD_STAGES) ? 0 : pslot + 1;
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

template <bool RV, bool DOT, bool LREG, int MINB>
static cudaError_t launch_ad_k(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
                               const IncIn &din, const IncOut &dout, const ADOpts &opt, int grid,
                               cudaStream_t s) {
  const size_t smem = (size_t)AD_STAGES * AD_NF * NT * sizeof(double);
  auto kern = k_cloudsc2_ad<RV, DOT, LREG, MINB>;
  static CSC2_SMEM_FLAGS smem_ok_on_device{0};
  if (cudaError_t e0 = csc2_allow_smem(kern, smem, smem_ok_on_device)) return e0;
  // forward (trajectory) sweep: its own launch at the NL kernel's occupancy (12 warps/SM instead of
  // the 8 the adjoint level allows), check-pointing the fluxes the reverse sweep restarts from
  if (!opt.have_traj) {
    cudaError_t e = csc2_launch_nl_traj(c, g, in, out, s);
    if (e != cudaSuccess) return e;
  }
  kern<<<grid, CSC2_AD_THREADS, smem, s>>>(c, g, in, out, din, dout, opt);
  return cudaGetLastError();
}
template <bool RV, bool DOT>
static cudaError_t launch_ad_variant(const KConst &c, const Geom &g, const TrajIn &in,
                                     const TrajOut &out, const IncIn &din, const IncOut &dout,
                                     const ADOpts &opt, int grid, cudaStream_t s) {
#ifdef CSC2_EXPERIMENTS   // 3 CTAs/SM at 168 registers (CSC2_AD_MINB=3): slower, DESIGN.md 3.3
  static const int minb = [] { const char *e = getenv("CSC2_AD_MINB"); return e ? atoi(e) : 2; }();
  if (minb == 3) {
    if (c.lregcl) return launch_ad_k<RV, DOT, true, 3>(c, g, in, out, din, dout, opt, grid, s);
    return launch_ad_k<RV, DOT, false, 3>(c, g, in, out, din, dout, opt, grid, s);
  }
#endif
  if (c.lregcl) return launch_ad_k<RV, DOT, true, 2>(c, g, in, out, din, dout, opt, grid, s);
  return launch_ad_k<RV, DOT, false, 2>(c, g, in, out, din, dout, opt, grid, s);
}

cudaError_t csc2_launch_ad(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
                           const IncIn &din, const IncOut &dout, const ADOpts &opt, cudaStream_t s) {
  const long long ncol = (long long)g.nblocks * g.nproma;
  const int grid = (int)((ncol + CSC2_AD_THREADS - 1) / CSC2_AD_THREADS);
  const bool rv = c.rvtmp2 != 0.0;
  const bool dot = opt.dot_scale != 0.0;
  if (rv) {
    if (dot) return launch_ad_variant<true, true>(c, g, in, out, din, dout, opt, grid, s);
    return launch_ad_variant<true, false>(c, g, in, out, din, dout, opt, grid, s);
  }
  if (dot) return launch_ad_variant<false, true>(c, g, in, out, din, dout, opt, grid, s);
  return launch_ad_variant<false, false>(c, g, in, out, din, dout, opt, grid, s);
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

cudaError_t csc2_upload_levels_ad(const double *ceta, const double *zscalm, const double *sq1mceta,
                                  int klev, cudaStream_t s) {
  return csc2_upload_levels_impl(ceta, zscalm, sq1mceta, klev, s);
}
