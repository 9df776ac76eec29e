// cloudsc2_launch.h -- host-callable launchers of the CUDA kernels (one per .cu file).
#pragma once
#include "cloudsc2_common.cuh"

#define CSC2_NL_THREADS 128
#define CSC2_TL_THREADS 128
#define CSC2_AD_THREADS 128

// Fused SATUR + CLOUDSC2 over all blocks (cloudsc2_nl_kernel.cu).
cudaError_t csc2_launch_nl(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
                           cudaStream_t s);
// Device-side cyclic expansion (cloudsc2_nl_kernel.cu).
cudaError_t csc2_launch_expand(const double *src, int nlon, long long rows, double *dst, int nproma,
                               int ngptot, int nblocks, long long gcol0, cudaStream_t s);

// Tangent linear (cloudsc2_tl_kernel.cu).
//  pert_scale != 0 : increments are generated on load as pert_scale * trajectory input
//                    (the drivers' dx = 0.01 x); `din` is then ignored.
//  dout            : TL output fields (may hold NULL pointers when only sums are wanted)
//  colsum          : optional [10][ncol_pad] per-column sums over levels of the 10 TL outputs
//  colsq           : optional [ncol_pad] per-column sum of squares of the 10 TL outputs (ZNORM1)
struct TLOpts {
  double pert_scale;
  int zero_psupsat_pert;   // AD driver: ZSUPSAT = 0 (cloudsc_driver_ad_mod.F90:139)
  double *colsum;
  double *colsq;
  long long ncol_pad;
};
cudaError_t csc2_launch_tl(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
                           const IncIn &din, const IncOut &dout, const TLOpts &opt, cudaStream_t s);

// Perturbed nonlinear runs of the Taylor test (cloudsc2_tl_kernel.cu): for ilam = 1..10 (grid.y)
// run CLOUDSC2 on x + lambda*(0.01 x) with PQS = qsat(x) + lambda*(0.01 qsat(x)) and accumulate
// per column sum_levels(F - F5) for the 10 outputs into diffsum[ilam][field][col].
cudaError_t csc2_launch_taylor_nl(const KConst &c, const Geom &g, const TrajIn &in,
                                  const TrajOut &base, double *diffsum, long long ncol_pad,
                                  cudaStream_t s);
// ERROR_NORM + max over blocks (cloudsc_driver_tl_mod.F90:21-31, 233-252).
cudaError_t csc2_launch_taylor_finalize(const Geom &g, const double *tlsum, const double *diffsum,
                                        long long ncol_pad, double *ratios_blk, double *znormg,
                                        int *degenerate, cudaStream_t s);

// Adjoint (cloudsc2_ad_kernel.cu): forward sweep = the NL kernel writing the trajectory outputs (its
// PFPLSL / PFPLSN are the flux check-points), reverse sweep recomputing each level's trajectory.
//  dot_scale != 0 : input adjoints are not written; instead <dot_scale * x, x*> is accumulated
//                   per column into coldot (ZNORM2, cloudsc_driver_ad_mod.F90:240-256).
struct ADOpts {
  double dot_scale;
  int zero_psupsat_pert;
  double *coldot;
  long long ncol_pad;      // pitch of coldot
  int have_traj;           // the trajectory fluxes PFPLSL5/PFPLSN5 in `out` are already those of `in`
                           // (a CLOUDSC2 / CLOUDSC2TL call on the same inputs ran before, as in the
                           // adjoint test and in any 4D-Var inner loop): skip the forward sweep
  long long flux_pitch;    // level pitch (doubles) of out.pfplsl / out.pfplsn as read by the reverse sweep; 0 = NPROMA
};
cudaError_t csc2_launch_ad(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
                           const IncIn &din, const IncOut &dout, const ADOpts &opt, cudaStream_t s);
// ZNORM3 per column and max (cloudsc_driver_ad_mod.F90:260-267).
cudaError_t csc2_launch_ad_finalize(const Geom &g, const double *n1, const double *n2,
                                    double *norms_col, double *znormg, cudaStream_t s);

// Accuracy probe of cloudsc2_math.cuh: y[i] = fn(x[i]); fn 0 rcp, 1 exp, 2 expn, 3 sqrt,
// 4 tanh+1, 5 sech^2 (device pointers).
cudaError_t csc2_launch_math_probe(int fn, const double *x, double *y, int n, cudaStream_t s);
// NL launch configuration (see csc2_launch_nl); also settable with CSC2_NL_VARIANT.
void csc2_set_nl_variant(int v);

// Device-side validation statistics (cloudsc2_validate_kernel.cu): out5 = min(field), max(field),
// max|err|, sum|err|, sum|ref| against the un-expanded reference columns ref_src (nlon, rows).
#define CSC2_VALIDATE_MAX_CTAS (148 * 8)
size_t csc2_validate_scratch_bytes();
cudaError_t csc2_launch_validate(const double *ref_src, int nlon, const double *field, int nproma,
                                 long long rows, long long blk_stride, int ngptot, int nblocks,
                                 long long gcol0, void *scratch, double *out5, cudaStream_t s);

cudaError_t csc2_launch_validate_split(const double *ref_src, int nlon, const double *field, int nproma,
                                       long long rows, long long blk_stride, int ngptot, int nblocks,
                                       long long gcol0, void *scratch, double *o_min, double *o_max2,
                                       double *o_sum2, cudaStream_t s);

// Cross-rank reductions of the test norms (cloudsc2_validate_kernel.cu): see k_norms_prepare / k_fill.
cudaError_t csc2_launch_norms_prepare(double *z, int n, const int *deg, double *deg_as_double, cudaStream_t s);
cudaError_t csc2_launch_fill(double *z, int n, double v, cudaStream_t s);

// SATUR alone, elementwise over n points (device pointers).
cudaError_t csc2_launch_satur(const KConst &c, const double *pap, const double *pt, double *pqsat,
                              long long n, cudaStream_t s);

// Upload CETA / ZSCALM / SQRT(1-CETA) (klev values each, host pointers) into the __constant__
// level table of each kernel translation unit.  Called by cloudsc2_gpu_init.
cudaError_t csc2_upload_levels_nl(const double *ceta, const double *zscalm, const double *sq1mceta,
                                  int klev, cudaStream_t s);
cudaError_t csc2_upload_levels_tl(const double *ceta, const double *zscalm, const double *sq1mceta,
                                  int klev, cudaStream_t s);
cudaError_t csc2_upload_levels_ad(const double *ceta, const double *zscalm, const double *sq1mceta,
                                  int klev, cudaStream_t s);

// Forward (trajectory) sweep of the adjoint as its own launch (cloudsc2_nl_kernel.cu): the NL kernel without the
// driver-level zeroing; its PFPLSL / PFPLSN outputs are the flux check-points of the reverse sweep.  Columns
// beyond ngptot are not touched.
cudaError_t csc2_launch_nl_traj(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
                                cudaStream_t s);
