// cloudsc2_nl_kernel.cu -- the nonlinear kernel: fused SATUR + CLOUDSC2 for every column of every
// NPROMA block in one launch.  Replaces the OpenMP block loop of CLOUDSC_DRIVER
// (reference src/cloudsc2_nl/cloudsc_driver_mod.F90:82-111).
//
// Mapping: thread <-> column, CTA = 128 consecutive columns, the KLEV loop keeps the column state
// in registers while the next level's 15 inputs are copied by cp.async into a per-thread slot of a
// shared-memory ring (cloudsc2_stage.cuh).  Loads/stores are coalesced along NPROMA (JL); every
// input is read exactly once (plus the ~40-level tropopause pre-pass over PT/PGTENT), PQSAT never
// touches memory.  The same kernel, with flux check-points, is the forward sweep of the adjoint.
#include <cstdint>
#include <cstdlib>

#include "cloudsc2_nl.cuh"
#include "cloudsc2_stage.cuh"
#include "cloudsc2_launch.h"

namespace {

__device__ __forceinline__ double ldin(const double *p) { return __ldg(p); }
__device__ __forceinline__ void stout(double *p, double v) { __stcs(p, v); }

constexpr int NL_NF = CSC2_NTRAJ;   // staged fields per level: 15 inputs + optional PQS

// Template parameters: HAS_PQS (PQS supplied by the caller instead of the fused SATUR), STAGES =
// depth of the shared-memory ring (levels in flight + the one being computed), NT = threads per
// CTA, MAXREG = register cap (-> CTAs per SM), RV = (RVTMP2 != 0).
// The same kernel is the forward (trajectory) sweep of the adjoint: the rain / snow flux entering level
// JK+1 is its output PFPLSL / PFPLSN(JK+1), which the reverse sweep restarts from -- no check-point array.

// PROBE != 0 exists only in the experiments build (tools/probes): extra dummy instructions per level.
#ifdef CSC2_EXPERIMENTS
#include "experiments/cloudsc2_nl_probe.cuh"
#else
template <int PROBE>
struct NlProbe {
  __device__ __forceinline__ void level(int) {}
  __device__ __forceinline__ bool fired() const { return false; }
};
#endif
template <bool HAS_PQS, int STAGES, int NT, int MAXREG, bool RV, int PROBE = 0>
__global__ void __maxnreg__(MAXREG)
k_cloudsc2_nl(const __grid_constant__ KConst c, const Geom g, const TrajIn in, const TrajOut out) {
  extern __shared__ double ring_all[];
  double *ring = ring_all + threadIdx.x;
  csc2_math_init();
  const int gcol = blockIdx.x * blockDim.x + threadIdx.x;
  const int ibl = gcol / g.nproma;
  if (ibl >= g.nblocks) return;
  const int jl = gcol - ibl * g.nproma;
  const int klev = g.klev, nproma = g.nproma;
  const ColOffsets o = csc2_col_offsets(ibl, jl, nproma, klev, in.bs_cld, in.bs_cml, out.bs_loc);

  if (gcol >= g.ngptot) {
    // padding column of the last block: the DRIVER zeroes PCOVPTOT and TENDENCY_LOC%CLD(:,:,NCLV)
    // of whole blocks (driver_mod.F90:87-88); CLOUDSC2 itself never touches columns beyond ICEND.
    // loc_last != NULL marks the driver-level call (NULL: plain CLOUDSC2 semantics, e.g. cloudsc2_).
    if (out.loc_last)
      for (int jk = 0; jk < klev; ++jk) {
        stout(out.pcovptot + o.o1 + (size_t)jk * nproma, 0.0);
        stout(out.loc_last + o.oloc + (size_t)jk * nproma, 0.0);
      }
    return;
  }

  // start the pipeline before the tropopause pre-pass so that its latency is covered
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < klev) csc2_stage_traj<NT, false, HAS_PQS ? 1 : 0>(ring + s * (NL_NF * NT), in, o, s, klev, nproma);
    csc2_cp_async_commit();
  }

  const CritRH crh = make_critrh(tropopause_eta(c, in.pt, in.gt, o.o1, o.ocml, nproma));

  Carry st;
  st.paph0 = ldin(in.paph + o.oh);
  st.rfl = 0.0;
  st.sfl = 0.0;
  // flux rows at the model top (cloudsc2.F90:308-309, :732-733 -> -0.0)
  stout(out.pfplsl + o.oh, 0.0);
  stout(out.pfplsn + o.oh, 0.0);
  stout(out.pfhpsl + o.oh, -0.0 * c.rlvtt);
  stout(out.pfhpsn + o.oh, -0.0 * c.rlstt);

  int slot = 0, pslot = STAGES - 1;
  NlProbe<PROBE> probe;
  for (int jk = 0; jk < klev; ++jk) {
    probe.level(jk);
    const int pf = jk + STAGES - 1;
    if (pf < klev) csc2_stage_traj<NT, false, HAS_PQS ? 1 : 0>(ring + pslot * (NL_NF * NT), in, o, pf, klev, nproma);
    csc2_cp_async_commit();
    csc2_cp_async_wait<STAGES - 1>();
    const LevIn cur = csc2_read_level<NT>(ring + slot * (NL_NF * NT), jk, klev);
    const double pqs = HAS_PQS ? ring[(size_t)slot * (NL_NF * NT) + 15 * NT]
                               : satur_point(c, cur.pt, csc2_rcp(cur.pap));
    LevOut y;
    nl_level<RV>(c, crh, jk, cur, pqs, st, y);

    const size_t l = (size_t)jk * nproma;
    stout(out.tent + o.oloc + l, y.tent);
    stout(out.tenq + o.oloc + l, y.tenq);
    stout(out.tenl + o.oloc + l, y.tenl);
    stout(out.teni + o.oloc + l, y.teni);
    if (out.loc_last) stout(out.loc_last + o.oloc + l, 0.0);
    stout(out.pclc + o.o1 + l, y.pclc);
    stout(out.pcovptot + o.o1 + l, 0.0);
    stout(out.pfplsl + o.oh + l + nproma, y.rfln);
    stout(out.pfplsn + o.oh + l + nproma, y.sfln);
    stout(out.pfhpsl + o.oh + l + nproma, -y.rfln * c.rlvtt);   // cloudsc2.F90:730-735
    stout(out.pfhpsn + o.oh + l + nproma, -y.sfln * c.rlstt);
    slot = (slot + 1 == STAGES) ? 0 : slot + 1;
    pslot = (pslot + 1 == STAGES) ? 0 : pslot + 1;
  }
  if (probe.fired()) stout(out.pcovptot + o.o1, 1.0);
}

#ifdef CSC2_EXPERIMENTS
#include "experiments/cloudsc2_nl_experiments.cuh"   // rejected variants + PROBE builds: never in the product
#endif

// expand_mod.F90:270-302 on the device: dst(nproma, rows, nblocks) <- src(nlon, rows), local
// column j <- source column (gcol0 + j) mod nlon, zero beyond ngptot.  A thread owns one column of the
// blocked array (its block / lane / source column are computed once) and walks the rows with the grid's
// y stride: no per-element division, stores coalesced along NPROMA, the 100 source columns stay in L1/L2.
__global__ void __launch_bounds__(256)
k_expand(const double *__restrict__ src, int nlon, int rows, double *__restrict__ dst, int nproma,
         int ngptot, long long gcol0, int ncol) {
  for (int gcol = blockIdx.x * blockDim.x + threadIdx.x; gcol < ncol; gcol += gridDim.x * blockDim.x) {
    const int b = gcol / nproma, jl = gcol - b * nproma;
    const bool valid = gcol < ngptot;
    const double *sp = src + (int)((gcol0 + gcol) % nlon);
    double *dp = dst + (size_t)b * rows * nproma + jl;
    // four rows in flight per thread: the loads hit L1 / L2 (the source is 100 columns), the stores stream out
#pragma unroll 4
    for (int r = blockIdx.y; r < rows; r += gridDim.y)
      __stcs(dp + (size_t)r * nproma, valid ? __ldg(sp + (size_t)r * nlon) : 0.0);
  }
}

// SATUR on its own (satur.F90:106-123, LDPHYLIN branch): elementwise over n points
__global__ void k_satur(const __grid_constant__ KConst c, const double *__restrict__ pap,
                        const double *__restrict__ pt, double *__restrict__ pqsat, long long n) {
  csc2_math_init();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pqsat[i] = satur_point(c, pt[i], csc2_rcp(pap[i]));
}

// accuracy probe of the branch-free elementary functions (tests/test_gpu_math.py)
__global__ void k_math_probe(int fn, const double *__restrict__ x, double *__restrict__ y, int n) {
  csc2_math_init();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = x[i];
  double r, t;
  switch (fn) {
    case 0: r = csc2_rcp(v); break;
    case 1: r = csc2_exp(v); break;
    case 2: r = csc2_expn(v); break;
    case 3: r = csc2_sqrt(v); break;
    case 4: r = csc2_tanh_p1(v); break;
    default: csc2_tanh_p1_sech2(v, t, r); break;
  }
  y[i] = r;
}

}  // namespace

cudaError_t csc2_launch_satur(const KConst &c, const double *pap, const double *pt, double *pqsat,
                              long long n, cudaStream_t s) {
  k_satur<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(c, pap, pt, pqsat, n);
  return cudaGetLastError();
}

cudaError_t csc2_launch_math_probe(int fn, const double *x, double *y, int n, cudaStream_t s) {
  k_math_probe<<<(n + 127) / 128, 128, 0, s>>>(fn, x, y, n);
  return cudaGetLastError();
}

// The product shape: 2-stage ring, CTA of 128 columns, 128 registers -> 4 CTAs = 16 warps per SM
// (measured at 163 840 columns: 16 warps 0.849 ms, 12 warps 0.885, 14 warps 0.878, 18-20 warps 0.92-0.99:
// more warps cost registers -> instructions, and the kernel is issue-bound).
template <bool HAS_PQS, int STAGES, int NT, int MAXREG, bool RV, int PROBE = 0>
static cudaError_t launch_nl_rv(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
                                cudaStream_t s) {
  const long long ncol = (long long)g.nblocks * g.nproma;
  const int grid = (int)((ncol + NT - 1) / NT);
#ifdef CSC2_EXPERIMENTS
  // CSC2_NL_EXTRA_SMEM_KB (probe): unused extra dynamic shared memory per CTA, to measure how the size of the
  // shared-memory carve-out (= what is left for L1) affects the kernel at unchanged occupancy
  static const size_t extra = [] { const char *e = getenv("CSC2_NL_EXTRA_SMEM_KB"); return e ? (size_t)atoi(e) * 1024 : 0; }();
#else
  const size_t extra = 0;
#endif
  const size_t smem = (size_t)STAGES * NL_NF * NT * sizeof(double) + extra;
  auto kern = k_cloudsc2_nl<HAS_PQS, STAGES, NT, MAXREG, RV, PROBE>;
  static CSC2_SMEM_FLAGS smem_ok{0};
  if (cudaError_t e0 = csc2_allow_smem(kern, smem, smem_ok)) return e0;
  kern<<<grid, NT, smem, s>>>(c, g, in, out);
  return cudaGetLastError();
}

#ifdef CSC2_EXPERIMENTS
#include "experiments/cloudsc2_nl_experiments_launch.cuh"
#else
void csc2_set_nl_variant(int) {}
#endif

cudaError_t csc2_launch_nl(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
                           cudaStream_t s) {
#ifdef CSC2_EXPERIMENTS
  const cudaError_t e = csc2_launch_nl_experiment(c, g, in, out, s);
  if (e != cudaErrorNotSupported) return e;
#endif
  // RV = (RVTMP2 != 0) and HAS_PQS are template switches so that the level stays one basic block
  if (c.rvtmp2 != 0.0) {
    if (in.pqs) return launch_nl_rv<true, 2, 128, 128, true>(c, g, in, out, s);
    return launch_nl_rv<false, 2, 128, 128, true>(c, g, in, out, s);
  }
  if (in.pqs) return launch_nl_rv<true, 2, 128, 128, false>(c, g, in, out, s);
  return launch_nl_rv<false, 2, 128, 128, false>(c, g, in, out, s);
}

// Forward (trajectory) sweep of the adjoint: the NL kernel itself, without the driver-level zeroing.  Its
// outputs PFPLSL5 / PFPLSN5 (cloudsc2ad.F90:847-848) are the flux check-points the reverse sweep restarts from.
cudaError_t csc2_launch_nl_traj(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
                                cudaStream_t s) {
  TrajOut o = out;
  o.loc_last = nullptr;
  if (c.rvtmp2 != 0.0) {
    if (in.pqs) return launch_nl_rv<true, 2, 128, 128, true>(c, g, in, o, s);
    return launch_nl_rv<false, 2, 128, 128, true>(c, g, in, o, s);
  }
  if (in.pqs) return launch_nl_rv<true, 2, 128, 128, false>(c, g, in, o, s);
  return launch_nl_rv<false, 2, 128, 128, false>(c, g, in, o, s);
}

cudaError_t csc2_launch_expand(const double *src, int nlon, long long rows, double *dst, int nproma,
                               int ngptot, int nblocks, long long gcol0, cudaStream_t s) {
  const long long ncol = (long long)nproma * nblocks;
  if (ncol > 0x7fffffffLL || rows > 0x7fffffffLL) return cudaErrorInvalidValue;
  long long gx = (ncol + 255) / 256;
  if (gx > 148LL * 8) gx = 148LL * 8;
  // enough CTAs to fill the machine a few times over, never more rows than there are
  long long gy = (148LL * 16 + gx - 1) / gx;
  if (gy > rows) gy = rows;
  if (gy < 1) gy = 1;
  k_expand<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, s>>>(src, nlon, (int)rows, dst, nproma, ngptot, gcol0, (int)ncol);
  return cudaGetLastError();
}

cudaError_t csc2_upload_levels_nl(const double *ceta, const double *zscalm, const double *sq1mceta,
                                  int klev, cudaStream_t s) {
  return csc2_upload_levels_impl(ceta, zscalm, sq1mceta, klev, s);
}
