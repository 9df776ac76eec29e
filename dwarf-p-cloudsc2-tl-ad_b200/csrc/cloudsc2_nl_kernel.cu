// cloudsc2_nl_kernel.cu -- the nonlinear kernel: fused SATUR + CLOUDSC2 for every column of every
// NPROMA block in one launch.  Replaces the OpenMP block loop of CLOUDSC_DRIVER
// (reference src/cloudsc2_nl/cloudsc_driver_mod.F90:82-111).
//
// Mapping: thread <-> column, CTA = 128 consecutive columns, the KLEV loop runs in registers with
// the next level's 15 inputs prefetched while the current level is computed.  Loads/stores are
// coalesced along NPROMA (JL); every input is read exactly once (plus the ~40-level tropopause
// pre-pass over PT/PGTENT, which the main sweep re-reads out of L2), PQSAT never touches memory.
#include "cloudsc2_nl.cuh"
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

template <bool HAS_PQS>
__global__ void __launch_bounds__(CSC2_NL_THREADS)
k_cloudsc2_nl(const __grid_constant__ KConst c, const Geom g, const TrajIn in, const TrajOut out) {
  const int gcol = blockIdx.x * blockDim.x + threadIdx.x;
  const int ibl = gcol / g.nproma;
  if (ibl >= g.nblocks) return;
  const int jl = gcol - ibl * g.nproma;
  const int klev = g.klev, nproma = g.nproma;
  const size_t n2 = (size_t)nproma * klev;
  ColOffsets o;
  o.o1 = (size_t)ibl * n2 + jl;
  o.oh = (size_t)ibl * (n2 + nproma) + jl;
  o.ocld = (size_t)ibl * in.bs_cld + jl;
  o.ocml = (size_t)ibl * in.bs_cml + jl;
  o.oloc = (size_t)ibl * out.bs_loc + jl;

  if (gcol >= g.ngptot) {
    // padding column of the last block: the driver zeroes whole blocks (driver_mod.F90:87-88),
    // the kernel itself never computes columns beyond ICEND.
    for (int jk = 0; jk < klev; ++jk) {
      stout(out.pcovptot + o.o1 + (size_t)jk * nproma, 0.0);
      if (out.loc_last) stout(out.loc_last + o.oloc + (size_t)jk * nproma, 0.0);
    }
    return;
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

  LevIn cur = load_level(in, o, 0, klev, nproma);
  double pqs_cur = HAS_PQS ? ldin(in.pqs + o.o1) : 0.0;
  for (int jk = 0; jk < klev; ++jk) {
    LevIn nxt = cur;
    double pqs_nxt = 0.0;
    if (jk + 1 < klev) {
      nxt = load_level(in, o, jk + 1, klev, nproma);
      if (HAS_PQS) pqs_nxt = ldin(in.pqs + o.o1 + (size_t)(jk + 1) * nproma);
    }
    const double pqs = HAS_PQS ? pqs_cur : satur_point(c, cur.pt, 1.0 / cur.pap);
    LevOut y;
    nl_level(c, crh, jk, cur, pqs, st, y);

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
    cur = nxt;
    pqs_cur = pqs_nxt;
  }
}

// expand_mod.F90:270-302 on the device: dst(nproma, rows, nblocks) <- src(nlon, rows), local
// column j <- source column (gcol0 + j) mod nlon, zero beyond ngptot.
__global__ void k_expand(const double *__restrict__ src, int nlon, long long rows,
                         double *__restrict__ dst, int nproma, int ngptot, long long gcol0,
                         long long total) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; idx < total; idx += stride) {
    const int jl = (int)(idx % nproma);
    const long long t = idx / nproma;
    const long long r = t % rows;
    const long long b = t / rows;
    const long long gcol = b * nproma + jl;
    dst[idx] = (gcol < ngptot) ? __ldg(src + r * nlon + ((gcol0 + gcol) % nlon)) : 0.0;
  }
}

// accuracy probe of the branch-free elementary functions (tests/test_gpu_math.py)
__global__ void k_math_probe(int fn, const double *__restrict__ x, double *__restrict__ y, int n) {
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

cudaError_t csc2_launch_math_probe(int fn, const double *x, double *y, int n, cudaStream_t s) {
  k_math_probe<<<(n + 127) / 128, 128, 0, s>>>(fn, x, y, n);
  return cudaGetLastError();
}

cudaError_t csc2_launch_nl(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
                           cudaStream_t s) {
  const long long ncol = (long long)g.nblocks * g.nproma;
  const int grid = (int)((ncol + CSC2_NL_THREADS - 1) / CSC2_NL_THREADS);
  if (in.pqs) k_cloudsc2_nl<true><<<grid, CSC2_NL_THREADS, 0, s>>>(c, g, in, out);
  else k_cloudsc2_nl<false><<<grid, CSC2_NL_THREADS, 0, s>>>(c, g, in, out);
  return cudaGetLastError();
}

cudaError_t csc2_launch_expand(const double *src, int nlon, long long rows, double *dst, int nproma,
                               int ngptot, int nblocks, long long gcol0, cudaStream_t s) {
  const long long total = (long long)nproma * rows * nblocks;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  k_expand<<<(int)blocks, 256, 0, s>>>(src, nlon, rows, dst, nproma, ngptot, gcol0, total);
  return cudaGetLastError();
}
