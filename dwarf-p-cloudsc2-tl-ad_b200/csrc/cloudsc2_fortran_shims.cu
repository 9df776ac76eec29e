// cloudsc2_fortran_shims.cu -- link-time substitutes for the reference's four kernel subroutines,
// with the gfortran calling convention (external name lower case + underscore, every argument by
// reference, explicit-shape arrays as bare pointers, LOGICAL as 4-byte int):
//     SATUR      (satur.F90:10)        -> satur_
//     CLOUDSC2   (cloudsc2.F90:10)     -> cloudsc2_
//     CLOUDSC2TL (cloudsc2tl.F90:10)   -> cloudsc2tl_
//     CLOUDSC2AD (cloudsc2ad.F90:10)   -> cloudsc2ad_
// With these, the reference's driver modules compile and link UNCHANGED (only the four kernel
// object files are left out); each call processes one NPROMA block: H2D of the block's arrays,
// one kernel launch, D2H.  Functional, not fast -- the throughput path is the
// whole-problem entry cloudsc2_gpu_nl & co.  The module constants the Fortran kernels USE are not
// arguments, so cloudsc2_gpu_init must have been called (fortran/cloudsc2_gpu_mod.F90:
// CLOUDSC2_GPU_SETUP); errors abort like ABOR1 (abor1.F90:10-14).  The reference calls these from
// inside an OpenMP parallel region (cloudsc_driver_mod.F90:73-81): calls are serialised by a mutex.
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "cloudsc2_launch.h"
#include "../../include/cloudsc2_fortran.h"   // prototypes must agree with the definitions below

// provided by cloudsc2_api.cu
int csc2_shim_context(KConst *kc, double ptsphy, int klev, cudaStream_t *stream, long long *launches);
void csc2_shim_count_launch();

namespace {

std::mutex g_mu;

[[noreturn]] void abor1(const char *msg) {
  fprintf(stderr, " CLOUDSC2 GPU shim: %s\n", msg);    // ABOR1: WRITE(0,*) + abort()
  abort();
}
#define CKA(call)                                                              \
  do {                                                                         \
    cudaError_t e_ = (call);                                                   \
    if (e_ != cudaSuccess) abor1(cudaGetErrorString(e_));                      \
  } while (0)

// one device slab holding n arrays of `each` doubles
struct Slab {
  double *p = nullptr;
  size_t cap = 0;
  double *get(size_t doubles) {
    if (doubles > cap) {
      if (p) cudaFree(p);
      CKA(cudaMalloc(&p, doubles * sizeof(double)));
      cap = doubles;
    }
    return p;
  }
};
Slab g_slab;

struct Mover {
  cudaStream_t s;
  double *next;
  std::vector<std::pair<double *, std::pair<double *, size_t>>> back;   // device -> host on finish
  double *in(const double *h, size_t n) {
    double *d = next;
    next += n;
    CKA(cudaMemcpyAsync(d, h, n * sizeof(double), cudaMemcpyHostToDevice, s));
    return d;
  }
  double *inout(double *h, size_t n) {
    double *d = in(h, n);
    back.push_back({h, {d, n}});
    return d;
  }
  void finish() {
    for (auto &b : back)
      CKA(cudaMemcpyAsync(b.first, b.second.first, b.second.second * sizeof(double),
                          cudaMemcpyDeviceToHost, s));
    CKA(cudaStreamSynchronize(s));
  }
};

void check_dims(const int *kidia, const int *kfdia, const int *klon, const int *ktdia, const int *ldrain1d) {
  if (*kidia != 1 || *ktdia != 1) abor1("KIDIA and KTDIA must be 1 (whole blocks from the top level)");
  if (*kfdia < 1 || *kfdia > *klon) abor1("KFDIA outside 1..KLON");
  if (ldrain1d && *ldrain1d) abor1("LDRAIN1D=.TRUE. is not supported (the dwarf never sets it)");
}

}  // namespace

extern "C" {

void satur_(const int *kidia, const int *kfdia, const int *klon, const int *ktdia, const int *klev,
            const int *ldphylin, const double *paprsf, const double *pt, double *pqsat,
            const int *kflag) {
  std::lock_guard<std::mutex> lk(g_mu);
  (void)kflag;   // KFLAG only selects FOEEWM vs FOEEWMCU in the non-LDPHYLIN branch (satur.F90:126-140)
  check_dims(kidia, kfdia, klon, ktdia, nullptr);
  if (!*ldphylin) abor1("SATUR: only the LDPHYLIN=.TRUE. branch exists on the GPU");
  KConst kc; cudaStream_t s; long long dummy;
  if (csc2_shim_context(&kc, 1.0, *klev, &s, &dummy)) abor1("cloudsc2_gpu_init has not been called (or KLEV differs)");
  const size_t n2 = (size_t)*klon * *klev;
  Mover m{s, g_slab.get(3 * n2), {}};
  const double *dp = m.in(paprsf, n2), *dt = m.in(pt, n2);
  double *dq = m.inout(pqsat, n2);       // columns beyond KFDIA keep the caller's values
  // the kernel is elementwise; restrict it to KIDIA..KFDIA level by level when KFDIA < KLON
  if (*kfdia == *klon) {
    CKA(csc2_launch_satur(kc, dp, dt, dq, (long long)n2, s));
    csc2_shim_count_launch();
  } else {
    for (int jk = 0; jk < *klev; ++jk) {
      CKA(csc2_launch_satur(kc, dp + (size_t)jk * *klon, dt + (size_t)jk * *klon, dq + (size_t)jk * *klon, *kfdia, s));
      csc2_shim_count_launch();
    }
  }
  m.finish();
}

void cloudsc2_(const int *kidia, const int *kfdia, const int *klon, const int *ktdia, const int *klev,
               const int *ldrain1d, const double *ptsphy,
               const double *paphp1, const double *papp1, const double *pqm1, const double *pqs,
               const double *ptm1, const double *pl, const double *pi, const double *plude,
               const double *plu, const double *pmfu, const double *pmfd,
               double *ptent, const double *pgtent, double *ptenq, const double *pgtenq,
               double *ptenl, const double *pgtenl, double *pteni, const double *pgteni,
               const double *psupsat, double *pclc, double *pfplsl, double *pfplsn, double *pfhpsl,
               double *pfhpsn, double *pcovptot) {
  std::lock_guard<std::mutex> lk(g_mu);
  check_dims(kidia, kfdia, klon, ktdia, ldrain1d);
  KConst kc; cudaStream_t s; long long dummy;
  if (csc2_shim_context(&kc, *ptsphy, *klev, &s, &dummy)) abor1("cloudsc2_gpu_init has not been called (or KLEV differs)");
  const size_t n2 = (size_t)*klon * *klev, n2h = (size_t)*klon * (*klev + 1);
  Mover m{s, g_slab.get(21 * n2 + 5 * n2h), {}};
  TrajIn in{};
  in.paph = m.in(paphp1, n2h); in.pap = m.in(papp1, n2); in.pq = m.in(pqm1, n2); in.pqs = m.in(pqs, n2);
  in.pt = m.in(ptm1, n2); in.pl = m.in(pl, n2); in.pi = m.in(pi, n2); in.plude = m.in(plude, n2);
  in.plu = m.in(plu, n2); in.pmfu = m.in(pmfu, n2); in.pmfd = m.in(pmfd, n2);
  in.gt = m.in(pgtent, n2); in.gq = m.in(pgtenq, n2); in.gl = m.in(pgtenl, n2); in.gi = m.in(pgteni, n2);
  in.psupsat = m.in(psupsat, n2);
  in.bs_cld = in.bs_cml = (long long)n2;
  TrajOut out{};
  out.tent = m.inout(ptent, n2); out.tenq = m.inout(ptenq, n2); out.tenl = m.inout(ptenl, n2);
  out.teni = m.inout(pteni, n2); out.pclc = m.inout(pclc, n2); out.pfplsl = m.inout(pfplsl, n2h);
  out.pfplsn = m.inout(pfplsn, n2h); out.pfhpsl = m.inout(pfhpsl, n2h); out.pfhpsn = m.inout(pfhpsn, n2h);
  out.pcovptot = m.inout(pcovptot, n2);
  out.loc_last = nullptr;      // zeroing TENDENCY_LOC%CLD(:,:,NCLV) is the driver's job (:88)
  out.bs_loc = (long long)n2;
  Geom g{*klon, *klev, *kfdia, 1};
  CKA(csc2_launch_nl(kc, g, in, out, s));
  csc2_shim_count_launch();
  m.finish();
}

#define TRAJ_ARGS                                                                                   \
  const double *paphp15, const double *papp15, const double *pqm15, const double *pqs5,             \
      const double *ptm15, const double *pl5, const double *pi5, const double *plude5,              \
      const double *plu5, const double *pmfu5, const double *pmfd5, double *ptent5,                 \
      const double *pgtent5, double *ptenq5, const double *pgtenq5, double *ptenl5,                 \
      const double *pgtenl5, double *pteni5, const double *pgteni5, const double *psupsat5,         \
      double *pclc5, double *pfplsl5, double *pfplsn5, double *pfhpsl5, double *pfhpsn5,            \
      double *pcovptot5
#define INCR_ARGS                                                                                   \
  double *paphp1, double *papp1, double *pqm1, double *pqs, double *ptm1, double *pl, double *pi,   \
      double *plude, double *plu, double *pmfu, double *pmfd, double *ptent, double *pgtent,        \
      double *ptenq, double *pgtenq, double *ptenl, double *pgtenl, double *pteni, double *pgteni,  \
      double *psupsat, double *pclc, double *pfplsl, double *pfplsn, double *pfhpsl,                \
      double *pfhpsn, double *pcovptot

static void tlad(bool is_ad, const int *kidia, const int *kfdia, const int *klon, const int *ktdia,
                 const int *klev, const int *ldrain1d, const double *ptsphy, TRAJ_ARGS, INCR_ARGS) {
  std::lock_guard<std::mutex> lk(g_mu);
  check_dims(kidia, kfdia, klon, ktdia, ldrain1d);
  KConst kc; cudaStream_t s; long long dummy;
  if (csc2_shim_context(&kc, *ptsphy, *klev, &s, &dummy)) abor1("cloudsc2_gpu_init has not been called (or KLEV differs)");
  const size_t n2 = (size_t)*klon * *klev, n2h = (size_t)*klon * (*klev + 1);
  double *base = g_slab.get(2 * (21 * n2 + 5 * n2h));
  Mover m{s, base, {}};
  TrajIn in{};
  in.paph = m.in(paphp15, n2h); in.pap = m.in(papp15, n2); in.pq = m.in(pqm15, n2); in.pqs = m.in(pqs5, n2);
  in.pt = m.in(ptm15, n2); in.pl = m.in(pl5, n2); in.pi = m.in(pi5, n2); in.plude = m.in(plude5, n2);
  in.plu = m.in(plu5, n2); in.pmfu = m.in(pmfu5, n2); in.pmfd = m.in(pmfd5, n2);
  in.gt = m.in(pgtent5, n2); in.gq = m.in(pgtenq5, n2); in.gl = m.in(pgtenl5, n2); in.gi = m.in(pgteni5, n2);
  in.psupsat = m.in(psupsat5, n2);
  in.bs_cld = in.bs_cml = (long long)n2;
  TrajOut out{};
  out.tent = m.inout(ptent5, n2); out.tenq = m.inout(ptenq5, n2); out.tenl = m.inout(ptenl5, n2);
  out.teni = m.inout(pteni5, n2); out.pclc = m.inout(pclc5, n2); out.pfplsl = m.inout(pfplsl5, n2h);
  out.pfplsn = m.inout(pfplsn5, n2h); out.pfhpsl = m.inout(pfhpsl5, n2h); out.pfhpsn = m.inout(pfhpsn5, n2h);
  out.pcovptot = m.inout(pcovptot5, n2);
  out.loc_last = nullptr;
  out.bs_loc = (long long)n2;
  // increments: TL reads the 16 and writes the 10; AD reads+zeroes the 10 and accumulates into the 16
  IncIn din{};
  auto mv16 = [&](double *h, size_t n) { return is_ad ? m.inout(h, n) : m.in(h, n); };
  din.paph = mv16(paphp1, n2h); din.pap = mv16(papp1, n2); din.pq = mv16(pqm1, n2); din.pqs = mv16(pqs, n2);
  din.pt = mv16(ptm1, n2); din.pl = mv16(pl, n2); din.pi = mv16(pi, n2); din.plude = mv16(plude, n2);
  din.plu = mv16(plu, n2); din.pmfu = mv16(pmfu, n2); din.pmfd = mv16(pmfd, n2);
  din.gt = mv16(pgtent, n2); din.gq = mv16(pgtenq, n2); din.gl = mv16(pgtenl, n2); din.gi = mv16(pgteni, n2);
  din.psupsat = mv16(psupsat, n2);
  IncOut dout{};
  dout.tent = m.inout(ptent, n2); dout.tenq = m.inout(ptenq, n2); dout.tenl = m.inout(ptenl, n2);
  dout.teni = m.inout(pteni, n2); dout.pclc = m.inout(pclc, n2); dout.pfplsl = m.inout(pfplsl, n2h);
  dout.pfplsn = m.inout(pfplsn, n2h); dout.pfhpsl = m.inout(pfhpsl, n2h); dout.pfhpsn = m.inout(pfhpsn, n2h);
  dout.pcovptot = m.inout(pcovptot, n2);
  Geom g{*klon, *klev, *kfdia, 1};
  if (is_ad) {
    ADOpts opt{0.0, 0, nullptr, 0, 0};
    CKA(csc2_launch_ad(kc, g, in, out, din, dout, opt, s));
    csc2_shim_count_launch();      // forward + reverse sweep = two launches
  } else {
    TLOpts opt{0.0, 0, nullptr, nullptr, 0};
    CKA(csc2_launch_tl(kc, g, in, out, din, dout, opt, s));
  }
  csc2_shim_count_launch();
  m.finish();
}

void cloudsc2tl_(const int *kidia, const int *kfdia, const int *klon, const int *ktdia, const int *klev,
                 const int *ldrain1d, const double *ptsphy, TRAJ_ARGS, INCR_ARGS) {
  tlad(false, kidia, kfdia, klon, ktdia, klev, ldrain1d, ptsphy, paphp15, papp15, pqm15, pqs5, ptm15, pl5,
       pi5, plude5, plu5, pmfu5, pmfd5, ptent5, pgtent5, ptenq5, pgtenq5, ptenl5, pgtenl5, pteni5, pgteni5,
       psupsat5, pclc5, pfplsl5, pfplsn5, pfhpsl5, pfhpsn5, pcovptot5, paphp1, papp1, pqm1, pqs, ptm1, pl, pi,
       plude, plu, pmfu, pmfd, ptent, pgtent, ptenq, pgtenq, ptenl, pgtenl, pteni, pgteni, psupsat, pclc,
       pfplsl, pfplsn, pfhpsl, pfhpsn, pcovptot);
}
void cloudsc2ad_(const int *kidia, const int *kfdia, const int *klon, const int *ktdia, const int *klev,
                 const int *ldrain1d, const double *ptsphy, TRAJ_ARGS, INCR_ARGS) {
  tlad(true, kidia, kfdia, klon, ktdia, klev, ldrain1d, ptsphy, paphp15, papp15, pqm15, pqs5, ptm15, pl5,
       pi5, plude5, plu5, pmfu5, pmfd5, ptent5, pgtent5, ptenq5, pgtenq5, ptenl5, pgtenl5, pteni5, pgteni5,
       psupsat5, pclc5, pfplsl5, pfplsn5, pfhpsl5, pfhpsn5, pcovptot5, paphp1, papp1, pqm1, pqs, ptm1, pl, pi,
       plude, plu, pmfu, pmfd, ptent, pgtent, ptenq, pgtenq, ptenl, pgtenl, pteni, pgteni, psupsat, pclc,
       pfplsl, pfplsn, pfhpsl, pfhpsn, pcovptot);
}

}  // extern "C"
