// cloudsc2_host.cc -- host-side mirror of the reference's state container, expansion and
// validation, plus the synthetic stand-in for the missing config-files/input.h5.
// Interfaces and citations: include/cloudsc2_host.h.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "cloudsc2_host.h"

namespace {

// splitmix64: tiny, seedable, identical on every platform (no libm, no std::distribution).
struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
  double uni(double a, double b) { return a + (b - a) * uni(); }
  // sum of 4 uniforms, variance 1/3 -> scaled to unit variance; bounded "normal-like" noise
  double gauss() { return (uni() + uni() + uni() + uni() - 2.0) * 1.7320508075688772; }
};

double *dalloc(size_t n) { return static_cast<double *>(std::calloc(n ? n : 1, sizeof(double))); }

// SATUR formula (satur.F90:106-123) used only to place the synthetic humidity relative to
// saturation; it is input generation, not the product path.
double qsat_for_synth(const cloudsc2_params *p, double t, double pres) {
  double x = (std::max(p->rtice, std::min(p->rtwat, t)) - p->rtice) * p->rtwat_rtice_r;
  double alfa = std::min(1.0, x * x);
  double el = p->r2es * std::exp(p->r3les * (t - p->rtt) / (t - p->r4les));
  double ei = p->r2es * std::exp(p->r3ies * (t - p->rtt) / (t - p->r4ies));
  double qs = std::min(0.5, (alfa * el + (1.0 - alfa) * ei) / pres);
  return qs / (1.0 - p->retv * qs);
}

}  // namespace

extern "C" {

void cloudsc2_default_params(cloudsc2_params *p) {
  std::memset(p, 0, sizeof(*p));
  p->rg = 9.80665;
  p->rd = 287.0596736665907;      // R/RMD, R=8.31451*... IFS: 1000*R/28.9644
  const double rv = 461.5249933083879;
  p->rcpd = 3.5 * p->rd;
  p->retv = rv / p->rd - 1.0;
  p->rlvtt = 2.5008e6;
  p->rlstt = 2.8345e6;
  p->rlmlt = p->rlstt - p->rlvtt;
  p->rtt = 273.16;
  p->r2es = 611.21 * p->rd / rv;
  p->r3les = 17.502;
  p->r3ies = 22.587;
  p->r4les = 32.19;
  p->r4ies = -0.7;
  p->r5les = p->r3les * (p->rtt - p->r4les);
  p->r5ies = p->r3ies * (p->rtt - p->r4ies);
  p->r5alvcp = p->r5les * p->rlvtt / p->rcpd;
  p->r5alscp = p->r5ies * p->rlstt / p->rcpd;
  p->ralvdcp = p->rlvtt / p->rcpd;
  p->ralsdcp = p->rlstt / p->rcpd;
  p->rtwat = p->rtt;
  p->rtice = p->rtt - 23.0;
  p->rtwat_rtice_r = 1.0 / (p->rtwat - p->rtice);
  p->rvtmp2 = 0.0;  // never loaded by the reference (yoethf.F90:30 vs :79-99) -> static 0
  p->rclcrit = 4.0e-4;
  p->rkconv = 1.0 / 6000.0;
  p->rlmin = 1.0e-8;
  p->rpecons = 5.547e-5;
  p->rlptrc = p->rtice + (p->rtwat - p->rtice) / std::sqrt(2.0);
  p->lphylin = 1;
  p->levapls2 = 0;
  p->lregcl = 0;
  p->ldrain1d = 0;
}

int cloudsc2_nblocks(int ngptot, int nproma) {
  // cloudsc_driver_mod.F90:62 : NGPBLKS = (NGPTOT / NPROMA) + MIN(MOD(NGPTOT,NPROMA), 1)
  return ngptot / nproma + std::min(ngptot % nproma, 1);
}

int cloudsc2_source_synth(cloudsc2_source *s, unsigned long long seed, int klon, int klev,
                          const cloudsc2_params *p) {
  if (!s || klon <= 0 || klev < 8) return 1;
  std::memset(s, 0, sizeof(*s));
  s->klon = klon;
  s->klev = klev;
  s->ptsphy = 3600.0;
  const size_t n = (size_t)klon * klev;
  s->pt = dalloc(n); s->pq = dalloc(n); s->pap = dalloc(n); s->paph = dalloc(n + klon);
  s->plu = dalloc(n); s->plude = dalloc(n); s->pmfu = dalloc(n); s->pmfd = dalloc(n);
  s->pa = dalloc(n); s->psupsat = dalloc(n);
  s->pclv = dalloc(n * CLOUDSC2_NCLV);
  s->tend_cml = dalloc(n * CLOUDSC2_NSTATE);
  s->ceta = dalloc(klev);
  if (!s->pt || !s->pq || !s->pap || !s->paph || !s->plu || !s->plude || !s->pmfu || !s->pmfd ||
      !s->pa || !s->psupsat || !s->pclv || !s->tend_cml || !s->ceta) {
    cloudsc2_source_free(s);
    return 2;
  }
#define S2(a, jl, jk) (a)[(size_t)(jk) * klon + (jl)]
#define S3(a, jl, jk, f) (a)[((size_t)(f) * klev + (jk)) * klon + (jl)]
  for (int jl = 0; jl < klon; ++jl) {
    Rng r(seed * 1000003ull + (uint64_t)jl);
    // hybrid-sigma-like half levels: ~1 Pa at the top, surface pressure 900..1030 hPa
    const double ps = r.uni(9.0e4, 1.03e5);
    const double ptop = 1.0;
    for (int k = 0; k <= klev; ++k) {
      double x = (double)k / klev;
      S2(s->paph, jl, k) = ptop * (1.0 - x) + ps * std::pow(x, 2.6);
    }
    // temperature: troposphere T = Ts*(p/ps)^0.19 above a stratosphere warming with height
    const double ts = r.uni(252.0, 303.0);
    const double ttrop = r.uni(205.0, 222.0);
    const double ptrop = r.uni(1.2e4, 3.0e4);           // tropopause inside CETA in (0.1,0.4)
    const double inv = (r.uni() < 0.3) ? r.uni(1.0, 4.0) : 0.0;  // low-level inversion
    const double moist = r.uni(0.55, 1.0);               // column moisture regime
    for (int k = 0; k < klev; ++k) {
      const double pf = 0.5 * (S2(s->paph, jl, k) + S2(s->paph, jl, k + 1));
      S2(s->pap, jl, k) = pf;
      double t;
      if (pf < ptrop) {
        t = ttrop + 12.0 * std::log10(ptrop / pf);       // stratosphere: warmer aloft
      } else {
        const double tt = ts * std::pow(pf / ps, 0.19);
        const double t0 = ts * std::pow(ptrop / ps, 0.19);
        t = tt + (ttrop - t0) * std::pow(ptrop / pf, 3.0);  // blend to the tropopause value
        if (pf > 0.92 * ps) t -= inv * (pf / ps - 0.92) / 0.08;
      }
      t += 0.4 * r.gauss();
      S2(s->pt, jl, k) = t;
      const double qs = qsat_for_synth(p, t, pf);
      // relative humidity 0.2 .. 1.1 with a wet bias in mid/low troposphere
      double rh = r.uni(0.2, 1.1);
      if (pf > 3.0e4 && r.uni() < moist) rh = r.uni(0.75, 1.1);
      if (pf < ptrop) rh = r.uni(0.02, 0.6);
      S2(s->pq, jl, k) = rh * qs;
      if (pf < ptrop) S2(s->pq, jl, k) = std::min(rh * qs, 3.0e-6 * (1.0 + r.uni()));  // dry stratosphere
      // prior cloud condensate where it is moist
      const bool cloudy = rh > 0.7 && pf > 1.0e4;
      S3(s->pclv, jl, k, 0) = (cloudy && t > 250.0) ? r.uni(0.0, 1.0e-5) : 0.0;  // QL
      S3(s->pclv, jl, k, 1) = (cloudy && t < p->rtt) ? r.uni(0.0, 1.0e-5) : 0.0; // QI
      S3(s->pclv, jl, k, 2) = (cloudy) ? r.uni(0.0, 1.0e-6) : 0.0;               // QR (unused)
      S3(s->pclv, jl, k, 3) = (cloudy) ? r.uni(0.0, 1.0e-6) : 0.0;               // QS (unused)
      S3(s->pclv, jl, k, 4) = S2(s->pq, jl, k);                                  // QV (unused)
      S2(s->pa, jl, k) = cloudy ? r.uni(0.0, 1.0) : 0.0;
      // convection (troposphere only)
      const bool conv = pf > ptrop;
      S2(s->plude, jl, k) = (conv && r.uni() < 0.10) ? r.uni(0.0, 1.0e-6) : 0.0;
      S2(s->plu, jl, k) = (conv && r.uni() < 0.8) ? r.uni(0.0, 1.0e-4) : 0.0;
      S2(s->pmfu, jl, k) = conv ? r.uni(0.0, 0.05) : 0.0;
      S2(s->pmfd, jl, k) = conv ? -r.uni(0.0, 0.02) : 0.0;
      S2(s->psupsat, jl, k) = (r.uni() < 0.05) ? r.uni(0.0, 1.0e-6) : 0.0;
      // cumulative tendencies
      S3(s->tend_cml, jl, k, 0) = 1.0e-5 * r.gauss();                            // T
      S3(s->tend_cml, jl, k, 1) = 0.0;                                           // A
      S3(s->tend_cml, jl, k, 2) = 1.0e-9 * r.gauss() * std::min(1.0, qs / 1e-3); // Q
      S3(s->tend_cml, jl, k, 3) = 2.0e-10 * r.uni(-0.2, 1.0);                    // QL
      S3(s->tend_cml, jl, k, 4) = 2.0e-10 * r.uni(-0.2, 1.0);                    // QI
      S3(s->tend_cml, jl, k, 5) = 0.0;
      S3(s->tend_cml, jl, k, 6) = 0.0;
      S3(s->tend_cml, jl, k, 7) = 0.0;
    }
  }
  // dwarf_cloudsc.F90:100-102 : CETA(JK) = PAP(1,JK,1)/PAPH(1,KLEV+1,1) (column 1 for ALL columns)
  for (int k = 0; k < klev; ++k) s->ceta[k] = S2(s->pap, 0, k) / S2(s->paph, 0, klev);
#undef S2
#undef S3
  return 0;
}

void cloudsc2_source_free(cloudsc2_source *s) {
  if (!s) return;
  std::free(s->pt); std::free(s->pq); std::free(s->pap); std::free(s->paph); std::free(s->plu);
  std::free(s->plude); std::free(s->pmfu); std::free(s->pmfd); std::free(s->pa);
  std::free(s->psupsat); std::free(s->pclv); std::free(s->tend_cml); std::free(s->ceta);
  std::memset(s, 0, sizeof(*s));
}

void cloudsc2_expand_host(const double *src, int nlon, int nlev, int ndim, double *dst,
                          int nproma, int ngptot) {
  const int nblocks = cloudsc2_nblocks(ngptot, nproma);
  const size_t rows = (size_t)nlev * ndim;  // (nlev, ndim) collapse: both layouts keep them adjacent
#pragma omp parallel for schedule(static)
  for (int b = 0; b < nblocks; ++b) {
    const int g0 = b * nproma;
    const int bsize = std::min(nproma, ngptot - g0);
    for (size_t r = 0; r < rows; ++r) {
      double *d = dst + ((size_t)b * rows + r) * nproma;
      const double *sr = src + r * nlon;
      for (int jl = 0; jl < bsize; ++jl) d[jl] = sr[(g0 + jl) % nlon];
      for (int jl = bsize; jl < nproma; ++jl) d[jl] = 0.0;   // expand_mod.F90:298
    }
  }
}

int cloudsc2_state_load(cloudsc2_state *st, const cloudsc2_source *s, int nproma, int ngptot) {
  if (!st || !s || nproma <= 0 || ngptot <= 0) return 1;
  std::memset(st, 0, sizeof(*st));
  st->nproma = nproma; st->klev = s->klev; st->ngptot = ngptot;
  st->nblocks = cloudsc2_nblocks(ngptot, nproma);
  const size_t n = (size_t)nproma * s->klev * st->nblocks;
  const size_t nh = (size_t)nproma * (s->klev + 1) * st->nblocks;
  double *pt = dalloc(n), *pq = dalloc(n), *pap = dalloc(n), *paph = dalloc(nh), *plu = dalloc(n),
         *plude = dalloc(n), *pmfu = dalloc(n), *pmfd = dalloc(n), *psupsat = dalloc(n),
         *pclv = dalloc(n * CLOUDSC2_NCLV), *b_cml = dalloc(n * CLOUDSC2_NSTATE);
  st->f.pt = pt; st->f.pq = pq; st->f.pap = pap; st->f.paph = paph; st->f.plu = plu;
  st->f.plude = plude; st->f.pmfu = pmfu; st->f.pmfd = pmfd; st->f.psupsat = psupsat;
  st->f.pclv = pclv; st->f.b_cml = b_cml;
  st->f.b_loc = dalloc(n * CLOUDSC2_NSTATE);
  st->f.pa = dalloc(n);
  st->f.pcovptot = dalloc(n);
  st->f.pfplsl = dalloc(nh); st->f.pfplsn = dalloc(nh);
  st->f.pfhpsl = dalloc(nh); st->f.pfhpsn = dalloc(nh);
  if (!pt || !pq || !pap || !paph || !plu || !plude || !pmfu || !pmfd || !psupsat || !pclv ||
      !b_cml || !st->f.b_loc || !st->f.pa || !st->f.pcovptot || !st->f.pfplsl || !st->f.pfplsn ||
      !st->f.pfhpsl || !st->f.pfhpsn) {
    cloudsc2_state_free(st);
    return 2;
  }
  // cloudsc2_array_state_mod.F90:167-183
  cloudsc2_expand_host(s->pt, s->klon, s->klev, 1, pt, nproma, ngptot);
  cloudsc2_expand_host(s->pq, s->klon, s->klev, 1, pq, nproma, ngptot);
  cloudsc2_expand_host(s->pap, s->klon, s->klev, 1, pap, nproma, ngptot);
  cloudsc2_expand_host(s->paph, s->klon, s->klev + 1, 1, paph, nproma, ngptot);
  cloudsc2_expand_host(s->plu, s->klon, s->klev, 1, plu, nproma, ngptot);
  cloudsc2_expand_host(s->plude, s->klon, s->klev, 1, plude, nproma, ngptot);
  cloudsc2_expand_host(s->pmfu, s->klon, s->klev, 1, pmfu, nproma, ngptot);
  cloudsc2_expand_host(s->pmfd, s->klon, s->klev, 1, pmfd, nproma, ngptot);
  cloudsc2_expand_host(s->pa, s->klon, s->klev, 1, st->f.pa, nproma, ngptot);
  cloudsc2_expand_host(s->pclv, s->klon, s->klev, CLOUDSC2_NCLV, pclv, nproma, ngptot);
  cloudsc2_expand_host(s->psupsat, s->klon, s->klev, 1, psupsat, nproma, ngptot);
  cloudsc2_expand_host(s->tend_cml, s->klon, s->klev, CLOUDSC2_NSTATE, b_cml, nproma, ngptot);
  return 0;
}

void cloudsc2_state_free(cloudsc2_state *st) {
  if (!st) return;
  cloudsc2_fields *f = &st->f;
  std::free(const_cast<double *>(f->pt)); std::free(const_cast<double *>(f->pq));
  std::free(const_cast<double *>(f->pap)); std::free(const_cast<double *>(f->paph));
  std::free(const_cast<double *>(f->plu)); std::free(const_cast<double *>(f->plude));
  std::free(const_cast<double *>(f->pmfu)); std::free(const_cast<double *>(f->pmfd));
  std::free(const_cast<double *>(f->psupsat)); std::free(const_cast<double *>(f->pclv));
  std::free(const_cast<double *>(f->b_cml));
  std::free(f->b_loc); std::free(f->pa); std::free(f->pcovptot); std::free(f->pfplsl);
  std::free(f->pfplsn); std::free(f->pfhpsl); std::free(f->pfhpsn);
  std::memset(st, 0, sizeof(*st));
}

void cloudsc2_validate_host(const double *ref, const double *field, int nproma, int nlev,
                            int ngptot, double out[5]) {
  // validate_mod.F90:165-211 : min/max over whole blocks (incl. padding), errors over bsize
  const int nblocks = cloudsc2_nblocks(ngptot, nproma);
  double zmin = std::numeric_limits<double>::max(), zmax = -std::numeric_limits<double>::max();
  double zmaxerr = 0.0, zsumerr = 0.0, zsumref = 0.0;
  for (int b = 0; b < nblocks; ++b) {
    const int bsize = std::min(nproma, ngptot - b * nproma);
    for (int jk = 0; jk < nlev; ++jk) {
      const size_t o = ((size_t)b * nlev + jk) * nproma;
      for (int jl = 0; jl < nproma; ++jl) {
        zmin = std::min(zmin, field[o + jl]);
        zmax = std::max(zmax, field[o + jl]);
      }
      for (int jl = 0; jl < bsize; ++jl) {
        // a NaN result must fail the validation (std::max drops NaN): give it the largest error
        const double d0 = std::fabs(field[o + jl] - ref[o + jl]);
        const double d = (d0 == d0) ? d0 : std::numeric_limits<double>::max();
        zmaxerr = std::max(zmaxerr, d);
        zsumerr += d;
        zsumref += std::fabs(ref[o + jl]);
      }
    }
  }
  out[0] = zmin; out[1] = zmax; out[2] = zmaxerr; out[3] = zsumerr; out[4] = zsumref;
}

double cloudsc2_error_rel(const double stats[5], int *flag_out) {
  // validate_mod.F90:263-296 ERROR_PRINT: relative error in %, flag = the "!!!!" warning
  const double zeps = std::numeric_limits<double>::epsilon();
  const double zerrsum = stats[3], zsum = stats[4];
  double zrelerr;
  if (zerrsum < zeps) zrelerr = 0.0;
  else if (zsum < zeps) zrelerr = zerrsum / (1.0 + zsum);
  else zrelerr = zerrsum / zsum;
  // the "!!!!" warning; a non-finite statistic (NaN results or NaN reference values) is always flagged --
  // `NaN > x` is false, which would otherwise let it pass silently
  bool finite = true;
  for (int i = 0; i < 5; ++i) finite = finite && std::isfinite(stats[i]) && std::fabs(stats[i]) < 1.0e300;
  if (flag_out) *flag_out = !finite || zrelerr > 10.0 * zeps;
  return 100.0 * zrelerr;
}

}  // extern "C"
