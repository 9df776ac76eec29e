/*
 * cloudsc2_drivers.c -- oracle (TEST INFRASTRUCTURE, see cloudsc2_oracle.h): plain-C restatement
 * of the reference's three block-loop drivers and their self tests:
 *   CLOUDSC_DRIVER     src/cloudsc2_nl/cloudsc_driver_mod.F90:22-125
 *   CLOUDSC_DRIVER_TL  src/cloudsc2_tl/cloudsc_driver_tl_mod.F90:21-31 (ERROR_NORM), 33-311
 *   CLOUDSC_DRIVER_AD  src/cloudsc2_ad/cloudsc_driver_ad_mod.F90:22-294
 * OpenMP over NPROMA blocks exactly as the reference (schedule(runtime) -> OMP_SCHEDULE).
 */
#include <float.h>
#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "cloudsc2_oracle.h"

#define NSTATE CLOUDSC2_NSTATE
#define NCLV CLOUDSC2_NCLV

int orc_max_threads(void) { return omp_get_max_threads(); }

static int nblocks_of(int ngptot, int nproma) {
  return ngptot / nproma + ((ngptot % nproma) < 1 ? (ngptot % nproma) : 1);
}

/* pointers to the per-block slabs the reference passes as array sections */
typedef struct blk {
  const double *pt, *pq, *pap, *paph, *plu, *plude, *pmfu, *pmfd, *psupsat, *pl, *pi;
  const double *cml_t, *cml_q, *cml_l, *cml_i;
  double *loc_t, *loc_q, *loc_l, *loc_i, *loc_last;
  double *pa, *pcovptot, *pfplsl, *pfplsn, *pfhpsl, *pfhpsn;
} blk;

static blk get_blk(const cloudsc2_fields *F, int nproma, int klev, int ibl) {
  const size_t n2 = (size_t)nproma * klev, n2h = (size_t)nproma * (klev + 1);
  blk b;
  b.pt = F->pt + n2 * ibl; b.pq = F->pq + n2 * ibl; b.pap = F->pap + n2 * ibl;
  b.paph = F->paph + n2h * ibl; b.plu = F->plu + n2 * ibl; b.plude = F->plude + n2 * ibl;
  b.pmfu = F->pmfu + n2 * ibl; b.pmfd = F->pmfd + n2 * ibl; b.psupsat = F->psupsat + n2 * ibl;
  b.pl = F->pclv + n2 * ((size_t)NCLV * ibl + 0);   /* PCLV(:,:,NCLDQL,IBL) */
  b.pi = F->pclv + n2 * ((size_t)NCLV * ibl + 1);   /* PCLV(:,:,NCLDQI,IBL) */
  const double *cml = F->b_cml + n2 * (size_t)NSTATE * ibl;
  double *loc = F->b_loc + n2 * (size_t)NSTATE * ibl;
  b.cml_t = cml; b.cml_q = cml + 2 * n2; b.cml_l = cml + 3 * n2; b.cml_i = cml + 4 * n2;
  b.loc_t = loc; b.loc_q = loc + 2 * n2; b.loc_l = loc + 3 * n2; b.loc_i = loc + 4 * n2;
  b.loc_last = loc + 7 * n2;                          /* %CLD(:,:,NCLV) */
  b.pa = F->pa + n2 * ibl; b.pcovptot = F->pcovptot + n2 * ibl;
  b.pfplsl = F->pfplsl + n2h * ibl; b.pfplsn = F->pfplsn + n2h * ibl;
  b.pfhpsl = F->pfhpsl + n2h * ibl; b.pfhpsn = F->pfhpsn + n2h * ibl;
  return b;
}

/* cloudsc_driver_mod.F90:82-111 */
int orc_driver_nl(const cloudsc2_params *P, const double *ceta, int numomp, int nproma, int klev,
                  int ngptot, double ptsphy, const cloudsc2_fields *F, double *elapsed_s) {
  const int ngpblks = nblocks_of(ngptot, nproma);
  const size_t n2 = (size_t)nproma * klev;
  int err = 0;
  double t0 = omp_get_wtime();
#pragma omp parallel num_threads(numomp)
  {
    double *zqsat = (double *)calloc(n2, sizeof(double));
#pragma omp for schedule(runtime)
    for (int ibl = 0; ibl < ngpblks; ++ibl) {
      const int jkglo = ibl * nproma + 1;
      int icend = ngptot - jkglo + 1;
      if (icend > nproma) icend = nproma;
      blk b = get_blk(F, nproma, klev, ibl);
      memset(b.pcovptot, 0, sizeof(double) * n2);     /* :87 */
      memset(b.loc_last, 0, sizeof(double) * n2);     /* :88 */
      orc_satur(P, 1, icend, nproma, klev, b.pap, b.pt, zqsat);   /* :91 */
      int rc = orc_cloudsc2(P, ceta, 1, icend, nproma, klev, ptsphy, b.paph, b.pap, b.pq, zqsat,
                            b.pt, b.pl, b.pi, b.plude, b.plu, b.pmfu, b.pmfd, b.loc_t, b.cml_t,
                            b.loc_q, b.cml_q, b.loc_l, b.cml_l, b.loc_i, b.cml_i, b.psupsat,
                            b.pa, b.pfplsl, b.pfplsn, b.pfhpsl, b.pfhpsn, b.pcovptot); /* :94 */
      if (rc) {
#pragma omp atomic write
        err = rc;
      }
    }
    free(zqsat);
  }
  if (elapsed_s) *elapsed_s = omp_get_wtime() - t0;
  return err;
}

/* ---------------------------------------------------------------------------------------- */
/* TL / AD drivers                                                                          */
/* ---------------------------------------------------------------------------------------- */

typedef struct scratch26 {   /* 16 input-like + 10 output-like (NPROMA,KLEV[+1]) arrays */
  orc_in16 in;
  orc_out10 out;
  double *base;
} scratch26;

static int scratch_alloc(scratch26 *s, int nproma, int klev) {
  const size_t n2 = (size_t)nproma * klev, n2h = (size_t)nproma * (klev + 1);
  const size_t tot = 15 * n2 + n2h + 6 * n2 + 4 * n2h;
  s->base = (double *)calloc(tot, sizeof(double));
  if (!s->base) return -2;
  double *p = s->base;
  s->in.paphp1 = p; p += n2h;
  s->in.papp1 = p; p += n2; s->in.pqm1 = p; p += n2; s->in.pqs = p; p += n2;
  s->in.ptm1 = p; p += n2; s->in.pl = p; p += n2; s->in.pi = p; p += n2;
  s->in.plude = p; p += n2; s->in.plu = p; p += n2; s->in.pmfu = p; p += n2;
  s->in.pmfd = p; p += n2; s->in.pgtent = p; p += n2; s->in.pgtenq = p; p += n2;
  s->in.pgtenl = p; p += n2; s->in.pgteni = p; p += n2; s->in.psupsat = p; p += n2;
  s->out.ptent = p; p += n2; s->out.ptenq = p; p += n2; s->out.ptenl = p; p += n2;
  s->out.pteni = p; p += n2; s->out.pclc = p; p += n2; s->out.pcovptot = p; p += n2;
  s->out.pfplsl = p; p += n2h; s->out.pfplsn = p; p += n2h;
  s->out.pfhpsl = p; p += n2h; s->out.pfhpsn = p; p += n2h;
  return 0;
}
static void scratch_free(scratch26 *s) { free(s->base); s->base = NULL; }

/* trajectory views of one block (in5 reads the caller's arrays, out5 writes TENDENCY_LOC etc.) */
static void traj_views(const blk *b, double *zqsat, orc_in16 *in5, orc_out10 *out5) {
  in5->paphp1 = (double *)b->paph; in5->papp1 = (double *)b->pap; in5->pqm1 = (double *)b->pq;
  in5->pqs = zqsat; in5->ptm1 = (double *)b->pt; in5->pl = (double *)b->pl;
  in5->pi = (double *)b->pi; in5->plude = (double *)b->plude; in5->plu = (double *)b->plu;
  in5->pmfu = (double *)b->pmfu; in5->pmfd = (double *)b->pmfd;
  in5->pgtent = (double *)b->cml_t; in5->pgtenq = (double *)b->cml_q;
  in5->pgtenl = (double *)b->cml_l; in5->pgteni = (double *)b->cml_i;
  in5->psupsat = (double *)b->psupsat;
  out5->ptent = b->loc_t; out5->ptenq = b->loc_q; out5->ptenl = b->loc_l; out5->pteni = b->loc_i;
  out5->pclc = b->pa; out5->pfplsl = b->pfplsl; out5->pfplsn = b->pfplsn;
  out5->pfhpsl = b->pfhpsl; out5->pfhpsn = b->pfhpsn; out5->pcovptot = b->pcovptot;
}

/* cloudsc_driver_tl_mod.F90:156-171 : increments = 0.01 * whole (NPROMA,NLEV) slab */
static void scale16(orc_in16 *d, const orc_in16 *x, double f, size_t n2, size_t n2h, int zero_sup) {
  for (size_t i = 0; i < n2h; ++i) d->paphp1[i] = x->paphp1[i] * f;
  for (size_t i = 0; i < n2; ++i) {
    d->papp1[i] = x->papp1[i] * f; d->pqm1[i] = x->pqm1[i] * f; d->pqs[i] = x->pqs[i] * f;
    d->ptm1[i] = x->ptm1[i] * f; d->pl[i] = x->pl[i] * f; d->pi[i] = x->pi[i] * f;
    d->plude[i] = x->plude[i] * f; d->plu[i] = x->plu[i] * f; d->pmfu[i] = x->pmfu[i] * f;
    d->pmfd[i] = x->pmfd[i] * f; d->pgtent[i] = x->pgtent[i] * f; d->pgtenq[i] = x->pgtenq[i] * f;
    d->pgtenl[i] = x->pgtenl[i] * f; d->pgteni[i] = x->pgteni[i] * f;
    d->psupsat[i] = zero_sup ? 0.0 : x->psupsat[i] * f;
  }
}
/* :200-215 : x5 = x + lambda*dx */
static void axpy16(orc_in16 *y, const orc_in16 *x, double lam, const orc_in16 *d, size_t n2,
                   size_t n2h) {
  for (size_t i = 0; i < n2h; ++i) y->paphp1[i] = x->paphp1[i] + lam * d->paphp1[i];
  for (size_t i = 0; i < n2; ++i) {
    y->papp1[i] = x->papp1[i] + lam * d->papp1[i]; y->pqm1[i] = x->pqm1[i] + lam * d->pqm1[i];
    y->pqs[i] = x->pqs[i] + lam * d->pqs[i]; y->ptm1[i] = x->ptm1[i] + lam * d->ptm1[i];
    y->pl[i] = x->pl[i] + lam * d->pl[i]; y->pi[i] = x->pi[i] + lam * d->pi[i];
    y->plude[i] = x->plude[i] + lam * d->plude[i]; y->plu[i] = x->plu[i] + lam * d->plu[i];
    y->pmfu[i] = x->pmfu[i] + lam * d->pmfu[i]; y->pmfd[i] = x->pmfd[i] + lam * d->pmfd[i];
    y->pgtent[i] = x->pgtent[i] + lam * d->pgtent[i];
    y->pgtenq[i] = x->pgtenq[i] + lam * d->pgtenq[i];
    y->pgtenl[i] = x->pgtenl[i] + lam * d->pgtenl[i];
    y->pgteni[i] = x->pgteni[i] + lam * d->pgteni[i];
    y->psupsat[i] = x->psupsat[i] + lam * d->psupsat[i];
  }
}

/* ERROR_NORM, cloudsc_driver_tl_mod.F90:21-31 ; Fortran SUM over (1:NLON,:) = JL fastest */
static void error_norm(int nlon, int nproma, int nlev, const double *field, const double *pert5,
                       const double *pert, double *znorm, double *zcount, double zlambda) {
  double s_tl = 0.0, s_d = 0.0;
  for (int jk = 0; jk < nlev; ++jk)
    for (int jl = 0; jl < nlon; ++jl) {
      size_t i = (size_t)jk * nproma + jl;
      s_tl += pert[i] * zlambda;
      s_d += field[i] - pert5[i];
    }
  if (fabs(s_tl) > DBL_EPSILON) {
    *zcount += 1.0;
    *znorm += fabs(s_d / s_tl);
  }
}

int orc_cloudsc2_call(const cloudsc2_params *P, const double *ceta, int icend, int nproma, int klev,
                      double ptsphy, const orc_in16 *x, const orc_out10 *y) {
  return orc_cloudsc2(P, ceta, 1, icend, nproma, klev, ptsphy, x->paphp1, x->papp1, x->pqm1,
                      x->pqs, x->ptm1, x->pl, x->pi, x->plude, x->plu, x->pmfu, x->pmfd, y->ptent,
                      x->pgtent, y->ptenq, x->pgtenq, y->ptenl, x->pgtenl, y->pteni, x->pgteni,
                      x->psupsat, y->pclc, y->pfplsl, y->pfplsn, y->pfhpsl, y->pfhpsn,
                      y->pcovptot);
}

/* cloudsc_driver_tl_mod.F90:126-254 */
int orc_driver_tl(const cloudsc2_params *P, const double *ceta, int numomp, int nproma, int klev,
                  int ngptot, double ptsphy, const cloudsc2_fields *F, double znormg[10],
                  double *ratios_blk, double *elapsed_s) {
  const int ngpblks = nblocks_of(ngptot, nproma);
  const size_t n2 = (size_t)nproma * klev, n2h = (size_t)nproma * (klev + 1);
  int err = 0;
  double zng[10];
  for (int i = 0; i < 10; ++i) zng[i] = 0.0;   /* :112 */
  double t0 = omp_get_wtime();
#pragma omp parallel num_threads(numomp)
  {
    double *zqsat = (double *)calloc(n2, sizeof(double));
    scratch26 d, x5;     /* d = increments (in) + TL outputs; x5 = perturbed state + NL outputs */
    int arc = scratch_alloc(&d, nproma, klev) | scratch_alloc(&x5, nproma, klev);
    double zloc[10];
    for (int i = 0; i < 10; ++i) zloc[i] = 0.0;
#pragma omp for schedule(runtime)
    for (int ibl = 0; ibl < ngpblks; ++ibl) {
      if (arc) { err = -2; continue; }
      const int jkglo = ibl * nproma + 1;
      int icend = ngptot - jkglo + 1;
      if (icend > nproma) icend = nproma;
      blk b = get_blk(F, nproma, klev, ibl);
      memset(b.pcovptot, 0, sizeof(double) * n2);
      memset(b.loc_last, 0, sizeof(double) * n2);
      orc_satur(P, 1, icend, nproma, klev, b.pap, b.pt, zqsat);              /* :135 */
      orc_in16 in5; orc_out10 out5;
      traj_views(&b, zqsat, &in5, &out5);
      int rc = orc_cloudsc2_call(P, ceta, icend, nproma, klev, ptsphy, &in5, &out5); /* :138 */
      scale16(&d.in, &in5, 0.01, n2, n2h, 0);                                 /* :156-171 */
      rc |= orc_cloudsc2tl(P, ceta, 1, icend, nproma, klev, ptsphy, &in5, &out5, &d.in, &d.out);
      for (int ilam = 1; ilam <= 10; ++ilam) {                                /* :197 */
        const double zlambda = pow(10.0, -(double)ilam);
        axpy16(&x5.in, &in5, zlambda, &d.in, n2, n2h);                        /* :200-215 */
        rc |= orc_cloudsc2_call(P, ceta, icend, nproma, klev, ptsphy, &x5.in, &x5.out);
        double zcount = 0.0, znorm = 0.0;                                     /* :233-244 */
        error_norm(icend, nproma, klev, out5.ptent, x5.out.ptent, d.out.ptent, &znorm, &zcount, zlambda);
        error_norm(icend, nproma, klev, out5.ptenq, x5.out.ptenq, d.out.ptenq, &znorm, &zcount, zlambda);
        error_norm(icend, nproma, klev, out5.ptenl, x5.out.ptenl, d.out.ptenl, &znorm, &zcount, zlambda);
        error_norm(icend, nproma, klev, out5.pteni, x5.out.pteni, d.out.pteni, &znorm, &zcount, zlambda);
        error_norm(icend, nproma, klev, out5.pclc, x5.out.pclc, d.out.pclc, &znorm, &zcount, zlambda);
        error_norm(icend, nproma, klev + 1, out5.pfplsl, x5.out.pfplsl, d.out.pfplsl, &znorm, &zcount, zlambda);
        error_norm(icend, nproma, klev + 1, out5.pfplsn, x5.out.pfplsn, d.out.pfplsn, &znorm, &zcount, zlambda);
        error_norm(icend, nproma, klev + 1, out5.pfhpsl, x5.out.pfhpsl, d.out.pfhpsl, &znorm, &zcount, zlambda);
        error_norm(icend, nproma, klev + 1, out5.pfhpsn, x5.out.pfhpsn, d.out.pfhpsn, &znorm, &zcount, zlambda);
        error_norm(icend, nproma, klev, out5.pcovptot, x5.out.pcovptot, d.out.pcovptot, &znorm, &zcount, zlambda);
        if (znorm == 0.0 || zcount == 0.0) {                                  /* :247-249 STOP */
          if (ratios_blk) ratios_blk[(size_t)ibl * 10 + ilam - 1] = NAN;
          rc |= 3;
        } else {
          const double r = znorm / zcount;
          if (ratios_blk) ratios_blk[(size_t)ibl * 10 + ilam - 1] = r;
          if (r > zloc[ilam - 1]) zloc[ilam - 1] = r;                         /* :251 */
        }
      }
      if (rc) {
#pragma omp atomic write
        err = rc;
      }
    }
#pragma omp critical
    for (int i = 0; i < 10; ++i)
      if (zloc[i] > zng[i]) zng[i] = zloc[i];    /* reduction(max:znormg) :125 */
    free(zqsat);
    if (!arc) { scratch_free(&d); scratch_free(&x5); }
  }
  if (elapsed_s) *elapsed_s = omp_get_wtime() - t0;
  for (int i = 0; i < 10; ++i) znormg[i] = zng[i];
  return err;
}

/* cloudsc_driver_tl_mod.F90:273-311 */
int orc_taylor_verdict(const double znormg_in[10], int *istart_out) {
  double z[11];
  int istart = 0;
  for (int ilam = 1; ilam <= 10; ++ilam) {
    z[ilam] = fabs(1.0 - znormg_in[ilam - 1]);                  /* :278 */
    if (istart == 0 && z[ilam] < 0.5) istart = ilam;            /* :280 */
  }
  if (istart_out) *istart_out = istart;
  if (istart == 0 || istart > 4) return -13;                    /* :284-285 */
  int itest = -10, inegat = 1;
  for (int ilam = istart; ilam <= 9; ++ilam) {                  /* :290-298 */
    int itempnegat = (z[ilam + 1] / z[ilam] < 1.0) ? 1 : 0;
    if (inegat > itempnegat) itest += 10;
    inegat = itempnegat;
  }
  if (itest == -10) itest = 11;                                 /* :299 */
  double zmin = z[istart];
  for (int ilam = istart; ilam <= 10; ++ilam) if (z[ilam] < zmin) zmin = z[ilam];
  if (zmin > 0.00001) itest += 7;                               /* :301 */
  if (zmin > 0.000001) itest += 5;                              /* :302 */
  return itest;                                                 /* passed iff <= 5 (:304) */
}

/* cloudsc_driver_ad_mod.F90:108-271 */
int orc_driver_ad(const cloudsc2_params *P, const double *ceta, int numomp, int nproma, int klev,
                  int ngptot, double ptsphy, const cloudsc2_fields *F, double *znormg_out,
                  double *norms_col, double *elapsed_s) {
  const int ngpblks = nblocks_of(ngptot, nproma);
  const size_t n2 = (size_t)nproma * klev, n2h = (size_t)nproma * (klev + 1);
  int err = 0;
  double znormg = 0.0;
  double t0 = omp_get_wtime();
#pragma omp parallel num_threads(numomp)
  {
    double *zqsat = (double *)calloc(n2, sizeof(double));
    scratch26 d, d0;
    int arc = scratch_alloc(&d, nproma, klev) | scratch_alloc(&d0, nproma, klev);
    double zloc = 0.0;
#pragma omp for schedule(runtime)
    for (int ibl = 0; ibl < ngpblks; ++ibl) {
      if (arc) { err = -2; continue; }
      const int jkglo = ibl * nproma + 1;
      int icend = ngptot - jkglo + 1;
      if (icend > nproma) icend = nproma;
      blk b = get_blk(F, nproma, klev, ibl);
      memset(b.pcovptot, 0, sizeof(double) * n2);
      memset(b.loc_last, 0, sizeof(double) * n2);
      orc_satur(P, 1, icend, nproma, klev, b.pap, b.pt, zqsat);              /* :117 */
      orc_in16 in5; orc_out10 out5;
      traj_views(&b, zqsat, &in5, &out5);
      scale16(&d.in, &in5, 0.01, n2, n2h, 1);                                 /* :124-139 */
      scale16(&d0.in, &in5, 0.01, n2, n2h, 1);                                /* :142-157 */
      int rc = orc_cloudsc2tl(P, ceta, 1, icend, nproma, klev, ptsphy, &in5, &out5, &d.in, &d.out);
      for (int jrof = 0; jrof < icend; ++jrof) {                              /* :184-195 */
        double s[10] = {0};
        for (int jk = 0; jk < klev; ++jk) {
          size_t i = (size_t)jk * nproma + jrof;
          s[0] += d.out.ptent[i] * d.out.ptent[i]; s[1] += d.out.ptenq[i] * d.out.ptenq[i];
          s[2] += d.out.ptenl[i] * d.out.ptenl[i]; s[3] += d.out.pteni[i] * d.out.pteni[i];
          s[4] += d.out.pclc[i] * d.out.pclc[i]; s[9] += d.out.pcovptot[i] * d.out.pcovptot[i];
        }
        for (int jk = 0; jk < klev + 1; ++jk) {
          size_t i = (size_t)jk * nproma + jrof;
          s[5] += d.out.pfplsl[i] * d.out.pfplsl[i]; s[6] += d.out.pfplsn[i] * d.out.pfplsn[i];
          s[7] += d.out.pfhpsl[i] * d.out.pfhpsl[i]; s[8] += d.out.pfhpsn[i] * d.out.pfhpsn[i];
        }
        double n1 = s[0];
        for (int k = 1; k < 10; ++k) n1 += s[k];
        d0.out.ptent[jrof] = n1;   /* stash ZNORM1 in unused scratch (d0.out is otherwise idle) */
      }
      memset(d.in.paphp1, 0, sizeof(double) * (15 * n2 + n2h));               /* :198-213 */
      rc |= orc_cloudsc2ad(P, ceta, 1, icend, nproma, klev, ptsphy, &in5, &out5, &d.in, &d.out);
      for (int jrof = 0; jrof < icend; ++jrof) {                              /* :240-264 */
        double s[16] = {0};
        for (int jk = 0; jk < klev + 1; ++jk) {
          size_t i = (size_t)jk * nproma + jrof;
          s[0] += d0.in.paphp1[i] * d.in.paphp1[i];
        }
        for (int jk = 0; jk < klev; ++jk) {
          size_t i = (size_t)jk * nproma + jrof;
          s[1] += d0.in.papp1[i] * d.in.papp1[i]; s[2] += d0.in.pqm1[i] * d.in.pqm1[i];
          s[3] += d0.in.pqs[i] * d.in.pqs[i]; s[4] += d0.in.ptm1[i] * d.in.ptm1[i];
          s[5] += d0.in.pl[i] * d.in.pl[i]; s[6] += d0.in.pi[i] * d.in.pi[i];
          s[7] += d0.in.plude[i] * d.in.plude[i]; s[8] += d0.in.plu[i] * d.in.plu[i];
          s[9] += d0.in.pmfu[i] * d.in.pmfu[i]; s[10] += d0.in.pmfd[i] * d.in.pmfd[i];
          s[11] += d0.in.pgtent[i] * d.in.pgtent[i]; s[12] += d0.in.pgtenq[i] * d.in.pgtenq[i];
          s[13] += d0.in.pgtenl[i] * d.in.pgtenl[i]; s[14] += d0.in.pgteni[i] * d.in.pgteni[i];
          s[15] += d0.in.psupsat[i] * d.in.psupsat[i];
        }
        double n2v = s[0];
        for (int k = 1; k < 16; ++k) n2v += s[k];
        const double n1 = d0.out.ptent[jrof];
        double n3;
        if (n2v == 0.0) n3 = fabs(n1 - n2v) / DBL_EPSILON;                    /* :260-264 */
        else n3 = fabs(n1 - n2v) / DBL_EPSILON / n2v;
        if (norms_col) {
          size_t g = (size_t)ibl * nproma + jrof;
          norms_col[3 * g] = n1; norms_col[3 * g + 1] = n2v; norms_col[3 * g + 2] = n3;
        }
        if (n3 > zloc) zloc = n3;                                             /* :267 */
      }
      if (rc) {
#pragma omp atomic write
        err = rc;
      }
    }
#pragma omp critical
    if (zloc > znormg) znormg = zloc;
    free(zqsat);
    if (!arc) { scratch_free(&d); scratch_free(&d0); }
  }
  if (elapsed_s) *elapsed_s = omp_get_wtime() - t0;
  *znormg_out = znormg;
  return err;
}

int orc_adjoint_verdict(double znormg) { return znormg < 10000.0 ? 1 : 0; } /* :289 */

/* Timing-only TL / AD block loops for the CPU baseline */
static int bench_tlad(int is_ad, const cloudsc2_params *P, const double *ceta, int numomp,
                      int nproma, int klev, int ngptot, double ptsphy, const cloudsc2_fields *F,
                      double *elapsed_s) {
  const int ngpblks = nblocks_of(ngptot, nproma);
  const size_t n2 = (size_t)nproma * klev, n2h = (size_t)nproma * (klev + 1);
  int err = 0;
  double t0 = omp_get_wtime();
#pragma omp parallel num_threads(numomp)
  {
    double *zqsat = (double *)calloc(n2, sizeof(double));
    scratch26 d;
    int arc = scratch_alloc(&d, nproma, klev);
#pragma omp for schedule(runtime)
    for (int ibl = 0; ibl < ngpblks; ++ibl) {
      if (arc) { err = -2; continue; }
      int icend = ngptot - ibl * nproma;
      if (icend > nproma) icend = nproma;
      blk b = get_blk(F, nproma, klev, ibl);
      orc_satur(P, 1, icend, nproma, klev, b.pap, b.pt, zqsat);
      orc_in16 in5; orc_out10 out5;
      traj_views(&b, zqsat, &in5, &out5);
      int rc;
      if (!is_ad) {
        scale16(&d.in, &in5, 0.01, n2, n2h, 0);
        rc = orc_cloudsc2tl(P, ceta, 1, icend, nproma, klev, ptsphy, &in5, &out5, &d.in, &d.out);
      } else {
        memset(d.base, 0, sizeof(double) * (15 * n2 + n2h));
        for (size_t i = 0; i < 6 * n2 + 4 * n2h; ++i) d.out.ptent[i] = 1.0e-6;  /* some y */
        rc = orc_cloudsc2ad(P, ceta, 1, icend, nproma, klev, ptsphy, &in5, &out5, &d.in, &d.out);
      }
      if (rc) {
#pragma omp atomic write
        err = rc;
      }
    }
    free(zqsat);
    if (!arc) scratch_free(&d);
  }
  if (elapsed_s) *elapsed_s = omp_get_wtime() - t0;
  return err;
}
int orc_bench_tl(const cloudsc2_params *P, const double *ceta, int numomp, int nproma, int klev,
                 int ngptot, double ptsphy, const cloudsc2_fields *F, double *elapsed_s) {
  return bench_tlad(0, P, ceta, numomp, nproma, klev, ngptot, ptsphy, F, elapsed_s);
}
int orc_bench_ad(const cloudsc2_params *P, const double *ceta, int numomp, int nproma, int klev,
                 int ngptot, double ptsphy, const cloudsc2_fields *F, double *elapsed_s) {
  return bench_tlad(1, P, ceta, numomp, nproma, klev, ngptot, ptsphy, F, elapsed_s);
}
