/*
 * cloudsc2_tl.c -- oracle (TEST INFRASTRUCTURE, see cloudsc2_oracle.h): plain-C restatement of
 *   CUADJTQSTL  reference src/cloudsc2_tl/cuadjtqstl.F90:333-405 (KCALL==0; phase select on PT5)
 *   CLOUDSC2TL  reference src/cloudsc2_tl/cloudsc2tl.F90:310-1111
 * "5"-suffixed names are the trajectory, unsuffixed names the perturbation, as in the reference.
 * The LLO2 evaporation block (:845-943) is statically dead and not restated (see header).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "cloudsc2_oracle.h"

#define IX(jl, jk) ((size_t)(jk) * (size_t)klon + (size_t)(jl))
#define A(x) x[IX(jl, jk)]
#define SQ(x) ((x) * (x))
#define CUBE(x) ((x) * (x) * (x))

static inline double dmin(double a, double b) { return a < b ? a : b; }

/* cuadjtqstl.F90: phase selection on the trajectory T (same as cuadjtqs.F90:118-130), then
 * KCALL==0 :333-405 (two textually identical sweeps) */
void orc_cuadjtqstl(const cloudsc2_params *P, int kidia, int kfdia, int klon, int kk,
                    const double *psp5, double *pt5, double *pq5,
                    const double *psp, double *pt, double *pq) {
  const double zqmax = 0.5;
  for (int jl = kidia - 1; jl < kfdia; ++jl) {
    double z3es, z4es, z5alcp, zaldcp;
    if (pt5[IX(jl, kk)] > P->rtt) {
      z3es = P->r3les; z4es = P->r4les; z5alcp = P->r5alvcp; zaldcp = P->ralvdcp;
    } else {
      z3es = P->r3ies; z4es = P->r4ies; z5alcp = P->r5alscp; zaldcp = P->ralsdcp;
    }
    double zqp = -psp[jl] / SQ(psp5[jl]);
    double zqp5 = 1.0 / psp5[jl];
    for (int it = 0; it < 2; ++it) {
      double ztarg = pt[IX(jl, kk)];
      double ztarg5 = pt5[IX(jl, kk)];
      double zfoeew5 = P->r2es * exp(z3es * (ztarg5 - P->rtt) / (ztarg5 - z4es));
      double zfoeew = z3es * (P->rtt - z4es) * ztarg * zfoeew5 / SQ(ztarg5 - z4es);
      double zqsat = zqp5 * zfoeew + zqp * zfoeew5;
      double zqsat5 = zqp5 * zfoeew5;
      if (zqsat5 > zqmax) { zqsat = 0.0; zqsat5 = zqmax; }
      double zcor = (P->retv * zqsat) / SQ(1.0 - P->retv * zqsat5);
      double zcor5 = 1.0 / (1.0 - P->retv * zqsat5);
      zqsat = zqsat5 * zcor + zqsat * zcor5;
      zqsat5 = zqsat5 * zcor5;
      double z2s = -2.0 * ztarg * z5alcp / CUBE(ztarg5 - z4es);
      double z2s5 = z5alcp / SQ(ztarg5 - z4es);
      double zcond1 = (pq[IX(jl, kk)] - zqsat) / (1.0 + zqsat5 * zcor5 * z2s5) -
                      (pq5[IX(jl, kk)] - zqsat5) *
                          (zqsat * zcor5 * z2s5 + zqsat5 * zcor * z2s5 + zqsat5 * zcor5 * z2s) /
                          SQ(1.0 + zqsat5 * zcor5 * z2s5);
      double zcond15 = (pq5[IX(jl, kk)] - zqsat5) / (1.0 + zqsat5 * zcor5 * z2s5);
      pt[IX(jl, kk)] = pt[IX(jl, kk)] + zaldcp * zcond1;
      pt5[IX(jl, kk)] = pt5[IX(jl, kk)] + zaldcp * zcond15;
      pq[IX(jl, kk)] = pq[IX(jl, kk)] - zcond1;
      pq5[IX(jl, kk)] = pq5[IX(jl, kk)] - zcond15;
    }
  }
}

int orc_cloudsc2tl(const cloudsc2_params *P, const double *ceta, int kidia, int kfdia, int klon,
                   int klev, double ptsphy, const orc_in16 *in5, const orc_out10 *out5,
                   const orc_in16 *in, const orc_out10 *out) {
  if (P->levapls2 || P->ldrain1d || !P->lphylin) return -1;
  const int lregcl = P->lregcl;
  const double zscal = 0.9;
  const size_t n2 = (size_t)klon * (size_t)klev;

  /* trajectory inputs / outputs */
  const double *paphp15 = in5->paphp1, *papp15 = in5->papp1, *pqm15 = in5->pqm1, *pqs5 = in5->pqs,
               *ptm15 = in5->ptm1, *pl5 = in5->pl, *pi5 = in5->pi, *plude5 = in5->plude,
               *plu5 = in5->plu, *pmfu5 = in5->pmfu, *pmfd5 = in5->pmfd, *pgtent5 = in5->pgtent,
               *pgtenq5 = in5->pgtenq, *pgtenl5 = in5->pgtenl, *pgteni5 = in5->pgteni,
               *psupsat5 = in5->psupsat;
  double *ptent5 = out5->ptent, *ptenq5 = out5->ptenq, *ptenl5 = out5->ptenl,
         *pteni5 = out5->pteni, *pclc5 = out5->pclc, *pfplsl5 = out5->pfplsl,
         *pfplsn5 = out5->pfplsn, *pfhpsl5 = out5->pfhpsl, *pfhpsn5 = out5->pfhpsn,
         *pcovptot5 = out5->pcovptot;
  /* perturbation inputs / outputs */
  const double *paphp1 = in->paphp1, *papp1 = in->papp1, *pqm1 = in->pqm1, *pqs = in->pqs,
               *ptm1 = in->ptm1, *pl = in->pl, *pi = in->pi, *plude = in->plude, *plu = in->plu,
               *pmfu = in->pmfu, *pmfd = in->pmfd, *pgtent = in->pgtent, *pgtenq = in->pgtenq,
               *pgtenl = in->pgtenl, *pgteni = in->pgteni, *psupsat = in->psupsat;
  double *ptent = out->ptent, *ptenq = out->ptenq, *ptenl = out->ptenl, *pteni = out->pteni,
         *pclc = out->pclc, *pfplsl = out->pfplsl, *pfplsn = out->pfplsn, *pfhpsl = out->pfhpsl,
         *pfhpsn = out->pfhpsn, *pcovptot = out->pcovptot;

  /* work arrays: 15 (KLON,KLEV) pairs + ZSCALM + (KLON) vectors */
  const int NW2 = 30, NW1 = 46;
  double *w = (double *)calloc((size_t)NW2 * n2 + (size_t)klev + (size_t)NW1 * klon, sizeof(double));
  if (!w) return -2;
  double *q = w;
#define W2(name) double *name = q; q += n2
  W2(ztp1); W2(ztp15); W2(zqp1); W2(zqp15); W2(zl); W2(zl5); W2(zi); W2(zi5);
  W2(zlude); W2(zlude5); W2(zqc); W2(zqc5); W2(zqlwc); W2(zqlwc5); W2(zqiwc); W2(zqiwc5);
  W2(zdp); W2(zdp5); W2(zlsdcp); W2(zlsdcp5); W2(zlfdcp); W2(zlfdcp5); W2(zlvdcp); W2(zlvdcp5);
  W2(zrfreeze); W2(zrfreeze5); W2(zcondl); W2(zcondl5); W2(zcondi); W2(zcondi5);
#undef W2
  double *zscalm = q; q += klev;
#define W1(name) double *name = q; q += klon
  W1(zrfl); W1(zrfl5); W1(zsfl); W1(zsfl5); W1(zrfln); W1(zrfln5); W1(zsfln); W1(zsfln5);
  W1(zgdp); W1(zgdp5); W1(zdqdt); W1(zdqdt5); W1(zdtdt); W1(zdtdt5); W1(zdldt); W1(zdldt5);
  W1(zdidt); W1(zdidt5); W1(zqcrit); W1(zqcrit5); W1(zcovpclr); W1(zcovpclr5); W1(zcovptot);
  W1(zcovptot5); W1(zdqsdtemp); W1(zdqsdtemp5); W1(zcorqs); W1(zcorqs5); W1(zqold); W1(zqold5);
  W1(zpp); W1(zpp5); W1(zdq); W1(zdq5); W1(zqlim); W1(zqlim5); W1(zqsat); W1(zqsat5);
  W1(zfoeew); W1(zfoeew5); W1(zfwat); W1(zfwat5); W1(ztrpaus);
#undef W1
  (void)zcorqs; (void)zcorqs5; (void)zqlim; (void)zqlim5; /* feed only the dead LLO2 block */
  const double zevapr = 0.0, zevapr5 = 0.0, zevaps = 0.0, zevaps5 = 0.0;

  /* :321-333 */
  const double zckcodtl = 2.0 * P->rkconv * ptsphy;
  const double zckcodti = 5.0 * P->rkconv * ptsphy;
  const double zckcodtla = zckcodtl / 100.0;
  const double zckcodtia = zckcodti / 100.0;
  const double zcons2 = 1.0 / (ptsphy * P->rg);
  const double zcons3 = P->rlvtt / P->rcpd;
  const double zmeltp2 = P->rtt + 2.0;
  const double zqtmst = 1.0 / ptsphy;
  const double zqmax = 0.5, zeps1 = 1.e-12, zeps2 = 1.e-10;

  /* :342-353 */
  for (int jk = 0; jk < klev; ++jk)
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      A(ztp1) = A(ptm1) + ptsphy * A(pgtent);
      A(ztp15) = A(ptm15) + ptsphy * A(pgtent5);
      A(zqp1) = A(pqm1) + ptsphy * A(pgtenq) + A(psupsat);
      A(zqp15) = A(pqm15) + ptsphy * A(pgtenq5) + A(psupsat5);
      A(zl) = A(pl) + ptsphy * A(pgtenl);
      A(zl5) = A(pl5) + ptsphy * A(pgtenl5);
      A(zi) = A(pi) + ptsphy * A(pgteni);
      A(zi5) = A(pi5) + ptsphy * A(pgteni5);
    }
  /* :355-378 */
  for (int jk = 0; jk < klev; ++jk) {
    zscalm[jk] = zscal * pow(fmax(ceta[jk] - 0.2, zeps1), 0.2);
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      A(zdp) = paphp1[IX(jl, jk + 1)] - paphp1[IX(jl, jk)];
      A(zdp5) = paphp15[IX(jl, jk + 1)] - paphp15[IX(jl, jk)];
      double zzz = -P->rcpd * P->rvtmp2 * A(zqp1) / SQ(P->rcpd + P->rcpd * P->rvtmp2 * A(zqp15));
      double zzz5 = 1.0 / (P->rcpd + P->rcpd * P->rvtmp2 * A(zqp15));
      A(zlfdcp) = P->rlmlt * zzz;   A(zlfdcp5) = P->rlmlt * zzz5;
      A(zlsdcp) = P->rlstt * zzz;   A(zlsdcp5) = P->rlstt * zzz5;
      A(zlvdcp) = P->rlvtt * zzz;   A(zlvdcp5) = P->rlvtt * zzz5;
    }
  }
  /* :388-411 (work arrays are calloc'ed) */
  for (int jk = 0; jk < klev; ++jk)
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      A(pclc) = 0.0; A(pclc5) = 0.0; A(pcovptot5) = 0.0; A(pcovptot) = 0.0;
    }
  /* :415-428 */
  for (int jl = kidia - 1; jl < kfdia; ++jl) {
    zrfl[jl] = 0.0; zrfl5[jl] = 0.0; zsfl[jl] = 0.0; zsfl5[jl] = 0.0;
    pfplsl[IX(jl, 0)] = 0.0; pfplsl5[IX(jl, 0)] = 0.0;
    pfplsn[IX(jl, 0)] = 0.0; pfplsn5[IX(jl, 0)] = 0.0;
    zcovptot[jl] = 0.0; zcovptot5[jl] = 0.0; zcovpclr[jl] = 0.0; zcovpclr5[jl] = 0.0;
  }
  /* :431-442 : tropopause from the trajectory only */
  for (int jl = kidia - 1; jl < kfdia; ++jl) ztrpaus[jl] = 0.1;
  for (int jk = 0; jk < klev - 1; ++jk)
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      int llo1 = ceta[jk] > 0.1 && ceta[jk] < 0.4 && ztp15[IX(jl, jk)] > ztp15[IX(jl, jk + 1)];
      if (llo1) ztrpaus[jl] = ceta[jk];
    }

  for (int jk = 0; jk < klev; ++jk) {   /* :455 */
    /* :459-539 */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      double z3es, z4es;
      double zoealfaw = 0.545 * 0.17 * A(ztp1) / SQ(cosh(0.17 * (A(ztp15) - P->rlptrc)));
      double zoealfaw5 = 0.545 * (tanh(0.17 * (A(ztp15) - P->rlptrc)) + 1.0);
      if (A(ztp15) < P->rtt) {
        zfwat[jl] = zoealfaw; zfwat5[jl] = zoealfaw5; z3es = P->r3ies; z4es = P->r4ies;
      } else {
        zfwat[jl] = 0.0; zfwat5[jl] = 1.0; z3es = P->r3les; z4es = P->r4les;
      }
      zfoeew5[jl] = P->r2es * exp(z3es * (A(ztp15) - P->rtt) / (A(ztp15) - z4es));
      zfoeew[jl] = z3es * (P->rtt - z4es) * A(ztp1) * zfoeew5[jl] / SQ(A(ztp15) - z4es);
      double zesdp = zfoeew[jl] / A(papp15) - A(papp1) * zfoeew5[jl] / SQ(A(papp15));
      double zesdp5 = zfoeew5[jl] / A(papp15);
      if (zesdp5 > zqmax) { zesdp = 0.0; zesdp5 = zqmax; }
      double zfacw = -2.0 * P->r5les * A(ztp1) / CUBE(A(ztp15) - P->r4les);
      double zfacw5 = P->r5les / SQ(A(ztp15) - P->r4les);
      double zfaci = -2.0 * P->r5ies * A(ztp1) / CUBE(A(ztp15) - P->r4ies);
      double zfaci5 = P->r5ies / SQ(A(ztp15) - P->r4ies);
      double zfac = zfwat5[jl] * zfacw + zfacw5 * zfwat[jl] + (1.0 - zfwat5[jl]) * zfaci -
                    zfaci5 * zfwat[jl];
      double zfac5 = zfwat5[jl] * zfacw5 + (1.0 - zfwat5[jl]) * zfaci5;
      double zcor = P->retv * zesdp / SQ(1.0 - P->retv * zesdp5);
      double zcor5 = 1.0 / (1.0 - P->retv * zesdp5);
      zdqsdtemp[jl] = zfac5 * zcor5 * A(pqs) + zfac5 * A(pqs5) * zcor + zcor5 * A(pqs5) * zfac;
      zdqsdtemp5[jl] = zfac5 * zcor5 * A(pqs5);
      zcorqs[jl] = zcons3 * zdqsdtemp[jl];
      zcorqs5[jl] = 1.0 + zcons3 * zdqsdtemp5[jl];
      /* :495-501 */
      if (A(zqp15) > A(pqs5)) { zqlim[jl] = A(pqs); zqlim5[jl] = A(pqs5); }
      else { zqlim[jl] = A(zqp1); zqlim5[jl] = A(zqp15); }
      /* :505-520 */
      double zeta3 = ztrpaus[jl];
      double zrh1 = 1.0;
      double zrh2 = 0.35 + 0.14 * SQ((zeta3 - 0.25) / 0.15) + 0.04 * dmin(zeta3 - 0.25, 0.0) / 0.15;
      double zrh3 = 1.0;
      double zdeta2 = 0.3;
      double zdeta1 = 0.09 + 0.16 * (0.4 - zeta3) / 0.3;
      double zcrh2 = 0.0;
      if (ceta[jk] < zeta3) zcrh2 = zrh3;
      else if (ceta[jk] >= zeta3 && ceta[jk] < (zeta3 + zdeta2))
        zcrh2 = zrh3 + (zrh2 - zrh3) * ((ceta[jk] - zeta3) / zdeta2);
      else if (ceta[jk] >= (zeta3 + zdeta2) && ceta[jk] < (1.0 - zdeta1)) zcrh2 = zrh2;
      else if (ceta[jk] >= (1.0 - zdeta1))
        zcrh2 = zrh1 + (zrh2 - zrh1) * sqrt((1.0 - ceta[jk]) / zdeta1);
      /* :522-534 */
      double zsupsat5, zsupsat;
      if (A(ztp15) < P->rtice) { zsupsat5 = 1.8 - 3.e-03 * A(ztp15); zsupsat = -3.e-03 * A(ztp1); }
      else { zsupsat5 = 1.0; zsupsat = 0.0; }
      zqsat5[jl] = A(pqs5) * zsupsat5;
      zqsat[jl] = A(pqs) * zsupsat5 + A(pqs5) * zsupsat;
      zqcrit5[jl] = zcrh2 * zqsat5[jl];
      zqcrit[jl] = zcrh2 * zqsat[jl];
    }

    /* :543-593 */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      double zqt = A(zqp1) + A(zl) + A(zi);
      double zqt5 = A(zqp15) + A(zl5) + A(zi5);
      if (zqt5 <= zqcrit5[jl]) {
        A(pclc) = 0.0; A(pclc5) = 0.0; A(zqc) = 0.0; A(zqc5) = 0.0;
      } else if (zqt5 >= zqsat5[jl]) {
        A(pclc) = 0.0; A(pclc5) = 1.0;
        A(zqc) = (1.0 - zscalm[jk]) * (zqsat[jl] - zqcrit[jl]);
        A(zqc5) = (1.0 - zscalm[jk]) * (zqsat5[jl] - zqcrit5[jl]);
      } else {
        double zqpd = zqsat[jl] - zqt, zqpd5 = zqsat5[jl] - zqt5;
        double zqcd = zqsat[jl] - zqcrit[jl], zqcd5 = zqsat5[jl] - zqcrit5[jl];
        double zsqrt5 = sqrt(zqpd5 / (zqcd5 - zscalm[jk] * (zqt5 - zqcrit5[jl])));
        A(pclc5) = 1.0 - zsqrt5;
        A(pclc) = -(0.5 / zsqrt5) *
                  (zqpd * (zqcd5 - zscalm[jk] * (zqt5 - zqcrit5[jl])) -
                   zqpd5 * (zqcd - zscalm[jk] * (zqt - zqcrit[jl]))) /
                  SQ(zqcd5 - zscalm[jk] * (zqt5 - zqcrit5[jl]));
        if (lregcl) {   /* :575-580 */
          double zrat = zqpd5 / zqcd5;
          double zyyy = dmin(0.3, 3.5 * sqrt(zrat * CUBE(1.0 - zscalm[jk] * (1.0 - zrat))) /
                                      (1.0 - zscalm[jk]));
          A(pclc) = zyyy * A(pclc);
        }
        A(zqc) = (zscalm[jk] * zqpd + (1.0 - zscalm[jk]) * zqcd) * SQ(A(pclc5)) +
                 (zscalm[jk] * zqpd5 + (1.0 - zscalm[jk]) * zqcd5) * 2.0 * A(pclc5) * A(pclc);
        A(zqc5) = (zscalm[jk] * zqpd5 + (1.0 - zscalm[jk]) * zqcd5) * SQ(A(pclc5));
      }
    }

    /* :597-628 convective component */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      zgdp[jl] = -P->rg * (paphp1[IX(jl, jk + 1)] - paphp1[IX(jl, jk)]) /
                 SQ(paphp15[IX(jl, jk + 1)] - paphp15[IX(jl, jk)]);
      zgdp5[jl] = P->rg / (paphp15[IX(jl, jk + 1)] - paphp15[IX(jl, jk)]);
      A(zlude) = ptsphy * zgdp5[jl] * A(plude) + ptsphy * A(plude5) * zgdp[jl];
      A(zlude5) = A(plude5) * ptsphy * zgdp5[jl];
      int llo1;
      if (jk < klev - 1) llo1 = A(zlude5) >= P->rlmin && plu5[IX(jl, jk + 1)] >= zeps2;
      else llo1 = 0;
      if (llo1) {
        const double plu5n = plu5[IX(jl, jk + 1)], plun = plu[IX(jl, jk + 1)];
        A(pclc) = A(pclc) - A(pclc) * (1.0 - exp(-A(zlude5) / plu5n)) +
                  ((1.0 - A(pclc5)) / plu5n) * exp(-A(zlude5) / plu5n) * A(zlude) -
                  ((1.0 - A(pclc5)) * A(zlude5) / SQ(plu5n)) * exp(-A(zlude5) / plu5n) * plun;
        A(pclc5) = A(pclc5) + (1.0 - A(pclc5)) * (1.0 - exp(-A(zlude5) / plu5n));
        A(zqc) = A(zqc) + A(zlude);
        A(zqc5) = A(zqc5) + A(zlude5);
      }
    }

    /* :632-669 compensating subsidence */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      double zfac1 = 1.0 / (P->rd * A(ztp15));
      double zrho = (A(papp1) - A(ztp1) * A(papp15) / A(ztp15)) * zfac1;
      double zrho5 = A(papp15) * zfac1;
      double zfac2 = 1.0 / (A(papp15) - P->retv * zfoeew5[jl]);
      double zrodqsdp = (-zrho * A(pqs5) - zrho5 * A(pqs) +
                         zrho5 * A(pqs5) * (A(papp1) - P->retv * zfoeew[jl]) * zfac2) * zfac2;
      double zrodqsdp5 = -zrho5 * A(pqs5) * zfac2;
      double zldcp = zfwat[jl] * A(zlvdcp5) + zfwat5[jl] * A(zlvdcp) +
                     (1.0 - zfwat5[jl]) * A(zlsdcp) - zfwat[jl] * A(zlsdcp5);
      double zldcp5 = zfwat5[jl] * A(zlvdcp5) + (1.0 - zfwat5[jl]) * A(zlsdcp5);
      double zfac3 = 1.0 / (1.0 + zldcp5 * zdqsdtemp5[jl]);
      double dtdzmo5 = P->rg * (1.0 / P->rcpd - zldcp5 * zrodqsdp5) * zfac3;
      double dtdzmo = -(P->rg * (zldcp * zrodqsdp5 + zldcp5 * zrodqsdp) +
                        dtdzmo5 * (zldcp5 * zdqsdtemp[jl] + zldcp * zdqsdtemp5[jl])) * zfac3;
      double zdqsdz = zdqsdtemp5[jl] * dtdzmo + zdqsdtemp[jl] * dtdzmo5 - P->rg * zrodqsdp;
      double zdqsdz5 = zdqsdtemp5[jl] * dtdzmo5 - P->rg * zrodqsdp5;
      double zfac4 = 1.0 / zrho5;
      int llo3 = (zdqsdz5 * (A(pmfu5) + A(pmfd5)) * ptsphy * zfac4 < A(zqc5));
      double zdqc, zdqc5;
      if (llo3) {
        zdqc5 = zdqsdz5 * (A(pmfu5) + A(pmfd5)) * ptsphy * zfac4;
        zdqc = (ptsphy * (zdqsdz * (A(pmfu5) + A(pmfd5)) + zdqsdz5 * (A(pmfu) + A(pmfd))) -
                zdqc5 * zrho) * zfac4;
        if (lregcl) zdqc = zdqc * 0.1;   /* :657 */
      } else {
        zdqc5 = A(zqc5);
        zdqc = A(zqc);
      }
      A(zqc) = A(zqc) - zdqc;
      A(zqc5) = A(zqc5) - zdqc5;
    }

    /* :673-685 */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      A(zqlwc) = A(zqc) * zfwat5[jl] + A(zqc5) * zfwat[jl];
      A(zqlwc5) = A(zqc5) * zfwat5[jl];
      A(zqiwc) = A(zqc) * (1.0 - zfwat5[jl]) - A(zqc5) * zfwat[jl];
      A(zqiwc5) = A(zqc5) * (1.0 - zfwat5[jl]);
      A(zcondl) = (A(zqlwc) - A(zl)) * zqtmst;
      A(zcondl5) = (A(zqlwc5) - A(zl5)) * zqtmst;
      A(zcondi) = (A(zqiwc) - A(zi)) * zqtmst;
      A(zcondi5) = (A(zqiwc5) - A(zi5)) * zqtmst;
    }

    /* :690-701 precipitation overlap */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      if (A(pclc5) > zcovptot5[jl]) { zcovptot[jl] = A(pclc); zcovptot5[jl] = A(pclc5); }
      zcovpclr[jl] = zcovptot[jl] - A(pclc);
      zcovpclr5[jl] = zcovptot5[jl] - A(pclc5);
      if (zcovpclr5[jl] < 0.0) { zcovpclr[jl] = 0.0; zcovpclr5[jl] = 0.0; }
    }

    /* :707-738 melting of incoming snow */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      if (zsfl5[jl] != 0.0) {
        double zcons = zcons2 * (A(zdp) * A(zlfdcp5) - A(zdp5) * A(zlfdcp)) / SQ(A(zlfdcp5));
        double zcons5 = zcons2 * A(zdp5) / A(zlfdcp5);
        double zz2s, zz2s5, zsnmlt, zsnmlt5;
        if ((A(ztp15) - zmeltp2) > 0.0) {
          zz2s = zcons5 * A(ztp1) + zcons * (A(ztp15) - zmeltp2);
          zz2s5 = zcons5 * (A(ztp15) - zmeltp2);
        } else { zz2s = 0.0; zz2s5 = 0.0; }
        if (zsfl5[jl] <= zz2s5) { zsnmlt = zsfl[jl]; zsnmlt5 = zsfl5[jl]; }
        else { zsnmlt = zz2s; zsnmlt5 = zz2s5; }
        zrfln[jl] = zrfl[jl] + zsnmlt;    zrfln5[jl] = zrfl5[jl] + zsnmlt5;
        zsfln[jl] = zsfl[jl] - zsnmlt;    zsfln5[jl] = zsfl5[jl] - zsnmlt5;
        A(ztp1) = A(ztp1) - (zsnmlt * zcons5 - zcons * zsnmlt5) / SQ(zcons5);
        A(ztp15) = A(ztp15) - zsnmlt5 / zcons5;
      } else {
        zrfln[jl] = zrfl[jl]; zrfln5[jl] = zrfl5[jl];
        zsfln[jl] = zsfl[jl]; zsfln5[jl] = zsfl5[jl];
      }
    }

    /* :742-843 autoconversion and new precipitation */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      double zprr, zprr5, zprs, zprs5;
      if (A(pclc5) > zeps2) {   /* :743-779 liquid */
        double zlcrit = P->rclcrit * 2.0;
        double zcldl = A(zqlwc) / A(pclc5) - A(zqlwc5) * A(pclc) / SQ(A(pclc5));
        double zcldl5 = A(zqlwc5) / A(pclc5);
        double zexp35 = exp(-SQ(zcldl5 / zlcrit));
        double zd5 = zckcodtl * (1.0 - zexp35);
        double zexpdl5 = exp(-zd5);
        double zd;
        if (lregcl) zd = (2.0 * zckcodtla / SQ(zlcrit)) * exp(-SQ(zcldl5 / zlcrit)) * zcldl5 * zcldl;
        else zd = (2.0 * zckcodtl / SQ(zlcrit)) * exp(-SQ(zcldl5 / zlcrit)) * zcldl5 * zcldl;
        double zlnew = zcldl5 * zexpdl5 * A(pclc) + A(pclc5) * zexpdl5 * zcldl -
                       A(pclc5) * zcldl5 * zexpdl5 * zd;
        double zlnew5 = A(pclc5) * zcldl5 * zexpdl5;
        zprr = A(zqlwc) - zlnew;       zprr5 = A(zqlwc5) - zlnew5;
        A(zqlwc) = A(zqlwc) - zprr;    A(zqlwc5) = A(zqlwc5) - zprr5;
      } else { zprr = 0.0; zprr5 = 0.0; }
      if (A(pclc5) > zeps2) {   /* :783-819 ice */
        double zlcrit = P->rclcrit * 2.0;
        double zcldi = A(zqiwc) / A(pclc5) - A(zqiwc5) * A(pclc) / SQ(A(pclc5));
        double zcldi5 = A(zqiwc5) / A(pclc5);
        double zexp15 = exp(0.025 * (A(ztp15) - P->rtt));
        double zexp25 = exp(-SQ(zcldi5 / zlcrit));
        double zd5 = zckcodti * zexp15 * (1.0 - zexp25);
        double zexpdi5 = exp(-zd5);
        double zd;
        if (lregcl)
          zd = zckcodtia * zexp15 *
               (zexp25 * (2.0 * zcldi5 * zcldi / SQ(zlcrit) - 0.025 * A(ztp1)) + 0.025 * A(ztp1));
        else
          zd = zckcodti * zexp15 *
               (zexp25 * (2.0 * zcldi5 * zcldi / SQ(zlcrit) - 0.025 * A(ztp1)) + 0.025 * A(ztp1));
        double zinew = zcldi5 * zexpdi5 * A(pclc) + A(pclc5) * zexpdi5 * zcldi -
                       A(pclc5) * zcldi5 * zexpdi5 * zd;
        double zinew5 = A(pclc5) * zcldi5 * zexpdi5;
        zprs = A(zqiwc) - zinew;       zprs5 = A(zqiwc5) - zinew5;
        A(zqiwc) = A(zqiwc) - zprs;    A(zqiwc5) = A(zqiwc5) - zprs5;
      } else { zprs = 0.0; zprs5 = 0.0; }
      /* :823-843 */
      double zdr = zcons2 * (A(zdp5) * (zprr + zprs) + A(zdp) * (zprr5 + zprs5));
      double zdr5 = zcons2 * A(zdp5) * (zprr5 + zprs5);
      double zfwatr5, zfwatr;
      if (A(ztp15) < P->rtt) {
        A(zrfreeze5) = zcons2 * A(zdp5) * zprr5;
        A(zrfreeze) = zcons2 * (A(zdp) * zprr5 + A(zdp5) * zprr);
        zfwatr5 = 0.0; zfwatr = 0.0;
      } else { zfwatr5 = 1.0; zfwatr = 0.0; }
      double zrn = zfwatr5 * zdr + zdr5 * zfwatr;
      double zrn5 = zfwatr5 * zdr5;
      double zsn = -zdr5 * zfwatr + (1.0 - zfwatr5) * zdr;
      double zsn5 = (1.0 - zfwatr5) * zdr5;
      zrfln[jl] = zrfln[jl] + zrn;   zrfln5[jl] = zrfln5[jl] + zrn5;
      zsfln[jl] = zsfln[jl] + zsn;   zsfln5[jl] = zsfln5[jl] + zsn5;
      /* :845-943 precipitation evaporation: LLO2 statically false */
    }

    /* :949-989 incrementation of T and q */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      zdqdt[jl] = -(A(zcondl) + A(zcondi)) + (A(plude) + zevapr + zevaps) * zgdp5[jl] +
                  (A(plude5) + zevapr5 + zevaps5) * zgdp[jl];
      zdqdt5[jl] = -(A(zcondl5) + A(zcondi5)) + (A(plude5) + zevapr5 + zevaps5) * zgdp5[jl];
      zdtdt[jl] = A(zlvdcp) * A(zcondl5) + A(zlsdcp) * A(zcondi5) + A(zlvdcp5) * A(zcondl) +
                  A(zlsdcp5) * A(zcondi) -
                  (A(zlvdcp) * zevapr5 + A(zlsdcp) * zevaps5 + A(zlvdcp5) * zevapr +
                   A(zlsdcp5) * zevaps +
                   A(plude) * (zfwat5[jl] * A(zlvdcp5) + (1.0 - zfwat5[jl]) * A(zlsdcp5)) +
                   A(plude5) * (zfwat[jl] * (A(zlvdcp5) - A(zlsdcp5)) +
                                (zfwat5[jl] * A(zlvdcp) + (1.0 - zfwat5[jl]) * A(zlsdcp))) -
                   (A(zlsdcp) - A(zlvdcp)) * A(zrfreeze5) -
                   (A(zlsdcp5) - A(zlvdcp5)) * A(zrfreeze)) * zgdp5[jl] -
                  (A(zlvdcp5) * zevapr5 + A(zlsdcp5) * zevaps5 +
                   A(plude5) * (zfwat5[jl] * A(zlvdcp5) + (1.0 - zfwat5[jl]) * A(zlsdcp5)) -
                   (A(zlsdcp5) - A(zlvdcp5)) * A(zrfreeze5)) * zgdp[jl];
      zdtdt5[jl] = A(zlvdcp5) * A(zcondl5) + A(zlsdcp5) * A(zcondi5) -
                   (A(zlvdcp5) * zevapr5 + A(zlsdcp5) * zevaps5 +
                    A(plude5) * (zfwat5[jl] * A(zlvdcp5) + (1.0 - zfwat5[jl]) * A(zlsdcp5)) -
                    (A(zlsdcp5) - A(zlvdcp5)) * A(zrfreeze5)) * zgdp5[jl];
      A(ztp1) = A(ztp1) + ptsphy * zdtdt[jl];
      A(ztp15) = A(ztp15) + ptsphy * zdtdt5[jl];
      A(zqp1) = A(zqp1) + ptsphy * zdqdt[jl];
      A(zqp15) = A(zqp15) + ptsphy * zdqdt5[jl];
      zpp[jl] = A(papp1);   zpp5[jl] = A(papp15);
      zqold[jl] = A(zqp1);  zqold5[jl] = A(zqp15);
    }

    /* :993-997 */
    orc_cuadjtqstl(P, kidia, kfdia, klon, jk, zpp5, ztp15, zqp15, zpp, ztp1, zqp1);

    /* :999-1046 */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      if ((zqold5[jl] - A(zqp15)) >= 0.0) {
        zdq5[jl] = zqold5[jl] - A(zqp15);
        zdq[jl] = zqold[jl] - A(zqp1);
        if (lregcl) zdq[jl] = zdq[jl] * 0.7;   /* :1004-1006 */
      } else { zdq[jl] = 0.0; zdq5[jl] = 0.0; }
      double zdr2 = zcons2 * (A(zdp5) * zdq[jl] + zdq5[jl] * A(zdp));
      double zdr25 = zcons2 * A(zdp5) * zdq5[jl];
      double zrfreeze25, zrfreeze2, zfwatr5, zfwatr;
      if (A(ztp15) < P->rtt) {
        zrfreeze25 = zfwat5[jl] * zdr25;
        zrfreeze2 = zfwat[jl] * zdr25 + zfwat5[jl] * zdr2;
        zfwatr5 = 0.0; zfwatr = 0.0;
      } else {
        zrfreeze25 = 0.0; zrfreeze2 = 0.0; zfwatr5 = 1.0; zfwatr = 0.0;
      }
      double zrn = zfwatr5 * zdr2 + zdr25 * zfwatr;
      double zrn5 = zfwatr5 * zdr25;
      double zsn = (1.0 - zfwatr5) * zdr2 - zdr25 * zfwatr;
      double zsn5 = (1.0 - zfwatr5) * zdr25;
      A(zcondl) = A(zcondl) + (zfwatr5 * zdq[jl] + zfwatr * zdq5[jl]) * zqtmst;
      A(zcondl5) = A(zcondl5) + zfwatr5 * zdq5[jl] * zqtmst;
      A(zcondi) = A(zcondi) + ((1.0 - zfwatr5) * zdq[jl] - zfwatr * zdq5[jl]) * zqtmst;
      A(zcondi5) = A(zcondi5) + (1.0 - zfwatr5) * zdq5[jl] * zqtmst;
      zrfln[jl] = zrfln[jl] + zrn;   zrfln5[jl] = zrfln5[jl] + zrn5;
      zsfln[jl] = zsfln[jl] + zsn;   zsfln5[jl] = zsfln5[jl] + zsn5;
      A(zrfreeze5) = A(zrfreeze5) + zrfreeze25;
      A(zrfreeze) = A(zrfreeze) + zrfreeze2;
    }

    /* :1048-1096 */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      zdqdt[jl] = -(A(zcondl) + A(zcondi)) + (A(plude) + zevapr + zevaps) * zgdp5[jl] +
                  (A(plude5) + zevapr5 + zevaps5) * zgdp[jl];
      zdqdt5[jl] = -(A(zcondl5) + A(zcondi5)) + (A(plude5) + zevapr5 + zevaps5) * zgdp5[jl];
      zdtdt[jl] = A(zlvdcp) * A(zcondl5) + A(zlsdcp) * A(zcondi5) + A(zlvdcp5) * A(zcondl) +
                  A(zlsdcp5) * A(zcondi) -
                  (A(zlvdcp) * zevapr5 + A(zlsdcp) * zevaps5 + A(zlvdcp5) * zevapr +
                   A(zlsdcp5) * zevaps +
                   A(plude) * (zfwat5[jl] * A(zlvdcp5) + (1.0 - zfwat5[jl]) * A(zlsdcp5)) +
                   A(plude5) * (zfwat[jl] * (A(zlvdcp5) - A(zlsdcp5)) +
                                (zfwat5[jl] * A(zlvdcp) + (1.0 - zfwat5[jl]) * A(zlsdcp))) -
                   (A(zlsdcp) - A(zlvdcp)) * A(zrfreeze5) -
                   (A(zlsdcp5) - A(zlvdcp5)) * A(zrfreeze)) * zgdp5[jl] -
                  (A(zlvdcp5) * zevapr5 + A(zlsdcp5) * zevaps5 +
                   A(plude5) * (zfwat5[jl] * A(zlvdcp5) + (1.0 - zfwat5[jl]) * A(zlsdcp5)) -
                   (A(zlsdcp5) - A(zlvdcp5)) * A(zrfreeze5)) * zgdp[jl];
      zdtdt5[jl] = A(zlvdcp5) * A(zcondl5) + A(zlsdcp5) * A(zcondi5) -
                   (A(zlvdcp5) * zevapr5 + A(zlsdcp5) * zevaps5 +
                    A(plude5) * (zfwat5[jl] * A(zlvdcp5) + (1.0 - zfwat5[jl]) * A(zlsdcp5)) -
                    (A(zlsdcp5) - A(zlvdcp5)) * A(zrfreeze5)) * zgdp5[jl];
      zdldt[jl] = (A(zqlwc) - A(zl)) * zqtmst;     zdldt5[jl] = (A(zqlwc5) - A(zl5)) * zqtmst;
      zdidt[jl] = (A(zqiwc) - A(zi)) * zqtmst;     zdidt5[jl] = (A(zqiwc5) - A(zi5)) * zqtmst;
      A(ptenq) = zdqdt[jl];   A(ptenq5) = zdqdt5[jl];
      A(ptent) = zdtdt[jl];   A(ptent5) = zdtdt5[jl];
      A(ptenl) = zdldt[jl];   A(ptenl5) = zdldt5[jl];
      A(pteni) = zdidt[jl];   A(pteni5) = zdidt5[jl];
      pfplsl[IX(jl, jk + 1)] = zrfln[jl];   pfplsl5[IX(jl, jk + 1)] = zrfln5[jl];
      pfplsn[IX(jl, jk + 1)] = zsfln[jl];   pfplsn5[IX(jl, jk + 1)] = zsfln5[jl];
    }
    /* :1098-1103 */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      zrfl[jl] = zrfln[jl];   zrfl5[jl] = zrfln5[jl];
      zsfl[jl] = zsfln[jl];   zsfl5[jl] = zsfln5[jl];
    }
  } /* jk */

  /* :1108-1115 enthalpy fluxes */
  for (int jk = 0; jk < klev + 1; ++jk)
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      A(pfhpsl) = -A(pfplsl) * P->rlvtt;   A(pfhpsl5) = -A(pfplsl5) * P->rlvtt;
      A(pfhpsn) = -A(pfplsn) * P->rlstt;   A(pfhpsn5) = -A(pfplsn5) * P->rlstt;
    }

  free(w);
  return 0;
}
