/*
 * cloudsc2_ad.c -- oracle (TEST INFRASTRUCTURE, see cloudsc2_oracle.h): plain-C restatement of
 *   CUADJTQSAD  reference src/cloudsc2_ad/cuadjtqsad.F90:150-162 (phase), 314-367 (trajectory,
 *               KCALL==0), 542-641 (adjoint, KCALL==0)
 *   CLOUDSC2AD  reference src/cloudsc2_ad/cloudsc2ad.F90:340-1740: forward trajectory sweep that
 *               stores every intermediate in (KLON,KLEV) arrays (:364-866) exactly like the
 *               reference (this is the memory-hungry formulation the GPU kernel replaces by
 *               recomputation), adjoint initialisation (:877-921), reverse sweep (:934-1668),
 *               epilogue (:1677-1740).
 * "5"-suffixed = trajectory, unsuffixed = adjoint variable.  The LLO2 evaporation blocks
 * (:724-773 forward, :1152-1267 adjoint) are statically dead and not restated; the adjoint
 * accumulators ZEVAPR/ZEVAPS, ZCOVPCLR, ZCOVPTOT, ZCORQS, ZQLIM, ZDTGDP are only consumed inside
 * that block (they stay identically zero / unread) and are dropped with it.
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

void orc_cuadjtqsad(const cloudsc2_params *P, int kidia, int kfdia, int klon, int kk,
                    const double *psp5, double *pt5, double *pq5,
                    double *psp, double *pt, double *pq) {
  const double zqmax = 0.5;
  for (int jl = kidia - 1; jl < kfdia; ++jl) {
    /* :150-162 phase selection on the incoming (pre-adjustment) trajectory T */
    double z3es, z4es, z5alcp, zaldcp;
    if (pt5[IX(jl, kk)] > P->rtt) {
      z3es = P->r3les; z4es = P->r4les; z5alcp = P->r5alvcp; zaldcp = P->ralvdcp;
    } else {
      z3es = P->r3ies; z4es = P->r4ies; z5alcp = P->r5alscp; zaldcp = P->ralsdcp;
    }
    /* :314-367 trajectory of the two iterations; B = first, A = second */
    double zqp5 = 1.0 / psp5[jl];
    double ztarg5 = pt5[IX(jl, kk)];
    double zfoeew5 = P->r2es * exp(z3es * (ztarg5 - P->rtt) / (ztarg5 - z4es));
    const double zfoeew55b = zfoeew5;
    double zqsat5 = zqp5 * zfoeew5;
    const int lltest2 = zqsat5 > zqmax;
    if (zqsat5 > zqmax) zqsat5 = zqmax;
    double zcor5 = 1.0 / (1.0 - P->retv * zqsat5);
    const double zqsat55d = zqsat5;
    zqsat5 = zqsat5 * zcor5;
    const double ztarg55b = ztarg5;
    double z2s5 = z5alcp / SQ(ztarg5 - z4es);
    const double zqsat55b = zqsat5, zcor55b = zcor5, z2s55b = z2s5, zq55b = pq5[IX(jl, kk)];
    double zcond15 = (pq5[IX(jl, kk)] - zqsat5) / (1.0 + zqsat5 * zcor5 * z2s5);
    pt5[IX(jl, kk)] = pt5[IX(jl, kk)] + zaldcp * zcond15;
    pq5[IX(jl, kk)] = pq5[IX(jl, kk)] - zcond15;

    ztarg5 = pt5[IX(jl, kk)];
    zfoeew5 = P->r2es * exp(z3es * (ztarg5 - P->rtt) / (ztarg5 - z4es));
    const double zfoeew55a = zfoeew5;
    zqsat5 = zqp5 * zfoeew5;
    const int lltest1 = zqsat5 > zqmax;
    if (zqsat5 > zqmax) zqsat5 = zqmax;
    zcor5 = 1.0 / (1.0 - P->retv * zqsat5);
    const double zqsat55c = zqsat5;
    zqsat5 = zqsat5 * zcor5;
    const double ztarg55a = ztarg5;
    z2s5 = z5alcp / SQ(ztarg5 - z4es);
    const double zqsat55a = zqsat5, zcor55a = zcor5, z2s55a = z2s5, zq55a = pq5[IX(jl, kk)];
    zcond15 = (pq5[IX(jl, kk)] - zqsat5) / (1.0 + zqsat5 * zcor5 * z2s5);
    pt5[IX(jl, kk)] = pt5[IX(jl, kk)] + zaldcp * zcond15;
    pq5[IX(jl, kk)] = pq5[IX(jl, kk)] - zcond15;

    /* adjoint locals start at zero (cuadjtqsad.F90 initialisation section) */
    double zcond1 = 0.0, zqsat = 0.0, zcor = 0.0, z2s = 0.0, ztarg = 0.0, zfoeew = 0.0, zqp = 0.0;

    /* :546-593 adjoint of the second iteration */
    zcond1 = zcond1 - pq[IX(jl, kk)];
    zcond1 = zcond1 + zaldcp * pt[IX(jl, kk)];
    zqsat5 = zqsat55a; zcor5 = zcor55a; z2s5 = z2s55a;
    pq[IX(jl, kk)] = pq[IX(jl, kk)] + zcond1 / (1.0 + zqsat5 * zcor5 * z2s5);
    zqsat = zqsat - zcond1 / (1.0 + zqsat5 * zcor5 * z2s5);
    zqsat = zqsat - zcond1 * (zq55a - zqsat5) * zcor5 * z2s5 / SQ(1.0 + zqsat5 * zcor5 * z2s5);
    zcor = zcor - zcond1 * (zq55a - zqsat5) * zqsat5 * z2s5 / SQ(1.0 + zqsat5 * zcor5 * z2s5);
    z2s = z2s - zcond1 * (zq55a - zqsat5) * zqsat5 * zcor5 / SQ(1.0 + zqsat5 * zcor5 * z2s5);
    zcond1 = 0.0;
    ztarg5 = ztarg55a;
    ztarg = ztarg - 2.0 * z2s * z5alcp / CUBE(ztarg5 - z4es);
    z2s = 0.0;
    zqsat5 = zqsat55c;
    zcor = zcor + zqsat * zqsat5;
    zqsat = zqsat * zcor5;
    zqsat = zqsat + zcor * P->retv / SQ(1.0 - P->retv * zqsat5);
    zcor = 0.0;
    if (lltest1) zqsat = 0.0;
    zfoeew = zfoeew + zqsat * zqp5;
    zfoeew5 = zfoeew55a;
    zqp = zqp + zqsat * zfoeew5;
    zqsat = 0.0;
    ztarg = ztarg + zfoeew * P->r2es * z3es * (P->rtt - z4es) *
                        exp(z3es * (ztarg5 - P->rtt) / (ztarg5 - z4es)) / SQ(ztarg5 - z4es);
    zfoeew = 0.0;
    pt[IX(jl, kk)] = pt[IX(jl, kk)] + ztarg;
    ztarg = 0.0;

    /* :597-641 adjoint of the first iteration */
    zcond1 = zcond1 - pq[IX(jl, kk)];
    zcond1 = zcond1 + zaldcp * pt[IX(jl, kk)];
    zqsat5 = zqsat55b; zcor5 = zcor55b; z2s5 = z2s55b;
    pq[IX(jl, kk)] = pq[IX(jl, kk)] + zcond1 / (1.0 + zqsat5 * zcor5 * z2s5);
    zqsat = zqsat - zcond1 / (1.0 + zqsat5 * zcor5 * z2s5);
    zqsat = zqsat - zcond1 * (zq55b - zqsat5) * zcor5 * z2s5 / SQ(1.0 + zqsat5 * zcor5 * z2s5);
    zcor = zcor - zcond1 * (zq55b - zqsat5) * zqsat5 * z2s5 / SQ(1.0 + zqsat5 * zcor5 * z2s5);
    z2s = z2s - zcond1 * (zq55b - zqsat5) * zqsat5 * zcor5 / SQ(1.0 + zqsat5 * zcor5 * z2s5);
    zcond1 = 0.0;
    ztarg5 = ztarg55b;
    ztarg = ztarg - 2.0 * z2s * z5alcp / CUBE(ztarg5 - z4es);
    z2s = 0.0;
    zqsat5 = zqsat55d;
    zcor = zcor + zqsat * zqsat5;
    zqsat = zqsat * zcor5;
    zqsat = zqsat + zcor * P->retv / SQ(1.0 - P->retv * zqsat5);
    zcor = 0.0;
    if (lltest2) zqsat = 0.0;
    zfoeew = zfoeew + zqsat * zqp5;
    zfoeew5 = zfoeew55b;
    zqp = zqp + zqsat * zfoeew5;
    zqsat = 0.0;
    ztarg = ztarg + zfoeew * P->r2es * z3es * (P->rtt - z4es) *
                        exp(z3es * (ztarg5 - P->rtt) / (ztarg5 - z4es)) / SQ(ztarg5 - z4es);
    zfoeew = 0.0;
    pt[IX(jl, kk)] = pt[IX(jl, kk)] + ztarg;
    ztarg = 0.0;
    psp[jl] = psp[jl] - zqp / SQ(psp5[jl]);
    zqp = 0.0;
    (void)zqp; (void)ztarg; (void)zfoeew;
  }
}

int orc_cloudsc2ad(const cloudsc2_params *P, const double *ceta, int kidia, int kfdia, int klon,
                   int klev, double ptsphy, const orc_in16 *in5, const orc_out10 *out5,
                   const orc_in16 *in, const orc_out10 *out) {
  if (P->levapls2 || P->ldrain1d || !P->lphylin) return -1;
  const int lregcl = P->lregcl;
  const double zscal = 0.9;
  const size_t n2 = (size_t)klon * (size_t)klev;

  const double *paphp15 = in5->paphp1, *papp15 = in5->papp1, *pqm15 = in5->pqm1, *pqs5 = in5->pqs,
               *ptm15 = in5->ptm1, *pl5 = in5->pl, *pi5 = in5->pi, *plude5 = in5->plude,
               *plu5 = in5->plu, *pmfu5 = in5->pmfu, *pmfd5 = in5->pmfd, *pgtent5 = in5->pgtent,
               *pgtenq5 = in5->pgtenq, *pgtenl5 = in5->pgtenl, *pgteni5 = in5->pgteni,
               *psupsat5 = in5->psupsat;
  double *ptent5 = out5->ptent, *ptenq5 = out5->ptenq, *ptenl5 = out5->ptenl,
         *pteni5 = out5->pteni, *pclc5 = out5->pclc, *pfplsl5 = out5->pfplsl,
         *pfplsn5 = out5->pfplsn, *pfhpsl5 = out5->pfhpsl, *pfhpsn5 = out5->pfhpsn,
         *pcovptot5 = out5->pcovptot;
  double *paphp1 = in->paphp1, *papp1 = in->papp1, *pqm1 = in->pqm1, *pqs = in->pqs,
         *ptm1 = in->ptm1, *pl = in->pl, *pi = in->pi, *plude = in->plude, *plu = in->plu,
         *pmfu = in->pmfu, *pmfd = in->pmfd, *pgtent = in->pgtent, *pgtenq = in->pgtenq,
         *pgtenl = in->pgtenl, *pgteni = in->pgteni, *psupsat = in->psupsat;
  double *ptent = out->ptent, *ptenq = out->ptenq, *ptenl = out->ptenl, *pteni = out->pteni,
         *pclc = out->pclc, *pfplsl = out->pfplsl, *pfplsn = out->pfplsn, *pfhpsl = out->pfhpsl,
         *pfhpsn = out->pfhpsn, *pcovptot = out->pcovptot;

  /* (KLON,KLEV) storage: trajectory (:228-292) + adjoint accumulators (:253-266) */
  enum { NW2 = 96 };
  double *w = (double *)calloc((size_t)NW2 * n2 + 3 * (size_t)klon * (klev + 1) + (size_t)klev +
                                   24 * (size_t)klon, sizeof(double));
  char *llo3 = (char *)calloc(n2, 1);
  if (!w || !llo3) { free(w); free(llo3); return -2; }
  double *q = w;
#define W2(name) double *name = q; q += n2
  /* trajectory */
  W2(ztp15); W2(zqp15); W2(zl5); W2(zi5); W2(zlude5); W2(zqlwc5); W2(zqiwc5); W2(zdp5);
  W2(zlsdcp5); W2(zlfdcp5); W2(zlvdcp5); W2(zfwatr15); W2(zfwatr25); W2(zrfreeze15);
  W2(zrfreeze35); W2(zqt5); W2(zqc15); W2(zqc25); W2(zqc35); W2(zrho5); W2(zrodqsdp5);
  W2(zldcp5); W2(dtdzmo5); W2(zdqsdz5); W2(zcondl15); W2(zcondl25); W2(zcondi15); W2(zcondi25);
  W2(zqsat5); W2(zsupsat5); W2(zcldl5); W2(zcldi5); W2(zprr5); W2(zprs5); W2(zcons5);
  W2(zsnmlt5); W2(zz2s5); W2(zgdp5); W2(zqcrit5); W2(zdr15); W2(zdr25); W2(ztp25); W2(ztp35);
  W2(zqp25); W2(ztpb5); W2(zqpb5); W2(zqlwc15); W2(zqiwc15); W2(zclc5); W2(zexp15); W2(zexp25);
  W2(zexp35); W2(zexpdl5); W2(zexpdi5); W2(zfwat5); W2(zfoeew5); W2(zesdp5); W2(zesdp15);
  W2(zfac5); W2(zfacw5); W2(zfaci5); W2(zcor5); W2(zdqsdtemp5); W2(zqold5); W2(zdq5);
  W2(zqpd5); W2(zqcd5); W2(zsqrt5); W2(zdqc5); W2(zfac1); W2(zfac2); W2(zfac3); W2(zfac4);
  W2(zcrh2); W2(zcovptot15); W2(zcovpclr15);
  /* adjoint */
  W2(ztp1); W2(zqp1); W2(zl); W2(zi); W2(zlude); W2(zqlwc); W2(zqiwc); W2(zdp); W2(zlsdcp);
  W2(zlfdcp); W2(zlvdcp); W2(zrfreeze); W2(zcondl); W2(zcondi);
#undef W2
  double *zrfl5 = q; q += (size_t)klon * (klev + 1);        /* ZRFL5(KLON,KLEV+1)   */
  double *zsfl5 = q; q += (size_t)klon * (klev + 1);        /* ZSFL5(KLON,KLEV+1)   */
  double *zcovptot5 = q; q += (size_t)klon * (klev + 1);    /* ZCOVPTOT5(KLON,0:KLEV) -> index jk+1 */
  double *zscalm = q; q += klev;
#define W1(name) double *name = q; q += klon
  W1(zrfln5); W1(zsfln5); W1(zpp5); W1(ztrpaus);
  W1(zrfl); W1(zsfl); W1(zrfln); W1(zsfln); W1(zgdp); W1(zqcrit); W1(zdqsdtemp); W1(zqold);
  W1(zpp); W1(zdq); W1(zqc); W1(zqsat); W1(zfwat); W1(zfoeew);
#undef W1

  /* :344-356 */
  const double zckcodtl = 2.0 * P->rkconv * ptsphy;
  const double zckcodti = 5.0 * P->rkconv * ptsphy;
  const double zckcodtla = zckcodtl / 100.0;
  const double zckcodtia = zckcodti / 100.0;
  const double zcons2 = 1.0 / (ptsphy * P->rg);
  const double zcons3 = P->rlvtt / P->rcpd;
  const double zmeltp2 = P->rtt + 2.0;
  const double zqtmst = 1.0 / ptsphy;
  const double zqmax = 0.5, zeps1 = 1.e-12, zeps2 = 1.e-10;

  /* ===================== forward (trajectory) sweep, :364-866 ===================== */
  for (int jk = 0; jk < klev; ++jk)   /* :365-377 */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      A(ztp15) = A(ptm15) + ptsphy * A(pgtent5);
      A(zqp15) = A(pqm15) + ptsphy * A(pgtenq5) + A(psupsat5);
      A(zl5) = A(pl5) + ptsphy * A(pgtenl5);
      A(zi5) = A(pi5) + ptsphy * A(pgteni5);
      A(ztp25) = A(ztp15);
      A(ztp35) = A(ztp15);
      A(zqp25) = A(zqp15);
    }
  for (int jk = 0; jk < klev; ++jk) {   /* :379-394 */
    zscalm[jk] = zscal * pow(fmax(ceta[jk] - 0.2, zeps1), 0.2);
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      A(zdp5) = paphp15[IX(jl, jk + 1)] - paphp15[IX(jl, jk)];
      double zzz5 = 1.0 / (P->rcpd + P->rcpd * P->rvtmp2 * A(zqp15));
      A(zlfdcp5) = P->rlmlt * zzz5;
      A(zlsdcp5) = P->rlstt * zzz5;
      A(zlvdcp5) = P->rlvtt * zzz5;
    }
  }
  for (int jk = 0; jk < klev; ++jk)     /* :403-412 */
    for (int jl = kidia - 1; jl < kfdia; ++jl) { A(pclc5) = 0.0; A(pcovptot5) = 0.0; }
  for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :416-423 */
    zrfl5[IX(jl, 0)] = 0.0; zsfl5[IX(jl, 0)] = 0.0;
    pfplsl5[IX(jl, 0)] = 0.0; pfplsn5[IX(jl, 0)] = 0.0;
    zcovptot5[IX(jl, 0)] = 0.0;
  }
  for (int jl = kidia - 1; jl < kfdia; ++jl) ztrpaus[jl] = 0.1;   /* :426-437 */
  for (int jk = 0; jk < klev - 1; ++jk)
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      int llo1 = ceta[jk] > 0.1 && ceta[jk] < 0.4 && ztp15[IX(jl, jk)] > ztp15[IX(jl, jk + 1)];
      if (llo1) ztrpaus[jl] = ceta[jk];
    }

  for (int jk = 0; jk < klev; ++jk) {   /* :449 */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :453-527 */
      double z3es, z4es;
      A(zdr15) = 0.0; A(zdr25) = 0.0;
      double zoealfaw5 = 0.545 * (tanh(0.17 * (A(ztp25) - P->rlptrc)) + 1.0);
      if (A(ztp25) < P->rtt) { A(zfwat5) = zoealfaw5; z3es = P->r3ies; z4es = P->r4ies; }
      else { A(zfwat5) = 1.0; z3es = P->r3les; z4es = P->r4les; }
      A(zfoeew5) = P->r2es * exp(z3es * (A(ztp25) - P->rtt) / (A(ztp25) - z4es));
      A(zesdp15) = A(zfoeew5) / A(papp15);
      A(zesdp5) = A(zesdp15);
      if (A(zesdp15) > zqmax) A(zesdp5) = zqmax;
      A(zfacw5) = P->r5les / SQ(A(ztp25) - P->r4les);
      A(zfaci5) = P->r5ies / SQ(A(ztp25) - P->r4ies);
      A(zfac5) = A(zfwat5) * A(zfacw5) + (1.0 - A(zfwat5)) * A(zfaci5);
      A(zcor5) = 1.0 / (1.0 - P->retv * A(zesdp5));
      A(zdqsdtemp5) = A(zfac5) * A(zcor5) * A(pqs5);
      /* ZCORQS5, ZQLIM5 (:481-489) feed only the dead LLO2 block */
      double zeta3 = ztrpaus[jl];
      double zrh1 = 1.0;
      double zrh2 = 0.35 + 0.14 * SQ((zeta3 - 0.25) / 0.15) + 0.04 * dmin(zeta3 - 0.25, 0.0) / 0.15;
      double zrh3 = 1.0;
      double zdeta2 = 0.3;
      double zdeta1 = 0.09 + 0.16 * (0.4 - zeta3) / 0.3;
      if (ceta[jk] < zeta3) A(zcrh2) = zrh3;
      else if (ceta[jk] >= zeta3 && ceta[jk] < (zeta3 + zdeta2))
        A(zcrh2) = zrh3 + (zrh2 - zrh3) * ((ceta[jk] - zeta3) / zdeta2);
      else if (ceta[jk] >= (zeta3 + zdeta2) && ceta[jk] < (1.0 - zdeta1)) A(zcrh2) = zrh2;
      else if (ceta[jk] >= (1.0 - zdeta1))
        A(zcrh2) = zrh1 + (zrh2 - zrh1) * sqrt((1.0 - ceta[jk]) / zdeta1);
      if (A(ztp25) < P->rtice) A(zsupsat5) = 1.8 - 3.e-03 * A(ztp25);
      else A(zsupsat5) = 1.0;
      A(zqsat5) = A(pqs5) * A(zsupsat5);
      A(zqcrit5) = A(zcrh2) * A(zqsat5);
    }
    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :531-550 */
      A(zqt5) = A(zqp15) + A(zl5) + A(zi5);
      if (A(zqt5) <= A(zqcrit5)) { A(zclc5) = 0.0; A(zqc15) = 0.0; }
      else if (A(zqt5) >= A(zqsat5)) {
        A(zclc5) = 1.0;
        A(zqc15) = (1.0 - zscalm[jk]) * (A(zqsat5) - A(zqcrit5));
      } else {
        A(zqpd5) = A(zqsat5) - A(zqt5);
        A(zqcd5) = A(zqsat5) - A(zqcrit5);
        A(zsqrt5) = sqrt(A(zqpd5) / (A(zqcd5) - zscalm[jk] * (A(zqt5) - A(zqcrit5))));
        A(zclc5) = 1.0 - A(zsqrt5);
        A(zqc15) = (zscalm[jk] * A(zqpd5) + (1.0 - zscalm[jk]) * A(zqcd5)) * SQ(A(zclc5));
      }
    }
    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :554-574 */
      A(zgdp5) = P->rg / (paphp15[IX(jl, jk + 1)] - paphp15[IX(jl, jk)]);
      A(zlude5) = A(plude5) * ptsphy * A(zgdp5);
      int llo1;
      if (jk < klev - 1) llo1 = A(zlude5) >= P->rlmin && plu5[IX(jl, jk + 1)] >= zeps2;
      else llo1 = 0;
      if (llo1) {
        A(pclc5) = A(zclc5) + (1.0 - A(zclc5)) * (1.0 - exp(-A(zlude5) / plu5[IX(jl, jk + 1)]));
        A(zqc25) = A(zqc15) + A(zlude5);
      } else { A(pclc5) = A(zclc5); A(zqc25) = A(zqc15); }
    }
    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :578-600 */
      A(zfac1) = 1.0 / (P->rd * A(ztp25));
      A(zrho5) = A(papp15) * A(zfac1);
      A(zfac2) = 1.0 / (A(papp15) - P->retv * A(zfoeew5));
      A(zrodqsdp5) = -A(zrho5) * A(pqs5) * A(zfac2);
      A(zldcp5) = A(zfwat5) * A(zlvdcp5) + (1.0 - A(zfwat5)) * A(zlsdcp5);
      A(zfac3) = 1.0 / (1.0 + A(zldcp5) * A(zdqsdtemp5));
      A(dtdzmo5) = P->rg * (1.0 / P->rcpd - A(zldcp5) * A(zrodqsdp5)) * A(zfac3);
      A(zdqsdz5) = A(zdqsdtemp5) * A(dtdzmo5) - P->rg * A(zrodqsdp5);
      A(zfac4) = 1.0 / A(zrho5);
      A(llo3) = (A(zdqsdz5) * (A(pmfu5) + A(pmfd5)) * ptsphy * A(zfac4) < A(zqc25));
      if (A(llo3)) A(zdqc5) = A(zdqsdz5) * (A(pmfu5) + A(pmfd5)) * ptsphy * A(zfac4);
      else A(zdqc5) = A(zqc25);
      A(zqc35) = A(zqc25) - A(zdqc5);
    }
    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :604-609 */
      A(zqlwc15) = A(zqc35) * A(zfwat5);
      A(zqiwc15) = A(zqc35) * (1.0 - A(zfwat5));
      A(zcondl15) = (A(zqlwc15) - A(zl5)) * zqtmst;
      A(zcondi15) = (A(zqiwc15) - A(zi5)) * zqtmst;
    }
    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :614-627 (ZCOVPTOT5 index shifted by one) */
      if (A(pclc5) > zcovptot5[IX(jl, jk)]) A(zcovptot15) = A(pclc5);
      else A(zcovptot15) = zcovptot5[IX(jl, jk)];
      zcovptot5[IX(jl, jk + 1)] = A(zcovptot15);
      A(zcovpclr15) = zcovptot5[IX(jl, jk + 1)] - A(pclc5);
    }
    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :633-651 melting */
      if (zsfl5[IX(jl, jk)] != 0.0) {
        A(zcons5) = zcons2 * A(zdp5) / A(zlfdcp5);
        if ((A(ztp25) - zmeltp2) > 0.0) A(zz2s5) = A(zcons5) * (A(ztp25) - zmeltp2);
        else A(zz2s5) = 0.0;
        if (zsfl5[IX(jl, jk)] <= A(zz2s5)) A(zsnmlt5) = zsfl5[IX(jl, jk)];
        else A(zsnmlt5) = A(zz2s5);
        zrfln5[jl] = zrfl5[IX(jl, jk)] + A(zsnmlt5);
        zsfln5[jl] = zsfl5[IX(jl, jk)] - A(zsnmlt5);
        A(ztp15) = A(ztp25) - A(zsnmlt5) / A(zcons5);
      } else {
        zrfln5[jl] = zrfl5[IX(jl, jk)];
        zsfln5[jl] = zsfl5[IX(jl, jk)];
      }
    }
    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :655-722 */
      if (A(pclc5) > zeps2) {
        double zlcrit = P->rclcrit * 2.0;
        A(zcldl5) = A(zqlwc15) / A(pclc5);
        A(zexp35) = exp(-SQ(A(zcldl5) / zlcrit));
        double zdl5 = zckcodtl * (1.0 - A(zexp35));
        A(zexpdl5) = exp(-zdl5);
        double zlnew5 = A(pclc5) * A(zcldl5) * A(zexpdl5);
        A(zprr5) = A(zqlwc15) - zlnew5;
        A(zqlwc5) = A(zqlwc15) - A(zprr5);
      } else { A(zprr5) = 0.0; A(zqlwc5) = A(zqlwc15); }
      if (A(pclc5) > zeps2) {
        double zlcrit = P->rclcrit * 2.0;
        A(zcldi5) = A(zqiwc15) / A(pclc5);
        A(zexp15) = exp(0.025 * (A(ztp15) - P->rtt));
        A(zexp25) = exp(-SQ(A(zcldi5) / zlcrit));
        double zdi5 = zckcodti * A(zexp15) * (1.0 - A(zexp25));
        A(zexpdi5) = exp(-zdi5);
        double zinew5 = A(pclc5) * A(zcldi5) * A(zexpdi5);
        A(zprs5) = A(zqiwc15) - zinew5;
        A(zqiwc5) = A(zqiwc15) - A(zprs5);
      } else { A(zprs5) = 0.0; A(zqiwc5) = A(zqiwc15); }
      A(zdr15) = zcons2 * A(zdp5) * (A(zprr5) + A(zprs5));
      if (A(ztp15) < P->rtt) { A(zrfreeze15) = zcons2 * A(zdp5) * A(zprr5); A(zfwatr15) = 0.0; }
      else A(zfwatr15) = 1.0;
      double zrn5 = A(zfwatr15) * A(zdr15);
      double zsn5 = (1.0 - A(zfwatr15)) * A(zdr15);
      zrfln5[jl] = zrfln5[jl] + zrn5;
      zsfln5[jl] = zsfln5[jl] + zsn5;
      /* :724-773 LLO2 evaporation: statically dead */
    }
    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :777-797 */
      double zdqdt5 = -(A(zcondl15) + A(zcondi15)) + (A(plude5) + 0.0 + 0.0) * A(zgdp5);
      double zdtdt5 = A(zlvdcp5) * A(zcondl15) + A(zlsdcp5) * A(zcondi15) -
                      (A(zlvdcp5) * 0.0 + A(zlsdcp5) * 0.0 +
                       A(plude5) * (A(zfwat5) * A(zlvdcp5) + (1.0 - A(zfwat5)) * A(zlsdcp5)) -
                       (A(zlsdcp5) - A(zlvdcp5)) * A(zrfreeze15)) * A(zgdp5);
      A(ztp35) = A(ztp15) + ptsphy * zdtdt5;
      A(zqp15) = A(zqp25) + ptsphy * zdqdt5;
      A(ztpb5) = A(ztp35);
      A(zqpb5) = A(zqp15);
      zpp5[jl] = A(papp15);
      A(zqold5) = A(zqp15);
    }
    orc_cuadjtqs(P, kidia, kfdia, klon, jk, zpp5, ztp35, zqp15);   /* :803-804 */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :806-832 */
      if ((A(zqold5) - A(zqp15)) >= 0.0) A(zdq5) = A(zqold5) - A(zqp15);
      else A(zdq5) = 0.0;
      A(zdr25) = zcons2 * A(zdp5) * A(zdq5);
      double zrfreeze25;
      if (A(ztp35) < P->rtt) { zrfreeze25 = A(zfwat5) * A(zdr25); A(zfwatr25) = 0.0; }
      else { zrfreeze25 = 0.0; A(zfwatr25) = 1.0; }
      double zrn5 = A(zfwatr25) * A(zdr25);
      double zsn5 = (1.0 - A(zfwatr25)) * A(zdr25);
      A(zcondl25) = A(zcondl15) + A(zfwatr25) * A(zdq5) * zqtmst;
      A(zcondi25) = A(zcondi15) + (1.0 - A(zfwatr25)) * A(zdq5) * zqtmst;
      zrfln5[jl] = zrfln5[jl] + zrn5;
      zsfln5[jl] = zsfln5[jl] + zsn5;
      A(zrfreeze35) = A(zrfreeze15) + zrfreeze25;
    }
    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :834-856 */
      double zdqdt5 = -(A(zcondl25) + A(zcondi25)) + (A(plude5) + 0.0 + 0.0) * A(zgdp5);
      double zdtdt5 = A(zlvdcp5) * A(zcondl25) + A(zlsdcp5) * A(zcondi25) -
                      (A(zlvdcp5) * 0.0 + A(zlsdcp5) * 0.0 +
                       A(plude5) * (A(zfwat5) * A(zlvdcp5) + (1.0 - A(zfwat5)) * A(zlsdcp5)) -
                       (A(zlsdcp5) - A(zlvdcp5)) * A(zrfreeze35)) * A(zgdp5);
      double zdldt5 = (A(zqlwc5) - A(zl5)) * zqtmst;
      double zdidt5 = (A(zqiwc5) - A(zi5)) * zqtmst;
      A(ptenq5) = zdqdt5; A(ptent5) = zdtdt5; A(ptenl5) = zdldt5; A(pteni5) = zdidt5;
      pfplsl5[IX(jl, jk + 1)] = zrfln5[jl];
      pfplsn5[IX(jl, jk + 1)] = zsfln5[jl];
    }
    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :851-854 */
      zrfl5[IX(jl, jk + 1)] = zrfln5[jl];
      zsfl5[IX(jl, jk + 1)] = zsfln5[jl];
    }
  }
  for (int jk = 0; jk < klev + 1; ++jk)   /* :861-866 */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      A(pfhpsl5) = -A(pfplsl5) * P->rlvtt;
      A(pfhpsn5) = -A(pfplsn5) * P->rlstt;
    }

  /* ============================ adjoint computation ============================ */
  /* :877-909 adjoint accumulators are zero (calloc) */

  for (int jk = klev; jk >= 0; --jk)   /* :914-921 enthalpy fluxes */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      A(pfplsn) = A(pfplsn) - A(pfhpsn) * P->rlstt;
      A(pfhpsn) = 0.0;
      A(pfplsl) = A(pfplsl) - A(pfhpsl) * P->rlvtt;
      A(pfhpsl) = 0.0;
    }

  for (int jk = klev - 1; jk >= 0; --jk) {   /* :934 */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :939-944 */
      zsfln[jl] = zsfln[jl] + zsfl[jl]; zsfl[jl] = 0.0;
      zrfln[jl] = zrfln[jl] + zrfl[jl]; zrfl[jl] = 0.0;
    }
    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :946-1013 */
      double zdtdt = 0.0, zdqdt = 0.0, zdldt = 0.0, zdidt = 0.0;
      zsfln[jl] = zsfln[jl] + pfplsn[IX(jl, jk + 1)];
      pfplsn[IX(jl, jk + 1)] = 0.0;
      zrfln[jl] = zrfln[jl] + pfplsl[IX(jl, jk + 1)];
      pfplsl[IX(jl, jk + 1)] = 0.0;
      zdtdt = zdtdt + A(ptent); zdqdt = zdqdt + A(ptenq);
      zdldt = zdldt + A(ptenl); zdidt = zdidt + A(pteni);
      A(ptent) = 0.0; A(ptenq) = 0.0; A(ptenl) = 0.0; A(pteni) = 0.0;
      /* qice / qliq tendencies */
      A(zi) = A(zi) - zqtmst * zdidt;
      A(zqiwc) = A(zqiwc) + zqtmst * zdidt;
      zdidt = 0.0;
      A(zl) = A(zl) - zqtmst * zdldt;
      A(zqlwc) = A(zqlwc) + zqtmst * zdldt;
      zdldt = 0.0;
      /* T tendency (:977-1002), with ZEVAPR5 = ZEVAPS5 = 0 */
      zgdp[jl] = zgdp[jl] - zdtdt * (A(zlvdcp5) * 0.0 + A(zlsdcp5) * 0.0 +
                                     A(plude5) * (A(zfwat5) * A(zlvdcp5) +
                                                  (1.0 - A(zfwat5)) * A(zlsdcp5)) -
                                     (A(zlsdcp5) - A(zlvdcp5)) * A(zrfreeze35));
      A(zcondl) = A(zcondl) + zdtdt * A(zlvdcp5);
      A(zcondi) = A(zcondi) + zdtdt * A(zlsdcp5);
      A(zlvdcp) = A(zlvdcp) + zdtdt * (A(zcondl25) - 0.0 * A(zgdp5));
      A(zlsdcp) = A(zlsdcp) + zdtdt * (A(zcondi25) - 0.0 * A(zgdp5));
      A(plude) = A(plude) - zdtdt * A(zgdp5) * (A(zfwat5) * A(zlvdcp5) + (1.0 - A(zfwat5)) * A(zlsdcp5));
      A(zlvdcp) = A(zlvdcp) - zdtdt * A(plude5) * A(zgdp5) * A(zfwat5);
      A(zlsdcp) = A(zlsdcp) - zdtdt * A(plude5) * A(zgdp5) * (1.0 - A(zfwat5));
      zfwat[jl] = zfwat[jl] - zdtdt * A(plude5) * A(zgdp5) * (A(zlvdcp5) - A(zlsdcp5));
      A(zlsdcp) = A(zlsdcp) + zdtdt * A(zrfreeze35) * A(zgdp5);
      A(zlvdcp) = A(zlvdcp) - zdtdt * A(zrfreeze35) * A(zgdp5);
      A(zrfreeze) = A(zrfreeze) + zdtdt * (A(zlsdcp5) - A(zlvdcp5)) * A(zgdp5);
      zdtdt = 0.0;
      /* q tendency (:1005-1012) */
      zgdp[jl] = zgdp[jl] + zdqdt * (A(plude5) + 0.0 + 0.0);
      A(plude) = A(plude) + zdqdt * A(zgdp5);
      A(zcondl) = A(zcondl) - zdqdt;
      A(zcondi) = A(zcondi) - zdqdt;
      zdqdt = 0.0;
      (void)zdidt; (void)zdldt; (void)zdtdt; (void)zdqdt;
    }

    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :1017-1067 */
      double zfwatr = 0.0, zsn = 0.0, zrn = 0.0, zdr2 = 0.0;
      double zrfreeze2 = A(zrfreeze);
      zsn = zsn + zsfln[jl];
      zrn = zrn + zrfln[jl];
      zdq[jl] = zdq[jl] + A(zcondi) * (1.0 - A(zfwatr25)) * zqtmst;
      zfwatr = zfwatr - A(zcondi) * A(zdq5) * zqtmst;
      zdq[jl] = zdq[jl] + A(zcondl) * A(zfwatr25) * zqtmst;
      zfwatr = zfwatr + A(zcondl) * A(zdq5) * zqtmst;
      zdr2 = zdr2 + (1.0 - A(zfwatr25)) * zsn;
      zfwatr = zfwatr - A(zdr25) * zsn;
      zdr2 = zdr2 + A(zfwatr25) * zrn;
      zfwatr = zfwatr + A(zdr25) * zrn;
      if (A(ztp35) < P->rtt) {
        zfwatr = 0.0;
        zfwat[jl] = zfwat[jl] + A(zdr25) * zrfreeze2;
        zdr2 = zdr2 + A(zfwat5) * zrfreeze2;
      } else {
        zfwatr = 0.0;
      }
      zdq[jl] = zdq[jl] + zcons2 * A(zdp5) * zdr2;
      A(zdp) = A(zdp) + zcons2 * A(zdq5) * zdr2;
      zdr2 = 0.0;
      if ((A(zqold5) - A(zqp15)) >= 0.0) {
        if (lregcl) zdq[jl] = zdq[jl] * 0.7;
        zqold[jl] = zqold[jl] + zdq[jl];
        A(zqp1) = A(zqp1) - zdq[jl];
      }
      zdq[jl] = 0.0;
      zpp5[jl] = A(papp15);
      (void)zfwatr; (void)zdr2;
    }

    /* :1069-1072 ; trajectory copies ZTPB5/ZQPB5 are re-adjusted in place, as in the reference */
    orc_cuadjtqsad(P, kidia, kfdia, klon, jk, zpp5, ztpb5, zqpb5, zpp, ztp1, zqp1);

    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :1074-1126 */
      double zdtdt = 0.0, zdqdt = 0.0;
      A(zqp1) = A(zqp1) + zqold[jl];
      zqold[jl] = 0.0;
      A(papp1) = A(papp1) + zpp[jl];
      zpp[jl] = 0.0;
      zdqdt = zdqdt + ptsphy * A(zqp1);
      zdtdt = zdtdt + ptsphy * A(ztp1);
      zgdp[jl] = zgdp[jl] - zdtdt * (A(zlvdcp5) * 0.0 + A(zlsdcp5) * 0.0 +
                                     A(plude5) * (A(zfwat5) * A(zlvdcp5) +
                                                  (1.0 - A(zfwat5)) * A(zlsdcp5)) -
                                     (A(zlsdcp5) - A(zlvdcp5)) * A(zrfreeze15));
      A(zcondl) = A(zcondl) + zdtdt * A(zlvdcp5);
      A(zcondi) = A(zcondi) + zdtdt * A(zlsdcp5);
      A(zlvdcp) = A(zlvdcp) + zdtdt * (A(zcondl15) - 0.0 * A(zgdp5));
      A(zlsdcp) = A(zlsdcp) + zdtdt * (A(zcondi15) - 0.0 * A(zgdp5));
      A(plude) = A(plude) - zdtdt * A(zgdp5) * (A(zfwat5) * A(zlvdcp5) + (1.0 - A(zfwat5)) * A(zlsdcp5));
      A(zlvdcp) = A(zlvdcp) - zdtdt * A(plude5) * A(zgdp5) * A(zfwat5);
      A(zlsdcp) = A(zlsdcp) - zdtdt * A(plude5) * A(zgdp5) * (1.0 - A(zfwat5));
      zfwat[jl] = zfwat[jl] - zdtdt * A(plude5) * A(zgdp5) * (A(zlvdcp5) - A(zlsdcp5));
      A(zlsdcp) = A(zlsdcp) + zdtdt * A(zrfreeze15) * A(zgdp5);
      A(zlvdcp) = A(zlvdcp) - zdtdt * A(zrfreeze15) * A(zgdp5);
      A(zrfreeze) = A(zrfreeze) + zdtdt * (A(zlsdcp5) - A(zlvdcp5)) * A(zgdp5);
      zdtdt = 0.0;
      zgdp[jl] = zgdp[jl] + zdqdt * (A(plude5) + 0.0 + 0.0);
      A(plude) = A(plude) + zdqdt * A(zgdp5);
      A(zcondl) = A(zcondl) - zdqdt;
      A(zcondi) = A(zcondi) - zdqdt;
      zdqdt = 0.0;
      (void)zdtdt; (void)zdqdt;
    }

    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :1128-1358 */
      double zsn = 0.0, zrn = 0.0, zprr = 0.0, zprs = 0.0, zlnew = 0.0, zinew = 0.0;
      double zcldl = 0.0, zcldi = 0.0, zdl = 0.0, zdi = 0.0, zprtot = 0.0, zfwatr = 0.0, zdr = 0.0;
      /* :1152-1267 LLO2 evaporation adjoint: statically dead */
      zrfln[jl] = zrfln[jl] + zprtot;   /* :1271-1272 */
      zsfln[jl] = zsfln[jl] + zprtot;
      zsn = zsn + zsfln[jl];
      zrn = zrn + zrfln[jl];
      zfwatr = zfwatr - A(zdr15) * zsn;
      zdr = zdr + (1.0 - A(zfwatr15)) * zsn;
      zfwatr = zfwatr + A(zdr15) * zrn;
      zdr = zdr + A(zfwatr15) * zrn;
      if (A(ztp15) < P->rtt) {   /* :1284-1288 */
        A(zdp) = A(zdp) + A(zrfreeze) * zcons2 * A(zprr5);
        zprr = zprr + A(zrfreeze) * zcons2 * A(zdp5);
        A(zrfreeze) = 0.0;
      }
      zfwatr = 0.0;
      zprr = zprr + zcons2 * A(zdp5) * zdr;
      zprs = zprs + zcons2 * A(zdp5) * zdr;
      A(zdp) = A(zdp) + zcons2 * (A(zprr5) + A(zprs5)) * zdr;
      zdr = 0.0;
      if (A(pclc5) > zeps2) {   /* :1298-1327 ice */
        double zlcrit = P->rclcrit * 2.0;
        zprs = zprs - A(zqiwc);
        A(zqiwc) = A(zqiwc) + zprs;
        zinew = zinew - zprs;
        A(pclc) = A(pclc) + zinew * A(zcldi5) * A(zexpdi5);
        zcldi = zcldi + zinew * A(pclc5) * A(zexpdi5);
        zdi = zdi - zinew * A(pclc5) * A(zcldi5) * A(zexpdi5);
        if (lregcl) {
          A(ztp1) = A(ztp1) + zckcodtia * A(zexp15) * (1.0 - A(zexp25)) * 0.025 * zdi;
          zcldi = zcldi + (zckcodtia * A(zexp15) * A(zexp25) * 2.0 * A(zcldi5) / SQ(zlcrit)) * zdi;
        } else {
          A(ztp1) = A(ztp1) + zckcodti * A(zexp15) * (1.0 - A(zexp25)) * 0.025 * zdi;
          zcldi = zcldi + (zckcodti * A(zexp15) * A(zexp25) * 2.0 * A(zcldi5) / SQ(zlcrit)) * zdi;
        }
        A(zqiwc) = A(zqiwc) + zcldi / A(pclc5);
        A(pclc) = A(pclc) - A(zqiwc15) * zcldi / SQ(A(pclc5));
      }
      zprs = 0.0;
      if (A(pclc5) > zeps2) {   /* :1332-1356 liquid */
        double zlcrit = P->rclcrit * 2.0;
        zprr = zprr - A(zqlwc);
        A(zqlwc) = A(zqlwc) + zprr;
        zlnew = zlnew - zprr;
        A(pclc) = A(pclc) + zlnew * A(zcldl5) * A(zexpdl5);
        zcldl = zcldl + zlnew * A(pclc5) * A(zexpdl5);
        zdl = zdl - zlnew * A(pclc5) * A(zcldl5) * A(zexpdl5);
        if (lregcl) zcldl = zcldl + (2.0 * zckcodtla / SQ(zlcrit)) * A(zexp35) * A(zcldl5) * zdl;
        else zcldl = zcldl + (2.0 * zckcodtl / SQ(zlcrit)) * A(zexp35) * A(zcldl5) * zdl;
        A(zqlwc) = A(zqlwc) + zcldl / A(pclc5);
        A(pclc) = A(pclc) - A(zqlwc15) * zcldl / SQ(A(pclc5));
      }
      zprr = 0.0;
      (void)zprr; (void)zprs; (void)zfwatr; (void)zdr;
    }

    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :1362-1400 melting */
      double zsnmlt = 0.0, zcons = 0.0, zz2s = 0.0;
      if (zsfl5[IX(jl, jk)] != 0.0) {
        zsnmlt = zsnmlt - A(ztp1) / A(zcons5);
        zcons = zcons + (A(ztp1) * A(zsnmlt5)) / SQ(A(zcons5));
        zsfl[jl] = zsfl[jl] + zsfln[jl];
        zsnmlt = zsnmlt - zsfln[jl];
        zsfln[jl] = 0.0;
        zrfl[jl] = zrfl[jl] + zrfln[jl];
        zsnmlt = zsnmlt + zrfln[jl];
        zrfln[jl] = 0.0;
        if (zsfl5[IX(jl, jk)] <= A(zz2s5)) zsfl[jl] = zsfl[jl] + zsnmlt;
        else zz2s = zz2s + zsnmlt;
        if ((A(ztp25) - zmeltp2) > 0.0) {
          A(ztp1) = A(ztp1) + A(zcons5) * zz2s;
          zcons = zcons + (A(ztp25) - zmeltp2) * zz2s;
        }
        A(zdp) = A(zdp) + zcons2 * zcons / A(zlfdcp5);
        A(zlfdcp) = A(zlfdcp) - zcons2 * A(zdp5) * zcons / SQ(A(zlfdcp5));
      } else {
        zsfl[jl] = zsfl[jl] + zsfln[jl]; zsfln[jl] = 0.0;
        zrfl[jl] = zrfl[jl] + zrfln[jl]; zrfln[jl] = 0.0;
      }
    }

    /* :1407-1420 precipitation overlap: ZCOVPCLR / ZCOVPTOT adjoints are identically zero
       when LLO2 is false (only fed inside that block), so PCLC receives nothing here. */

    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :1424-1442 */
      zqc[jl] = 0.0;
      A(zqiwc) = A(zqiwc) + A(zcondi) * zqtmst;
      A(zi) = A(zi) - A(zcondi) * zqtmst;
      A(zcondi) = 0.0;
      A(zqlwc) = A(zqlwc) + A(zcondl) * zqtmst;
      A(zl) = A(zl) - A(zcondl) * zqtmst;
      A(zcondl) = 0.0;
      zqc[jl] = zqc[jl] + A(zqiwc) * (1.0 - A(zfwat5));
      zfwat[jl] = zfwat[jl] - A(zqiwc) * A(zqc35);
      A(zqiwc) = 0.0;
      zqc[jl] = zqc[jl] + A(zqlwc) * A(zfwat5);
      zfwat[jl] = zfwat[jl] + A(zqlwc) * A(zqc35);
      A(zqlwc) = 0.0;
    }

    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :1446-1496 subsidence */
      double zdqc = 0.0, zdqsdz = 0.0, dtdzmo = 0.0, zldcp = 0.0, zrodqsdp = 0.0, zrho = 0.0;
      zfoeew[jl] = 0.0;
      zdqc = zdqc - zqc[jl];
      if (A(llo3)) {
        if (lregcl) zdqc = zdqc * 0.1;
        zdqsdz = zdqsdz + zdqc * ptsphy * (A(pmfu5) + A(pmfd5)) * A(zfac4);
        A(pmfu) = A(pmfu) + zdqc * ptsphy * A(zdqsdz5) * A(zfac4);
        A(pmfd) = A(pmfd) + zdqc * ptsphy * A(zdqsdz5) * A(zfac4);
        zrho = zrho - zdqc * A(zdqc5) * A(zfac4);
      } else {
        zqc[jl] = zqc[jl] + zdqc;
      }
      zdqc = 0.0;
      dtdzmo = dtdzmo + zdqsdz * A(zdqsdtemp5);
      zdqsdtemp[jl] = zdqsdtemp[jl] + zdqsdz * A(dtdzmo5);
      zrodqsdp = zrodqsdp - zdqsdz * P->rg;
      zdqsdz = 0.0;
      zldcp = zldcp - dtdzmo * (P->rg * A(zrodqsdp5) + A(dtdzmo5) * A(zdqsdtemp5)) * A(zfac3);
      zrodqsdp = zrodqsdp - dtdzmo * P->rg * A(zldcp5) * A(zfac3);
      zdqsdtemp[jl] = zdqsdtemp[jl] - dtdzmo * A(dtdzmo5) * A(zldcp5) * A(zfac3);
      dtdzmo = 0.0;
      zfwat[jl] = zfwat[jl] + zldcp * (A(zlvdcp5) - A(zlsdcp5));
      A(zlvdcp) = A(zlvdcp) + zldcp * A(zfwat5);
      A(zlsdcp) = A(zlsdcp) + zldcp * (1.0 - A(zfwat5));
      zldcp = 0.0;
      zrho = zrho - zrodqsdp * A(pqs5) * A(zfac2);
      A(pqs) = A(pqs) - zrodqsdp * A(zrho5) * A(zfac2);
      A(papp1) = A(papp1) + zrodqsdp * A(zrho5) * A(pqs5) * SQ(A(zfac2));
      zfoeew[jl] = zfoeew[jl] - zrodqsdp * A(zrho5) * A(pqs5) * P->retv * SQ(A(zfac2));
      zrodqsdp = 0.0;
      A(papp1) = A(papp1) + zrho * A(zfac1);
      A(ztp1) = A(ztp1) - zrho * A(papp15) / A(ztp25) * A(zfac1);
      zrho = 0.0;
      (void)zdqc; (void)zdqsdz; (void)dtdzmo; (void)zldcp; (void)zrodqsdp; (void)zrho;
    }

    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :1500-1527 convective component */
      int llo1;
      if (jk < klev - 1) llo1 = A(zlude5) >= P->rlmin && plu5[IX(jl, jk + 1)] >= zeps2;
      else llo1 = 0;
      if (llo1) {
        const double plu5n = plu5[IX(jl, jk + 1)];
        A(zlude) = A(zlude) + zqc[jl];
        A(zlude) = A(zlude) + ((1.0 - A(zclc5)) / plu5n) * exp(-A(zlude5) / plu5n) * A(pclc);
        plu[IX(jl, jk + 1)] = plu[IX(jl, jk + 1)] -
            ((1.0 - A(zclc5)) * A(zlude5) / SQ(plu5n)) * exp(-A(zlude5) / plu5n) * A(pclc);
        A(pclc) = A(pclc) * (1.0 - (1.0 - exp(-A(zlude5) / plu5n)));
      }
      A(plude) = A(plude) + ptsphy * A(zgdp5) * A(zlude);
      zgdp[jl] = zgdp[jl] + ptsphy * A(plude5) * A(zlude);
      A(zlude) = 0.0;
      paphp1[IX(jl, jk + 1)] = paphp1[IX(jl, jk + 1)] -
          P->rg * zgdp[jl] / SQ(paphp15[IX(jl, jk + 1)] - paphp15[IX(jl, jk)]);
      paphp1[IX(jl, jk)] = paphp1[IX(jl, jk)] +
          P->rg * zgdp[jl] / SQ(paphp15[IX(jl, jk + 1)] - paphp15[IX(jl, jk)]);
      zgdp[jl] = 0.0;
    }

    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :1531-1583 uniform distribution */
      double zqpd = 0.0, zqcd = 0.0, zqt = 0.0;
      zqsat[jl] = 0.0;
      if (A(zqt5) <= A(zqcrit5)) {
        zqc[jl] = 0.0; A(pclc) = 0.0;
      } else if (A(zqt5) >= A(zqsat5)) {
        zqsat[jl] = zqsat[jl] + (1.0 - zscalm[jk]) * zqc[jl];
        zqcrit[jl] = zqcrit[jl] - (1.0 - zscalm[jk]) * zqc[jl];
        zqc[jl] = 0.0; A(pclc) = 0.0;
      } else {
        zqpd = zqpd + zscalm[jk] * zqc[jl] * SQ(A(zclc5));
        zqcd = zqcd + (1.0 - zscalm[jk]) * zqc[jl] * SQ(A(zclc5));
        A(pclc) = A(pclc) + (zscalm[jk] * A(zqpd5) + (1.0 - zscalm[jk]) * A(zqcd5)) * 2.0 *
                                A(zclc5) * zqc[jl];
        zqc[jl] = 0.0;
        if (lregcl) {
          double zrat = A(zqpd5) / A(zqcd5);
          double zyyy = dmin(0.3, 3.5 * sqrt(zrat * CUBE(1.0 - zscalm[jk] * (1.0 - zrat))) /
                                      (1.0 - zscalm[jk]));
          A(pclc) = zyyy * A(pclc);
        }
        const double den = A(zqcd5) - zscalm[jk] * (A(zqt5) - A(zqcrit5));
        zqpd = zqpd - (0.5 / A(zsqrt5)) * A(pclc) / den;
        zqcd = zqcd + (0.5 / A(zsqrt5)) * (A(zqpd5) * A(pclc)) / SQ(den);
        zqt = zqt - (0.5 / A(zsqrt5)) * (A(zqpd5) * zscalm[jk] * A(pclc)) / SQ(den);
        zqcrit[jl] = zqcrit[jl] + (0.5 / A(zsqrt5)) * (A(zqpd5) * zscalm[jk] * A(pclc)) / SQ(den);
        A(pclc) = 0.0;
        zqsat[jl] = zqsat[jl] + zqcd;
        zqcrit[jl] = zqcrit[jl] - zqcd;
        zqsat[jl] = zqsat[jl] + zqpd;
        zqt = zqt - zqpd;
      }
      A(zqp1) = A(zqp1) + zqt;
      A(zl) = A(zl) + zqt;
      A(zi) = A(zi) + zqt;
    }

    for (int jl = kidia - 1; jl < kfdia; ++jl) {   /* :1585-1666 */
      double zoealfaw = 0.0, zesdp = 0.0, zfacw = 0.0, zfaci = 0.0, zfac = 0.0, zcor = 0.0,
             zsupsat = 0.0;
      zqsat[jl] = zqsat[jl] + zqcrit[jl] * A(zcrh2);
      zqcrit[jl] = 0.0;
      A(pqs) = A(pqs) + zqsat[jl] * A(zsupsat5);
      zsupsat = zsupsat + zqsat[jl] * A(pqs5);
      zqsat[jl] = 0.0;
      if (A(ztp25) < P->rtice) A(ztp1) = A(ztp1) - zsupsat * 3.e-03;
      zsupsat = 0.0;
      /* :1610-1615 ZQLIM adjoint is identically zero (dead) */
      /* :1621-1622 ZCORQS adjoint is identically zero (dead): ZDQSDTEMP += ZCONS3*0 */
      zdqsdtemp[jl] = zdqsdtemp[jl] + zcons3 * 0.0;
      A(pqs) = A(pqs) + A(zfac5) * A(zcor5) * zdqsdtemp[jl];
      zcor = zcor + A(zfac5) * A(pqs5) * zdqsdtemp[jl];
      zfac = zfac + A(zcor5) * A(pqs5) * zdqsdtemp[jl];
      zdqsdtemp[jl] = 0.0;
      zesdp = zesdp + P->retv * zcor / SQ(1.0 - P->retv * A(zesdp5));
      zfacw = zfacw + A(zfwat5) * zfac;
      zfwat[jl] = zfwat[jl] + A(zfacw5) * zfac;
      zfaci = zfaci + (1.0 - A(zfwat5)) * zfac;
      zfwat[jl] = zfwat[jl] - A(zfaci5) * zfac;
      A(ztp1) = A(ztp1) - 2.0 * P->r5ies * zfaci / CUBE(A(ztp25) - P->r4ies);
      A(ztp1) = A(ztp1) - 2.0 * P->r5les * zfacw / CUBE(A(ztp25) - P->r4les);
      if (A(zesdp15) > zqmax) zesdp = 0.0;
      zfoeew[jl] = zfoeew[jl] + zesdp / A(papp15);
      A(papp1) = A(papp1) - zesdp * A(zfoeew5) / SQ(A(papp15));
      double z3es, z4es;
      if (A(ztp25) < P->rtt) { z3es = P->r3ies; z4es = P->r4ies; }
      else { z3es = P->r3les; z4es = P->r4les; }
      A(ztp1) = A(ztp1) + z3es * (P->rtt - z4es) * zfoeew[jl] * A(zfoeew5) / SQ(A(ztp25) - z4es);
      if (A(ztp25) < P->rtt) zoealfaw = zoealfaw + zfwat[jl];
      zfwat[jl] = 0.0;
      A(ztp1) = A(ztp1) + 0.545 * 0.17 * zoealfaw / SQ(cosh(0.17 * (A(ztp25) - P->rlptrc)));
      (void)zsupsat;
    }
  } /* jk */

  /* :1677-1692 */
  for (int jl = kidia - 1; jl < kfdia; ++jl) { pfplsn[IX(jl, 0)] = 0.0; pfplsl[IX(jl, 0)] = 0.0; }
  for (int jk = 0; jk < klev; ++jk)
    for (int jl = kidia - 1; jl < kfdia; ++jl) { A(pcovptot) = 0.0; A(pclc) = 0.0; }

  /* :1701-1717 thermodynamic constants */
  for (int jk = klev - 1; jk >= 0; --jk)
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      double zzz = 0.0;
      zzz = zzz + P->rlvtt * A(zlvdcp); A(zlvdcp) = 0.0;
      zzz = zzz + P->rlstt * A(zlsdcp); A(zlsdcp) = 0.0;
      zzz = zzz + P->rlmlt * A(zlfdcp); A(zlfdcp) = 0.0;
      A(zqp1) = A(zqp1) - zzz * P->rcpd * P->rvtmp2 / SQ(P->rcpd + P->rcpd * P->rvtmp2 * A(zqp15));
      paphp1[IX(jl, jk + 1)] = paphp1[IX(jl, jk + 1)] + A(zdp);
      paphp1[IX(jl, jk)] = paphp1[IX(jl, jk)] - A(zdp);
      A(zdp) = 0.0;
    }
  /* :1721-1740 first guess */
  for (int jk = klev - 1; jk >= 0; --jk)
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      A(pi) = A(pi) + A(zi);
      A(pgteni) = A(pgteni) + ptsphy * A(zi);
      A(zi) = 0.0;
      A(pl) = A(pl) + A(zl);
      A(pgtenl) = A(pgtenl) + ptsphy * A(zl);
      A(zl) = 0.0;
      A(pqm1) = A(pqm1) + A(zqp1);
      A(pgtenq) = A(pgtenq) + ptsphy * A(zqp1);
      A(psupsat) = ptsphy * A(zqp1);   /* :1733 assignment with PTSPHY, as in the reference */
      A(zqp1) = 0.0;
      A(ptm1) = A(ptm1) + A(ztp1);
      A(pgtent) = A(pgtent) + ptsphy * A(ztp1);
      A(ztp1) = 0.0;
    }

  free(w);
  free(llo3);
  return 0;
}
