/*
 * cloudsc2_nl.c -- oracle (TEST INFRASTRUCTURE, see cloudsc2_oracle.h): plain-C restatement of
 *   SATUR     reference src/cloudsc2_nl/satur.F90:106-123  (LDPHYLIN branch)
 *   CUADJTQS  reference src/cloudsc2_nl/cuadjtqs.F90:118-130, 212-244 (KCALL==0)
 *   CLOUDSC2  reference src/cloudsc2_nl/cloudsc2.F90:235-735
 * Statement order, branch conditions and operator association follow the Fortran so that the
 * two agree to rounding.  Arrays are (KLON,KLEV) column-major; IX(jl,jk) is 0-based.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "cloudsc2_oracle.h"

#define IX(jl, jk) ((size_t)(jk) * (size_t)klon + (size_t)(jl))

static inline double dmin(double a, double b) { return a < b ? a : b; }
static inline double dmax(double a, double b) { return a > b ? a : b; }

/* fcttre.func.h:73-75 */
static inline double foealfa(const cloudsc2_params *P, double ptare) {
  double x = (dmax(P->rtice, dmin(P->rtwat, ptare)) - P->rtice) * P->rtwat_rtice_r;
  return dmin(1.0, x * x);
}

/* satur.F90:106-123 */
void orc_satur(const cloudsc2_params *P, int kidia, int kfdia, int klon, int klev,
               const double *paprsf, const double *pt, double *pqsat) {
  const double zqmax = 0.5;
  for (int jk = 0; jk < klev; ++jk) {
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      double ztarg = pt[IX(jl, jk)];
      double zalfa = foealfa(P, ztarg);
      double zfoeewl = P->r2es * exp(P->r3les * (ztarg - P->rtt) / (ztarg - P->r4les));
      double zfoeewi = P->r2es * exp(P->r3ies * (ztarg - P->rtt) / (ztarg - P->r4ies));
      double zfoeew = zalfa * zfoeewl + (1.0 - zalfa) * zfoeewi;
      double zqs = zfoeew / paprsf[IX(jl, jk)];
      if (zqs > zqmax) zqs = zqmax;
      double zcor = 1.0 / (1.0 - P->retv * zqs);
      pqsat[IX(jl, jk)] = zqs * zcor;
    }
  }
}

/* cuadjtqs.F90:118-130 (phase selection on the incoming T) and :212-244 (KCALL==0) */
void orc_cuadjtqs(const cloudsc2_params *P, int kidia, int kfdia, int klon, int kk,
                  const double *psp, double *pt, double *pq) {
  const double zqmax = 0.5;
  for (int jl = kidia - 1; jl < kfdia; ++jl) {
    double z3es, z4es, z5alcp, zaldcp;
    if (pt[IX(jl, kk)] > P->rtt) {
      z3es = P->r3les; z4es = P->r4les; z5alcp = P->r5alvcp; zaldcp = P->ralvdcp;
    } else {
      z3es = P->r3ies; z4es = P->r4ies; z5alcp = P->r5alscp; zaldcp = P->ralsdcp;
    }
    double zqp = 1.0 / psp[jl];
    for (int it = 0; it < 2; ++it) {   /* the two textually identical iterations */
      double ztarg = pt[IX(jl, kk)];
      double zfoeew = P->r2es * exp(z3es * (ztarg - P->rtt) / (ztarg - z4es));
      double zqsat = zqp * zfoeew;
      if (zqsat > zqmax) zqsat = zqmax;
      double zcor = 1.0 / (1.0 - P->retv * zqsat);
      zqsat = zqsat * zcor;
      double z2s = z5alcp / ((ztarg - z4es) * (ztarg - z4es));
      double zcond1 = (pq[IX(jl, kk)] - zqsat) / (1.0 + zqsat * zcor * z2s);
      pt[IX(jl, kk)] = pt[IX(jl, kk)] + zaldcp * zcond1;
      pq[IX(jl, kk)] = pq[IX(jl, kk)] - zcond1;
    }
  }
}

/* cloudsc2.F90:10-741 */
int orc_cloudsc2(const cloudsc2_params *P, const double *ceta, int kidia, int kfdia, int klon,
                 int klev, double ptsphy, const double *paphp1, const double *papp1,
                 const double *pqm1, const double *pqs, const double *ptm1, const double *pl,
                 const double *pi, const double *plude, const double *plu, const double *pmfu,
                 const double *pmfd, double *ptent, const double *pgtent, double *ptenq,
                 const double *pgtenq, double *ptenl, const double *pgtenl, double *pteni,
                 const double *pgteni, const double *psupsat, double *pclc, double *pfplsl,
                 double *pfplsn, double *pfhpsl, double *pfhpsn, double *pcovptot) {
  if (P->levapls2 || P->ldrain1d || !P->lphylin) return -1; /* see cloudsc2_oracle.h */

  const double zscal = 0.9; /* :172 */
  const size_t n2 = (size_t)klon * (size_t)klev;
  /* :176-193 work arrays */
  double *w = (double *)malloc(sizeof(double) * (15 * n2 + (size_t)klev + 24 * (size_t)klon));
  if (!w) return -2;
  double *ztp1 = w, *zqp1 = ztp1 + n2, *zl = zqp1 + n2, *zi = zl + n2, *zlude = zi + n2;
  double *zqc = zlude + n2, *zqlwc = zqc + n2, *zqiwc = zqlwc + n2, *zdp = zqiwc + n2;
  double *zlsdcp = zdp + n2, *zlfdcp = zlsdcp + n2, *zlvdcp = zlfdcp + n2;
  double *zrfreeze = zlvdcp + n2, *zcondl = zrfreeze + n2, *zcondi = zcondl + n2;
  double *zscalm = zcondi + n2;
  double *zrfl = zscalm + klev, *zsfl = zrfl + klon, *zrfln = zsfl + klon, *zsfln = zrfln + klon;
  double *zgdp = zsfln + klon, *zdqdt = zgdp + klon, *zdtdt = zdqdt + klon;
  double *zdldt = zdtdt + klon, *zdidt = zdldt + klon, *zqcrit = zdidt + klon;
  double *zcovpclr = zqcrit + klon, *zcovptot = zcovpclr + klon, *zdqsdtemp = zcovptot + klon;
  double *zcorqs = zdqsdtemp + klon, *zqold = zcorqs + klon, *zpp = zqold + klon;
  double *zdq = zpp + klon, *zqlim = zdq + klon, *zqsat = zqlim + klon, *zfoeew = zqsat + klon;
  double *zfwat = zfoeew + klon, *ztrpaus = zfwat + klon;
  /* ZEVAPR/ZEVAPS are identically zero here (dead LLO2 branch), kept as literals below. */
  const double zevapr = 0.0, zevaps = 0.0;

  /* :235-244 */
  const double zckcodtl = 2.0 * P->rkconv * ptsphy;
  const double zckcodti = 5.0 * P->rkconv * ptsphy;
  const double zcons2 = 1.0 / (ptsphy * P->rg);
  const double zcons3 = P->rlvtt / P->rcpd;
  const double zmeltp2 = P->rtt + 2.0;
  const double zqtmst = 1.0 / ptsphy;
  const double zqmax = 0.5, zeps1 = 1.e-12, zeps2 = 1.e-10;

  /* :253-260 first guess */
  for (int jk = 0; jk < klev; ++jk)
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      ztp1[IX(jl, jk)] = ptm1[IX(jl, jk)] + ptsphy * pgtent[IX(jl, jk)];
      zqp1[IX(jl, jk)] = pqm1[IX(jl, jk)] + ptsphy * pgtenq[IX(jl, jk)] + psupsat[IX(jl, jk)];
      zl[IX(jl, jk)] = pl[IX(jl, jk)] + ptsphy * pgtenl[IX(jl, jk)];
      zi[IX(jl, jk)] = pi[IX(jl, jk)] + ptsphy * pgteni[IX(jl, jk)];
    }

  /* :262-279 */
  for (int jk = 0; jk < klev; ++jk) {
    zscalm[jk] = zscal * pow(dmax(ceta[jk] - 0.2, zeps1), 0.2);
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      zdp[IX(jl, jk)] = paphp1[IX(jl, jk + 1)] - paphp1[IX(jl, jk)];
      double zzz = 1.0 / (P->rcpd + P->rcpd * P->rvtmp2 * zqp1[IX(jl, jk)]);
      zlfdcp[IX(jl, jk)] = P->rlmlt * zzz;
      zlsdcp[IX(jl, jk)] = P->rlstt * zzz;
      zlvdcp[IX(jl, jk)] = P->rlvtt * zzz;
    }
  }

  /* :288-301 */
  for (int jk = 0; jk < klev; ++jk)
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      pclc[IX(jl, jk)] = 0.0;
      zqc[IX(jl, jk)] = 0.0;
      zqlwc[IX(jl, jk)] = 0.0;
      zqiwc[IX(jl, jk)] = 0.0;
      zrfreeze[IX(jl, jk)] = 0.0;
      zcondl[IX(jl, jk)] = 0.0;
      zcondi[IX(jl, jk)] = 0.0;
      pcovptot[IX(jl, jk)] = 0.0;
    }
  /* :305-312 */
  for (int jl = kidia - 1; jl < kfdia; ++jl) {
    zrfl[jl] = 0.0;
    zsfl[jl] = 0.0;
    pfplsl[IX(jl, 0)] = 0.0;
    pfplsn[IX(jl, 0)] = 0.0;
    zcovptot[jl] = 0.0;
    zcovpclr[jl] = 0.0;
  }
  /* :315-326 eta value at tropopause */
  for (int jl = kidia - 1; jl < kfdia; ++jl) ztrpaus[jl] = 0.1;
  for (int jk = 0; jk < klev - 1; ++jk)
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      int llo1 = ceta[jk] > 0.1 && ceta[jk] < 0.4 && ztp1[IX(jl, jk)] > ztp1[IX(jl, jk + 1)];
      if (llo1) ztrpaus[jl] = ceta[jk];
    }

  /* :339 main vertical loop */
  for (int jk = 0; jk < klev; ++jk) {
    /* :343-408 */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      double z3es, z4es, zesdp;
      /* LPHYLIN branch :349-364 */
      double zoealfaw = 0.545 * (tanh(0.17 * (ztp1[IX(jl, jk)] - P->rlptrc)) + 1.0);
      if (ztp1[IX(jl, jk)] < P->rtt) {
        zfwat[jl] = zoealfaw; z3es = P->r3ies; z4es = P->r4ies;
      } else {
        zfwat[jl] = 1.0; z3es = P->r3les; z4es = P->r4les;
      }
      zfoeew[jl] = P->r2es * exp(z3es * (ztp1[IX(jl, jk)] - P->rtt) / (ztp1[IX(jl, jk)] - z4es));
      zesdp = zfoeew[jl] / papp1[IX(jl, jk)];
      if (zesdp > zqmax) zesdp = zqmax;
      /* :370-375 */
      double tw = ztp1[IX(jl, jk)] - P->r4les, ti = ztp1[IX(jl, jk)] - P->r4ies;
      double zfacw = P->r5les / (tw * tw);
      double zfaci = P->r5ies / (ti * ti);
      double zfac = zfwat[jl] * zfacw + (1.0 - zfwat[jl]) * zfaci;
      double zcor = 1.0 / (1.0 - P->retv * zesdp);
      zdqsdtemp[jl] = zfac * zcor * pqs[IX(jl, jk)];
      zcorqs[jl] = 1.0 + zcons3 * zdqsdtemp[jl];
      /* :379-380 */
      zqlim[jl] = zqp1[IX(jl, jk)];
      if (zqp1[IX(jl, jk)] > pqs[IX(jl, jk)]) zqlim[jl] = pqs[IX(jl, jk)];
      /* :384-399 critical relative humidity */
      double zeta3 = ztrpaus[jl];
      double zrh1 = 1.0;
      double q = (zeta3 - 0.25) / 0.15;
      double zrh2 = 0.35 + 0.14 * (q * q) + 0.04 * dmin(zeta3 - 0.25, 0.0) / 0.15;
      double zrh3 = 1.0;
      double zdeta2 = 0.3;
      double zdeta1 = 0.09 + 0.16 * (0.4 - zeta3) / 0.3;
      double zcrh2 = 0.0;
      if (ceta[jk] < zeta3) {
        zcrh2 = zrh3;
      } else if (ceta[jk] >= zeta3 && ceta[jk] < (zeta3 + zdeta2)) {
        zcrh2 = zrh3 + (zrh2 - zrh3) * ((ceta[jk] - zeta3) / zdeta2);
      } else if (ceta[jk] >= (zeta3 + zdeta2) && ceta[jk] < (1.0 - zdeta1)) {
        zcrh2 = zrh2;
      } else if (ceta[jk] >= (1.0 - zdeta1)) {
        zcrh2 = zrh1 + (zrh2 - zrh1) * sqrt((1.0 - ceta[jk]) / zdeta1);
      }
      /* :401-407 */
      double zsupsat;
      if (ztp1[IX(jl, jk)] < P->rtice) zsupsat = 1.8 - 3.e-03 * ztp1[IX(jl, jk)];
      else zsupsat = 1.0;
      zqsat[jl] = pqs[IX(jl, jk)] * zsupsat;
      zqcrit[jl] = zcrh2 * zqsat[jl];
    }

    /* :412-427 uniform distribution of total water */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      double zqt = zqp1[IX(jl, jk)] + zl[IX(jl, jk)] + zi[IX(jl, jk)];
      if (zqt <= zqcrit[jl]) {
        pclc[IX(jl, jk)] = 0.0;
        zqc[IX(jl, jk)] = 0.0;
      } else if (zqt >= zqsat[jl]) {
        pclc[IX(jl, jk)] = 1.0;
        zqc[IX(jl, jk)] = (1.0 - zscalm[jk]) * (zqsat[jl] - zqcrit[jl]);
      } else {
        double zqpd = zqsat[jl] - zqt;
        double zqcd = zqsat[jl] - zqcrit[jl];
        pclc[IX(jl, jk)] = 1.0 - sqrt(zqpd / (zqcd - zscalm[jk] * (zqt - zqcrit[jl])));
        zqc[IX(jl, jk)] = (zscalm[jk] * zqpd + (1.0 - zscalm[jk]) * zqcd) *
                          (pclc[IX(jl, jk)] * pclc[IX(jl, jk)]);
      }
    }

    /* :431-444 convective component */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      zgdp[jl] = P->rg / (paphp1[IX(jl, jk + 1)] - paphp1[IX(jl, jk)]);
      zlude[IX(jl, jk)] = plude[IX(jl, jk)] * ptsphy * zgdp[jl];
      int llo1;
      if (jk < klev - 1) llo1 = zlude[IX(jl, jk)] >= P->rlmin && plu[IX(jl, jk + 1)] >= zeps2;
      else llo1 = 0;
      if (llo1) {
        pclc[IX(jl, jk)] = pclc[IX(jl, jk)] +
            (1.0 - pclc[IX(jl, jk)]) * (1.0 - exp(-zlude[IX(jl, jk)] / plu[IX(jl, jk + 1)]));
        zqc[IX(jl, jk)] = zqc[IX(jl, jk)] + zlude[IX(jl, jk)];
      }
    }

    /* :448-460 compensating subsidence */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      double zfac1 = 1.0 / (P->rd * ztp1[IX(jl, jk)]);
      double zrho = papp1[IX(jl, jk)] * zfac1;
      double zfac2 = 1.0 / (papp1[IX(jl, jk)] - P->retv * zfoeew[jl]);
      double zrodqsdp = -zrho * pqs[IX(jl, jk)] * zfac2;
      double zldcp = zfwat[jl] * zlvdcp[IX(jl, jk)] + (1.0 - zfwat[jl]) * zlsdcp[IX(jl, jk)];
      double zfac3 = 1.0 / (1.0 + zldcp * zdqsdtemp[jl]);
      double dtdzmo = P->rg * (1.0 / P->rcpd - zldcp * zrodqsdp) * zfac3;
      double zdqsdz = zdqsdtemp[jl] * dtdzmo - P->rg * zrodqsdp;
      double zfac4 = 1.0 / zrho;
      double zdqc = dmin(zdqsdz * (pmfu[IX(jl, jk)] + pmfd[IX(jl, jk)]) * ptsphy * zfac4,
                         zqc[IX(jl, jk)]);
      zqc[IX(jl, jk)] = zqc[IX(jl, jk)] - zdqc;
    }

    /* :464-469 */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      zqlwc[IX(jl, jk)] = zqc[IX(jl, jk)] * zfwat[jl];
      zqiwc[IX(jl, jk)] = zqc[IX(jl, jk)] * (1.0 - zfwat[jl]);
      zcondl[IX(jl, jk)] = (zqlwc[IX(jl, jk)] - zl[IX(jl, jk)]) * zqtmst;
      zcondi[IX(jl, jk)] = (zqiwc[IX(jl, jk)] - zi[IX(jl, jk)]) * zqtmst;
    }

    /* :475-481 precipitation overlap (maximum overlap) */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      if (pclc[IX(jl, jk)] > zcovptot[jl]) zcovptot[jl] = pclc[IX(jl, jk)];
      zcovpclr[jl] = zcovptot[jl] - pclc[IX(jl, jk)];
      zcovpclr[jl] = dmax(zcovpclr[jl], 0.0);
    }

    /* :487-498 melting of incoming snow */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      if (zsfl[jl] != 0.0) {
        double zcons = zcons2 * zdp[IX(jl, jk)] / zlfdcp[IX(jl, jk)];
        double zsnmlt = dmin(zsfl[jl], zcons * dmax(0.0, (ztp1[IX(jl, jk)] - zmeltp2)));
        zrfln[jl] = zrfl[jl] + zsnmlt;
        zsfln[jl] = zsfl[jl] - zsnmlt;
        ztp1[IX(jl, jk)] = ztp1[IX(jl, jk)] - zsnmlt / zcons;
      } else {
        zrfln[jl] = zrfl[jl];
        zsfln[jl] = zsfl[jl];
      }
    }

    /* :500-591 */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      double zprr, zprs;
      /* :504-516 rain production from cloud liquid water */
      if (pclc[IX(jl, jk)] > zeps2) {
        double zlcrit = P->rclcrit * 2.0;     /* LEVAPLS2.OR.LDRAIN1D is false */
        double zcldl = zqlwc[IX(jl, jk)] / pclc[IX(jl, jk)];
        double r = zcldl / zlcrit;
        double zd = zckcodtl * (1.0 - exp(-(r * r)));
        double zlnew = pclc[IX(jl, jk)] * zcldl * exp(-zd);
        zprr = zqlwc[IX(jl, jk)] - zlnew;
        zqlwc[IX(jl, jk)] = zqlwc[IX(jl, jk)] - zprr;
      } else {
        zprr = 0.0;
      }
      /* :520-534 snow production from cloud ice */
      if (pclc[IX(jl, jk)] > zeps2) {
        double zlcrit = P->rclcrit * 2.0;
        double zcldi = zqiwc[IX(jl, jk)] / pclc[IX(jl, jk)];
        double r = zcldi / zlcrit;
        double zd = zckcodti * exp(0.025 * (ztp1[IX(jl, jk)] - P->rtt)) * (1.0 - exp(-(r * r)));
        double zinew = pclc[IX(jl, jk)] * zcldi * exp(-zd);
        zprs = zqiwc[IX(jl, jk)] - zinew;
        zqiwc[IX(jl, jk)] = zqiwc[IX(jl, jk)] - zprs;
      } else {
        zprs = 0.0;
      }
      /* :538-552 new precipitation, rain fraction */
      double zdr = zcons2 * zdp[IX(jl, jk)] * (zprr + zprs);
      double zfwatr;
      if (ztp1[IX(jl, jk)] < P->rtt) {
        zrfreeze[IX(jl, jk)] = zcons2 * zdp[IX(jl, jk)] * zprr;
        zfwatr = 0.0;
      } else {
        zfwatr = 1.0;
      }
      double zrn = zfwatr * zdr;
      double zsn = (1.0 - zfwatr) * zdr;
      zrfln[jl] = zrfln[jl] + zrn;
      zsfln[jl] = zsfln[jl] + zsn;
      /* :556-591 precipitation evaporation: LLO2 is statically false (see header) */
    }

    /* :601-618 tendencies, first guess T and q */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      zdqdt[jl] = -(zcondl[IX(jl, jk)] + zcondi[IX(jl, jk)]) +
                  (plude[IX(jl, jk)] + zevapr + zevaps) * zgdp[jl];
      zdtdt[jl] = zlvdcp[IX(jl, jk)] * zcondl[IX(jl, jk)] + zlsdcp[IX(jl, jk)] * zcondi[IX(jl, jk)] -
                  (zlvdcp[IX(jl, jk)] * zevapr + zlsdcp[IX(jl, jk)] * zevaps +
                   plude[IX(jl, jk)] * (zfwat[jl] * zlvdcp[IX(jl, jk)] +
                                        (1.0 - zfwat[jl]) * zlsdcp[IX(jl, jk)]) -
                   (zlsdcp[IX(jl, jk)] - zlvdcp[IX(jl, jk)]) * zrfreeze[IX(jl, jk)]) *
                      zgdp[jl];
      ztp1[IX(jl, jk)] = ztp1[IX(jl, jk)] + ptsphy * zdtdt[jl];
      zqp1[IX(jl, jk)] = zqp1[IX(jl, jk)] + ptsphy * zdqdt[jl];
      zpp[jl] = papp1[IX(jl, jk)];
      zqold[jl] = zqp1[IX(jl, jk)];
    }

    /* :622-670 clipping of final qv: CUADJTQS (manually inlined in the NL reference) */
    orc_cuadjtqs(P, kidia, kfdia, klon, jk, zpp, ztp1, zqp1);

    /* :672-692 */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      zdq[jl] = dmax(0.0, zqold[jl] - zqp1[IX(jl, jk)]);
      double zdr2 = zcons2 * zdp[IX(jl, jk)] * zdq[jl];
      double zrfreeze2, zfwatr;
      if (ztp1[IX(jl, jk)] < P->rtt) {
        zrfreeze2 = zfwat[jl] * zdr2;
        zfwatr = 0.0;
      } else {
        zrfreeze2 = 0.0;
        zfwatr = 1.0;
      }
      double zrn = zfwatr * zdr2;
      double zsn = (1.0 - zfwatr) * zdr2;
      zcondl[IX(jl, jk)] = zcondl[IX(jl, jk)] + zfwatr * zdq[jl] * zqtmst;
      zcondi[IX(jl, jk)] = zcondi[IX(jl, jk)] + (1.0 - zfwatr) * zdq[jl] * zqtmst;
      zrfln[jl] = zrfln[jl] + zrn;
      zsfln[jl] = zsfln[jl] + zsn;
      zrfreeze[IX(jl, jk)] = zrfreeze[IX(jl, jk)] + zrfreeze2;
    }

    /* :694-716 final tendencies and fluxes */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      zdqdt[jl] = -(zcondl[IX(jl, jk)] + zcondi[IX(jl, jk)]) +
                  (plude[IX(jl, jk)] + zevapr + zevaps) * zgdp[jl];
      zdtdt[jl] = zlvdcp[IX(jl, jk)] * zcondl[IX(jl, jk)] + zlsdcp[IX(jl, jk)] * zcondi[IX(jl, jk)] -
                  (zlvdcp[IX(jl, jk)] * zevapr + zlsdcp[IX(jl, jk)] * zevaps +
                   plude[IX(jl, jk)] * (zfwat[jl] * zlvdcp[IX(jl, jk)] +
                                        (1.0 - zfwat[jl]) * zlsdcp[IX(jl, jk)]) -
                   (zlsdcp[IX(jl, jk)] - zlvdcp[IX(jl, jk)]) * zrfreeze[IX(jl, jk)]) *
                      zgdp[jl];
      zdldt[jl] = (zqlwc[IX(jl, jk)] - zl[IX(jl, jk)]) * zqtmst;
      zdidt[jl] = (zqiwc[IX(jl, jk)] - zi[IX(jl, jk)]) * zqtmst;
      ptenq[IX(jl, jk)] = zdqdt[jl];
      ptent[IX(jl, jk)] = zdtdt[jl];
      ptenl[IX(jl, jk)] = zdldt[jl];
      pteni[IX(jl, jk)] = zdidt[jl];
      pfplsl[IX(jl, jk + 1)] = zrfln[jl];
      pfplsn[IX(jl, jk + 1)] = zsfln[jl];
    }
    /* :720-723 */
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      zrfl[jl] = zrfln[jl];
      zsfl[jl] = zsfln[jl];
    }
  } /* jk */

  /* :730-735 enthalpy fluxes */
  for (int jk = 0; jk < klev + 1; ++jk)
    for (int jl = kidia - 1; jl < kfdia; ++jl) {
      pfhpsl[IX(jl, jk)] = -pfplsl[IX(jl, jk)] * P->rlvtt;
      pfhpsn[IX(jl, jk)] = -pfplsn[IX(jl, jk)] * P->rlstt;
    }

  free(w);
  return 0;
}
