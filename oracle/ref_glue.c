/*
 * ref_glue.c -- TEST INFRASTRUCTURE (see cloudsc2_oracle.h).  Hand-written glue around the C that
 * oracle/f90toc.py generates from the reference's Fortran kernels (oracle/_ref/ref_*.c).
 *
 * It re-exports the hand oracle's kernel interface (orc_satur / orc_cloudsc2 / orc_cloudsc2tl /
 * orc_cloudsc2ad / orc_cuadjtqs*) on top of the transliterated routines, so that
 *   oracle/_ref/libcloudsc2_ref.so = this glue + the transliterated kernels + a second compilation
 *                                    of the hand-written drivers (cloudsc2_drivers.c)
 * has the same exports as oracle/_build/liboracle.so and the tests can run every check twice.
 *
 * What the glue stands in for (none of it is kernel arithmetic):
 *  - the HDF5 loaders of the constant modules (YOMCST_LOAD_PARAMETERS yomcst.F90:167-177,
 *    YOETHF_LOAD_PARAMETERS yoethf.F90:79-99, YRECLDP_LOAD_PARAMETERS yoecldp.F90:242-370,
 *    YREPHLI_LOAD_PARAMETERS yoephli.F90:79-97, the CETA set-up cloudsc2_nl/dwarf_cloudsc.F90:100-102
 *    and the switches forced at :104-108): module variables are filled from a cloudsc2_params.
 *    Module variables the loaders do not set stay 0 like Fortran static storage (RVTMP2 comes from
 *    the params so that both values can be tested).
 *  - the call sites: SATUR as called at cloudsc_driver_mod.F90:91-92 (KTDIA=1, KFLAG=2,
 *    LDPHYLIN=YREPHLI%LPHYLIN), CLOUDSC2* as called at :95-109 (KTDIA=1, LDRAIN1D from the params),
 *    CUADJTQS* as called from the kernels (LDFLAG all true, KCALL=0).
 */
#include <stdlib.h>
#include <string.h>
#include "cloudsc2_oracle.h"
#include "_ref/ref_modules.h"
#include "_ref/ref_protos.h"

static cloudsc2_params g_p;
static double *g_ceta = NULL;
static int g_klev = -1, g_set = 0;

/* Fill the module variables; re-done only when params / CETA change (OpenMP block loops call this
 * from every thread with identical values). */
static void ref_setup(const cloudsc2_params *P, const double *ceta, int klev) {
#pragma omp critical(ref_setup)
  {
    int same = g_set && memcmp(&g_p, P, sizeof g_p) == 0 &&
               (!ceta || (g_klev == klev && memcmp(g_ceta, ceta, sizeof(double) * (size_t)klev) == 0));
    if (!same) {
      g_p = *P;
      g_set = 1;
      RG = P->rg; RD = P->rd; RCPD = P->rcpd; RETV = P->retv;
      RLVTT = P->rlvtt; RLSTT = P->rlstt; RLMLT = P->rlmlt; RTT = P->rtt;
      R2ES = P->r2es; R3LES = P->r3les; R3IES = P->r3ies; R4LES = P->r4les; R4IES = P->r4ies;
      R5LES = P->r5les; R5IES = P->r5ies; R5ALVCP = P->r5alvcp; R5ALSCP = P->r5alscp;
      RALVDCP = P->ralvdcp; RALSDCP = P->ralsdcp; RTWAT = P->rtwat; RTICE = P->rtice;
      RTWAT_RTICE_R = P->rtwat_rtice_r; RVTMP2 = P->rvtmp2;
      YRECLDP.RCLCRIT = P->rclcrit; YRECLDP.RKCONV = P->rkconv; YRECLDP.RLMIN = P->rlmin;
      YRECLDP.RPECONS = P->rpecons;
      YREPHLI.RLPTRC = P->rlptrc; YREPHLI.LPHYLIN = P->lphylin;
      YRPHNC.LEVAPLS2 = P->levapls2;
      YRNCL.LREGCL = P->lregcl;
      if (ceta) {
        double *c = (double *)malloc(sizeof(double) * (size_t)klev);
        memcpy(c, ceta, sizeof(double) * (size_t)klev);
        YRECLD.CETA = c;          /* the previous vector is leaked on purpose: another thread may */
        g_ceta = c;               /* still be reading it                                          */
        g_klev = klev;
      }
    }
  }
}

static int *all_true(int n) {
  int *f = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) f[i] = 1;
  return f;
}

void orc_satur(const cloudsc2_params *P, int kidia, int kfdia, int klon, int klev,
               const double *paprsf, const double *pt, double *pqsat) {
  ref_setup(P, NULL, klev);
  int ktdia = 1, ldphylin = P->lphylin, kflag = 2;
  ref_satur(&kidia, &kfdia, &klon, &ktdia, &klev, &ldphylin, (double *)paprsf, (double *)pt, pqsat, &kflag);
}

/* kk is 0-based like the hand oracle's; the arrays are (KLON, >= kk+1) */
void orc_cuadjtqs(const cloudsc2_params *P, int kidia, int kfdia, int klon, int kk,
                  const double *psp, double *pt, double *pq) {
  ref_setup(P, NULL, 0);
  int klev = kk + 1, k1 = kk + 1, kcall = 0, *f = all_true(klon);
  ref_cuadjtqs(&kidia, &kfdia, &klon, &klev, &k1, (double *)psp, pt, pq, f, &kcall);
  free(f);
}

void orc_cuadjtqstl(const cloudsc2_params *P, int kidia, int kfdia, int klon, int kk,
                    const double *psp5, double *pt5, double *pq5,
                    const double *psp, double *pt, double *pq) {
  ref_setup(P, NULL, 0);
  int klev = kk + 1, k1 = kk + 1, kcall = 0, *f = all_true(klon);
  ref_cuadjtqstl(&kidia, &kfdia, &klon, &klev, &k1, (double *)psp5, pt5, pq5, (double *)psp, pt, pq, f, &kcall);
  free(f);
}

void orc_cuadjtqsad(const cloudsc2_params *P, int kidia, int kfdia, int klon, int kk,
                    const double *psp5, double *pt5, double *pq5,
                    double *psp, double *pt, double *pq) {
  ref_setup(P, NULL, 0);
  int klev = kk + 1, k1 = kk + 1, kcall = 0, *f = all_true(klon);
  ref_cuadjtqsad(&kidia, &kfdia, &klon, &klev, &k1, (double *)psp5, pt5, pq5, psp, pt, pq, f, &kcall);
  free(f);
}

int orc_cloudsc2(const cloudsc2_params *P, const double *ceta, int kidia, int kfdia, int klon,
                 int klev, double ptsphy, const double *paphp1, const double *papp1,
                 const double *pqm1, const double *pqs, const double *ptm1, const double *pl,
                 const double *pi, const double *plude, const double *plu, const double *pmfu,
                 const double *pmfd, double *ptent, const double *pgtent, double *ptenq,
                 const double *pgtenq, double *ptenl, const double *pgtenl, double *pteni,
                 const double *pgteni, const double *psupsat, double *pclc, double *pfplsl,
                 double *pfplsn, double *pfhpsl, double *pfhpsn, double *pcovptot) {
  ref_setup(P, ceta, klev);
  int ktdia = 1, ldrain1d = P->ldrain1d;
#define CC(x) ((double *)(x))
  ref_cloudsc2(&kidia, &kfdia, &klon, &ktdia, &klev, &ldrain1d, &ptsphy, CC(paphp1), CC(papp1), CC(pqm1),
               CC(pqs), CC(ptm1), CC(pl), CC(pi), CC(plude), CC(plu), CC(pmfu), CC(pmfd), ptent,
               CC(pgtent), ptenq, CC(pgtenq), ptenl, CC(pgtenl), pteni, CC(pgteni), CC(psupsat), pclc,
               pfplsl, pfplsn, pfhpsl, pfhpsn, pcovptot);
  return 0;
}

/* argument order of cloudsc2tl.F90:10-24 and cloudsc2ad.F90:10-24 (identical) */
#define TLAD_ARGS(a, b)                                                                              \
  (a)->paphp1, (a)->papp1, (a)->pqm1, (a)->pqs, (a)->ptm1, (a)->pl, (a)->pi, (a)->plude, (a)->plu,   \
      (a)->pmfu, (a)->pmfd, (b)->ptent, (a)->pgtent, (b)->ptenq, (a)->pgtenq, (b)->ptenl,            \
      (a)->pgtenl, (b)->pteni, (a)->pgteni, (a)->psupsat, (b)->pclc, (b)->pfplsl, (b)->pfplsn,       \
      (b)->pfhpsl, (b)->pfhpsn, (b)->pcovptot

int orc_cloudsc2tl(const cloudsc2_params *P, const double *ceta, int kidia, int kfdia, int klon,
                   int klev, double ptsphy, const orc_in16 *in5, const orc_out10 *out5,
                   const orc_in16 *in, const orc_out10 *out) {
  ref_setup(P, ceta, klev);
  int ktdia = 1, ldrain1d = P->ldrain1d;
  ref_cloudsc2tl(&kidia, &kfdia, &klon, &ktdia, &klev, &ldrain1d, &ptsphy, TLAD_ARGS(in5, out5),
                 TLAD_ARGS(in, out));
  return 0;
}

int orc_cloudsc2ad(const cloudsc2_params *P, const double *ceta, int kidia, int kfdia, int klon,
                   int klev, double ptsphy, const orc_in16 *in5, const orc_out10 *out5,
                   const orc_in16 *in, const orc_out10 *out) {
  ref_setup(P, ceta, klev);
  int ktdia = 1, ldrain1d = P->ldrain1d;
  ref_cloudsc2ad(&kidia, &kfdia, &klon, &ktdia, &klev, &ldrain1d, &ptsphy, TLAD_ARGS(in5, out5),
                 TLAD_ARGS(in, out));
  return 0;
}

/* lets a test tell the two libraries apart */
const char *orc_flavour(void) { return "f90toc transliteration of the reference Fortran"; }
