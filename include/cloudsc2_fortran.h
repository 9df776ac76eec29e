/*
 * cloudsc2_fortran.h -- link-time substitutes for the reference's four kernel subroutines, in the
 * gfortran calling convention (lower-case name + '_', every argument by reference, explicit-shape
 * arrays as bare pointers to (KLON,KLEV[+1]) column-major storage, LOGICAL as 4-byte int).
 *
 * Leaving satur.F90, cloudsc2.F90, cloudsc2tl.F90, cloudsc2ad.F90 (and cuadjtqs*.F90) out of the
 * link and adding libcloudsc2_b200.so lets the reference's UNCHANGED drivers run on the GPU, one
 * NPROMA block per call (H2D, one launch, D2H: functional, not fast -- SURVEY 8b; the throughput
 * path is cloudsc2_gpu_nl & co. in cloudsc2_b200.h).  cloudsc2_gpu_init must have been called
 * (the Fortran kernels take their constants from modules, not arguments).  Restrictions, each
 * enforced with an abort in the style of ABOR1 (common/module/abor1.F90:10-14): KIDIA = KTDIA = 1,
 * LDRAIN1D = .FALSE., LDPHYLIN = .TRUE., KLEV as given to cloudsc2_gpu_init.  Columns
 * KFDIA+1..KLON keep the caller's values.  Thread-safe (calls are serialised), because the
 * reference calls the kernels from inside its OpenMP block loop (cloudsc_driver_mod.F90:73-111).
 */
#ifndef CLOUDSC2_FORTRAN_H
#define CLOUDSC2_FORTRAN_H
#ifdef __cplusplus
extern "C" {
#endif

/* SUBROUTINE SATUR(KIDIA,KFDIA,KLON,KTDIA,KLEV,LDPHYLIN,PAPRSF,PT,PQSAT,KFLAG)  satur.F90:10-11 */
void satur_(const int *kidia, const int *kfdia, const int *klon, const int *ktdia, const int *klev,
            const int *ldphylin, const double *paprsf, const double *pt, double *pqsat,
            const int *kflag);

/* SUBROUTINE CLOUDSC2(...)  cloudsc2.F90:10-18 */
void cloudsc2_(const int *kidia, const int *kfdia, const int *klon, const int *ktdia, const int *klev,
               const int *ldrain1d, const double *ptsphy,
               const double *paphp1, const double *papp1, const double *pqm1, const double *pqs,
               const double *ptm1, const double *pl, const double *pi, const double *plude,
               const double *plu, const double *pmfu, const double *pmfd,
               double *ptent, const double *pgtent, double *ptenq, const double *pgtenq,
               double *ptenl, const double *pgtenl, double *pteni, const double *pgteni,
               const double *psupsat, double *pclc, double *pfplsl, double *pfplsn, double *pfhpsl,
               double *pfhpsn, double *pcovptot);

/* The 26 trajectory ("5") arguments followed by the 26 perturbation / adjoint arguments, in the
 * order of cloudsc2tl.F90:10-24 and cloudsc2ad.F90:10-24. */
#define CLOUDSC2_TRAJ26                                                                          \
  const double *paphp15, const double *papp15, const double *pqm15, const double *pqs5,          \
      const double *ptm15, const double *pl5, const double *pi5, const double *plude5,           \
      const double *plu5, const double *pmfu5, const double *pmfd5, double *ptent5,              \
      const double *pgtent5, double *ptenq5, const double *pgtenq5, double *ptenl5,              \
      const double *pgtenl5, double *pteni5, const double *pgteni5, const double *psupsat5,      \
      double *pclc5, double *pfplsl5, double *pfplsn5, double *pfhpsl5, double *pfhpsn5,         \
      double *pcovptot5
#define CLOUDSC2_INCR26                                                                          \
  double *paphp1, double *papp1, double *pqm1, double *pqs, double *ptm1, double *pl, double *pi, \
      double *plude, double *plu, double *pmfu, double *pmfd, double *ptent, double *pgtent,     \
      double *ptenq, double *pgtenq, double *ptenl, double *pgtenl, double *pteni,               \
      double *pgteni, double *psupsat, double *pclc, double *pfplsl, double *pfplsn,             \
      double *pfhpsl, double *pfhpsn, double *pcovptot

/* SUBROUTINE CLOUDSC2TL(...)  cloudsc2tl.F90:10-24 : increments read, 10 output increments written */
void cloudsc2tl_(const int *kidia, const int *kfdia, const int *klon, const int *ktdia,
                 const int *klev, const int *ldrain1d, const double *ptsphy, CLOUDSC2_TRAJ26,
                 CLOUDSC2_INCR26);
/* SUBROUTINE CLOUDSC2AD(...)  cloudsc2ad.F90:10-24 : 10 output adjoints consumed and zeroed, 16 input
 * adjoints accumulated (PSUPSAT assigned, :1733) */
void cloudsc2ad_(const int *kidia, const int *kfdia, const int *klon, const int *ktdia,
                 const int *klev, const int *ldrain1d, const double *ptsphy, CLOUDSC2_TRAJ26,
                 CLOUDSC2_INCR26);

#ifdef __cplusplus
}
#endif
#endif
