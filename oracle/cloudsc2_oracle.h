/*
 * cloudsc2_oracle.h -- CPU restatement (plain C, FP64) of the reference's CLOUDSC2 NL/TL/AD path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity status: PINNED.
 *  (1) Against the reference's own FORTRAN TEXT: `make -C oracle ref` transliterates satur.F90, cuadjtqs.F90,
 *      cloudsc2.F90, cuadjtqstl.F90, cloudsc2tl.F90, cuadjtqsad.F90, cloudsc2ad.F90 from /root/reference/src,
 *      where they lie, into oracle/_ref/libcloudsc2_ref.so (oracle/f90toc.py; statement-by-statement C, same
 *      expression trees, same flags); tests/test_oracle_ref.py asserts this restatement == that library to
 *      1e-13 max|field| (measured: exactly 0) for NL, TL and AD, LREGCL off and on, RVTMP2 zero and
 *      non-zero, two atmospheres, and identical Taylor / adjoint self-test numbers.
 *  (2) Against the reference's own importable Python kernel (src/cloudsc2_nl_gt4py/cloudsc2_py.py) through
 *      the golden vectors in tests/golden/: NL values on two synthetic atmospheres and on edge-case columns
 *      sitting on the scheme's branch thresholds (make_golden.py, make_golden_edge.py); TL and AD against
 *      central finite differences of that kernel along the drivers' direction (make_golden_tl.py) and along
 *      16 single-input + 2 random directions (make_golden_tl_dirs.py).
 *  (3) By the reference's own known-answer properties (Taylor test, adjoint dot-product test).
 * Not available here: the reference's Fortran BINARIES on the real input.h5 (no Fortran compiler, no HDF5,
 * config-files/input.h5 absent) -- what a compiler could change relative to the source text (last bits).
 *
 * Statically dead code of the reference that is NOT restated: the precipitation-evaporation
 * branch guarded by LLO2 = ... .AND. (LEVAPLS2 .OR. LDRAIN1D) (cloudsc2.F90:556-591,
 * cloudsc2tl.F90:845-943, cloudsc2ad.F90:724-773,1152-1267): LEVAPLS2 is forced .FALSE. by all
 * three programs (cloudsc2_{nl,tl,ad}/dwarf_cloudsc.F90:105) and LDRAIN1D is a hard-coded .FALSE. in
 * all three drivers (cloudsc_driver*_mod.F90 "LOGICAL :: LDRAIN1D = .FALSE.").  This hand
 * restatement refuses (returns -1) if either switch is set; the transliterated reference in
 * oracle/_ref runs the branch (tests/test_oracle_ref.py::test_transliterated_reference_covers_...).  Likewise only LPHYLIN=.TRUE. (forced at
 * dwarf_cloudsc.F90:107) and KCALL==0 of CUADJTQS* are restated.
 */
#ifndef CLOUDSC2_ORACLE_H
#define CLOUDSC2_ORACLE_H
#include "../include/cloudsc2_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Per-block kernels: explicit-shape (KLON,KLEV) arrays, column-major, 1-based KIDIA..KFDIA. */

/* satur.F90:10, LDPHYLIN branch :106-123 */
void orc_satur(const cloudsc2_params *P, int kidia, int kfdia, int klon, int klev,
               const double *paprsf, const double *pt, double *pqsat);

/* cuadjtqs.F90:10, phase select :118-130, KCALL==0 :212-244; level kk is 0-based */
void orc_cuadjtqs(const cloudsc2_params *P, int kidia, int kfdia, int klon, int kk,
                  const double *psp, double *pt, double *pq);

/* cloudsc2.F90:10-741 */
int orc_cloudsc2(const cloudsc2_params *P, const double *ceta, int kidia, int kfdia, int klon,
                 int klev, double ptsphy, const double *paphp1, const double *papp1,
                 const double *pqm1, const double *pqs, const double *ptm1, const double *pl,
                 const double *pi, const double *plude, const double *plu, const double *pmfu,
                 const double *pmfd, double *ptent, const double *pgtent, double *ptenq,
                 const double *pgtenq, double *ptenl, const double *pgtenl, double *pteni,
                 const double *pgteni, const double *psupsat, double *pclc, double *pfplsl,
                 double *pfplsn, double *pfhpsl, double *pfhpsn, double *pcovptot);

/* Bundles of the 16 input-like and 10 output-like (KLON,KLEV[+1]) arrays of CLOUDSC2TL/AD. */
typedef struct orc_in16 {
  double *paphp1, *papp1, *pqm1, *pqs, *ptm1, *pl, *pi, *plude, *plu, *pmfu, *pmfd;
  double *pgtent, *pgtenq, *pgtenl, *pgteni, *psupsat;
} orc_in16;
typedef struct orc_out10 {
  double *ptent, *ptenq, *ptenl, *pteni, *pclc, *pfplsl, *pfplsn, *pfhpsl, *pfhpsn, *pcovptot;
} orc_out10;

/* cuadjtqstl.F90:10, KCALL==0 :333-405 */
void orc_cuadjtqstl(const cloudsc2_params *P, int kidia, int kfdia, int klon, int kk,
                    const double *psp5, double *pt5, double *pq5,
                    const double *psp, double *pt, double *pq);
/* cloudsc2tl.F90:10-1119 : x5 = trajectory (in5 read, out5 written), x = perturbation */
int orc_cloudsc2tl(const cloudsc2_params *P, const double *ceta, int kidia, int kfdia, int klon,
                   int klev, double ptsphy, const orc_in16 *in5, const orc_out10 *out5,
                   const orc_in16 *in, const orc_out10 *out);

/* cuadjtqsad.F90:10, trajectory :314-367, adjoint :542-641 */
void orc_cuadjtqsad(const cloudsc2_params *P, int kidia, int kfdia, int klon, int kk,
                    const double *psp5, double *pt5, double *pq5,
                    double *psp, double *pt, double *pq);
/* cloudsc2ad.F90:10-1746 : in5/out5 trajectory; `in` = input adjoints (accumulated),
 * `out` = output adjoints (consumed and zeroed) */
int orc_cloudsc2ad(const cloudsc2_params *P, const double *ceta, int kidia, int kfdia, int klon,
                   int klev, double ptsphy, const orc_in16 *in5, const orc_out10 *out5,
                   const orc_in16 *in, const orc_out10 *out);

/* Drivers on the blocked arrays (same layout as the C ABI). nthreads = NUMOMP. */

/* cloudsc_driver_mod.F90:22-125 ; elapsed_s = wall time of the block loop (:71,121) */
int orc_driver_nl(const cloudsc2_params *P, const double *ceta, int numomp, int nproma, int klev,
                  int ngptot, double ptsphy, const cloudsc2_fields *F, double *elapsed_s);
/* cloudsc_driver_tl_mod.F90:33-254 ; znormg[10] raw ratios, ratios_blk [nblocks][10] or NULL.
 * returns 3 when a block is degenerate (reference STOPs). */
int orc_driver_tl(const cloudsc2_params *P, const double *ceta, int numomp, int nproma, int klev,
                  int ngptot, double ptsphy, const cloudsc2_fields *F, double znormg[10],
                  double *ratios_blk, double *elapsed_s);
/* cloudsc_driver_ad_mod.F90:22-271 ; norms_col [ngptot][3] or NULL */
int orc_driver_ad(const cloudsc2_params *P, const double *ceta, int numomp, int nproma, int klev,
                  int ngptot, double ptsphy, const cloudsc2_fields *F, double *znormg,
                  double *norms_col, double *elapsed_s);

/* Timing-only drivers for the CPU baseline: the block loop calling SATUR + CLOUDSC2TL (resp.
 * SATUR + CLOUDSC2AD) once per block with dx = 0.01 x, writing increments into per-thread
 * scratch (the reference has no such executable; README.md:57,63). */
int orc_bench_tl(const cloudsc2_params *P, const double *ceta, int numomp, int nproma, int klev,
                 int ngptot, double ptsphy, const cloudsc2_fields *F, double *elapsed_s);
int orc_bench_ad(const cloudsc2_params *P, const double *ceta, int numomp, int nproma, int klev,
                 int ngptot, double ptsphy, const cloudsc2_fields *F, double *elapsed_s);

/* verdict logic, cloudsc_driver_tl_mod.F90:273-311 / cloudsc_driver_ad_mod.F90:286-294 */
int orc_taylor_verdict(const double znormg[10], int *istart_out);
int orc_adjoint_verdict(double znormg);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
