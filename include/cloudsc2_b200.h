/*
 * cloudsc2_b200.h -- C ABI of the B200-native CLOUDSC2 NL / TL / AD column physics.
 *
 * This is the drop-in boundary between the (unchanged) Fortran host of
 * ecmwf-ifs/dwarf-p-cloudsc2-tl-ad and the CUDA (sm_100a) implementation.  Every entry
 * point replaces one reference interface, cited as (file:line relative to the
 * reference's src/ directory).  Plain C types only: pointers, ints, doubles.
 *
 * Conventions
 *  - Working precision is FP64 (reference: common/module/parkind1.F90:40-44, JPRB).
 *  - Arrays keep the reference's blocked, column-major layout: a field declared
 *    X(NPROMA,KLEV,NBLOCKS) in Fortran is a double* to its first element; element
 *    (jl,jk,ibl) (0-based) lives at  ((ibl*KLEV + jk)*NPROMA + jl).  Half-level fields
 *    (PAPH, PFPLSL, PFPLSN, PFHPSL, PFHPSN) have KLEV+1 levels.
 *    PCLV is (NPROMA,KLEV,NCLV=5,NBLOCKS); the tendency states are the AOSOA buffers
 *    B_CML / B_LOC (NPROMA,KLEV,3+NCLV=8,NBLOCKS) with slabs T=0, A=1, Q=2, CLD=3..7
 *    (common/module/cloudsc2_array_state_mod.F90:129-151).
 *  - NBLOCKS = ceil(NGPTOT/NPROMA); columns ICEND+1..NPROMA of the last block are never
 *    computed (cloudsc2_nl/cloudsc_driver_mod.F90:82-84).
 *  - Every function returns 0 on success, non-zero on error; the text of the last error
 *    is returned by cloudsc2_gpu_last_error().  There is NO CPU fallback: without a
 *    CUDA device every compute entry point fails.  (Reference error convention:
 *    ABOR1 -> abort(), common/module/abor1.F90:10-14; the Fortran shim maps non-zero to ABOR1.)
 *  - "_dev" variants take device pointers (same layout) and enqueue on the library's
 *    stream (or a caller stream given as void* cudaStream_t); the others take host
 *    pointers and perform the H2D / D2H copies inside the call.
 */
#ifndef CLOUDSC2_B200_H
#define CLOUDSC2_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CLOUDSC2_NCLV 5      /* yoecldp.F90:86-91 : NCLV, NCLDQL=1, NCLDQI=2 (1-based) */
#define CLOUDSC2_NSTATE 8    /* 3 + NCLV slabs of a STATE_TYPE buffer                  */

/* Constants that reach the kernels.  Replaces the module variables of
 * common/module/yomcst.F90:167-177, yoethf.F90:79-99, yoecldp.F90:242-370,
 * yoephli.F90:79-97, yomncl.F90:26, yophnc.F90 -- read from input.h5 by the Fortran host
 * and passed by value.  All physics constants are run-time data, never compiled in. */
typedef struct cloudsc2_params {
  /* YOMCST */
  double rg, rd, rcpd, retv, rlvtt, rlstt, rlmlt, rtt;
  /* YOETHF */
  double r2es, r3les, r3ies, r4les, r4ies, r5les, r5ies, r5alvcp, r5alscp;
  double ralvdcp, ralsdcp, rtwat, rtice, rtwat_rtice_r, rvtmp2;
  /* YRECLDP */
  double rclcrit, rkconv, rlmin, rpecons;
  /* YREPHLI */
  double rlptrc;
  /* switches: YREPHLI%LPHYLIN, YRPHNC%LEVAPLS2, YRNCL%LREGCL, driver LDRAIN1D */
  int lphylin, levapls2, lregcl, ldrain1d;
} cloudsc2_params;

/* Blocked host/device arrays of one problem; mirrors the argument list of
 * CLOUDSC_DRIVER (cloudsc2_nl/cloudsc_driver_mod.F90:22-30), same names. */
typedef struct cloudsc2_fields {
  /* inputs */
  const double *pt, *pq, *pap, *paph, *plu, *plude, *pmfu, *pmfd, *psupsat;
  const double *pclv;   /* (NPROMA,KLEV,5,NBLOCKS), species 0=QL 1=QI used           */
  const double *b_cml;  /* TENDENCY_CML buffer (NPROMA,KLEV,8,NBLOCKS)               */
  /* outputs */
  double *b_loc;        /* TENDENCY_LOC buffer (NPROMA,KLEV,8,NBLOCKS): slabs 0,2,3,4 written, 7 zeroed */
  double *pa;           /* cloud fraction, passed as PCLC (driver_mod.F90:105)       */
  double *pcovptot, *pfplsl, *pfplsn, *pfhpsl, *pfhpsn;
} cloudsc2_fields;

/* The 16 increment (TL input / AD output) and 10 increment (TL output / AD input) arrays of
 * CLOUDSC2TL / CLOUDSC2AD (cloudsc2_tl/cloudsc2tl.F90:10-24, cloudsc2_ad/cloudsc2ad.F90:10-24).
 * All plain (NPROMA,KLEV[+1],NBLOCKS) arrays. */
typedef struct cloudsc2_incr_in {   /* perturbations of the 16 inputs  */
  double *paph, *pap, *pq, *pqs, *pt, *pl, *pi, *plude, *plu, *pmfu, *pmfd;
  double *gtent, *gtenq, *gtenl, *gteni, *psupsat;
} cloudsc2_incr_in;
typedef struct cloudsc2_incr_out {  /* perturbations of the 10 outputs */
  double *tent, *tenq, *tenl, *teni, *pclc, *pfplsl, *pfplsn, *pfhpsl, *pfhpsn, *pcovptot;
} cloudsc2_incr_out;

/* ---- life cycle ------------------------------------------------------------------ */

/* Select CUDA device `device`, upload the constants and the CETA(KLEV) vector
 * (cloudsc2_nl/dwarf_cloudsc.F90:100-102; YRECLD%CETA) and create the streams.
 * Replaces the implicit module state used by CLOUDSC2/TL/AD (cloudsc2.F90:104-111,222-224).
 * A single-device set: what one MPI rank of the reference's host uses (one rank per GPU). */
int cloudsc2_gpu_init(const cloudsc2_params *params, int klev, const double *ceta, int device);
/* The same for a set of `ngpus` devices driven by THIS process (devices 0..ngpus-1; ngpus <= 0: all
 * visible; cloudsc2_gpu_init_devices: an explicit list).  Creates one context, one stream set and one
 * host worker thread per device and an NCCL communicator over the set (ncclCommInitAll).  From then
 * on every host-pointer entry point (cloudsc2_gpu_nl / _tl / _ad / _tl_taylor / _ad_test) and the
 * cloudsc2_gpu_state_* entries shard the NPROMA blocks over the devices: device r of R owns blocks
 * [r*per, min(NB,(r+1)*per)), per = (NB-1)/R+1 -- the arithmetic the reference applies to MPI ranks
 * (cloudsc2_nl/dwarf_cloudsc.F90:65-69) -- and the test norms / validation statistics are all-reduced
 * on the devices over NVLink (replaces reduction(max:znormg), cloudsc_driver_tl_mod.F90:125,
 * cloudsc_driver_ad_mod.F90:107, and CLOUDSC_MPI_REDUCE_*, validate_mod.F90:197-199).  A single
 * Fortran process calling CLOUDSC_DRIVER once -- what the reference does -- thus uses all GPUs. */
int cloudsc2_gpu_init_multi(const cloudsc2_params *params, int klev, const double *ceta, int ngpus);
int cloudsc2_gpu_init_devices(const cloudsc2_params *params, int klev, const double *ceta, int ngpus,
                              const int *devices);
int cloudsc2_gpu_finalize(void);
/* The block shard of device / rank `index` of `nshards` (pure function, no GPU needed): per = (NB-1)/R + 1,
 * shard r owns blocks [r*per, min(NB,(r+1)*per)) -- the arithmetic of cloudsc2_nl/dwarf_cloudsc.F90:65-69
 * (NGPTOT = (NGPTOTG-1)/NUMPROC+1, the last rank takes the rest) applied to NPROMA blocks, so that a shard is
 * one contiguous byte range of every blocked array.  Outputs (any may be NULL): first block, number of
 * blocks, valid columns, first global column.  Returns 0, or 3 for bad arguments. */
int cloudsc2_shard_blocks(int index, int nshards, int nproma, int ngptot, int *block0, int *nblocks,
                          int *ngptot_local, long long *gcol0);
/* Number of devices in the set (0 before init). */
int cloudsc2_gpu_num_devices(void);
/* Make device `index` of the set the calling thread's current context for the _dev entry points and
 * the memory helpers (thread-local; index < 0: back to the default = device 0 / the whole set). */
int cloudsc2_gpu_select_device(int index);
/* Multi-PROCESS jobs (one process per GPU, e.g. MPI ranks or torchrun): attach the single-device
 * context to a job-wide NCCL communicator so that the test norms are all-reduced across processes
 * inside the library.  Rank 0 obtains the 128-byte id, the host broadcasts it (MPI_BCAST / any
 * channel), every rank calls comm_init_rank. */
int cloudsc2_gpu_comm_unique_id(void *id128, int bytes);
int cloudsc2_gpu_comm_init_rank(int rank, int nranks, const void *id128, int bytes);
/* rank / size of the current context's communicator (0 / 1 without one) and the NCCL version in use. */
int cloudsc2_gpu_comm_info(int *rank, int *size, int *nccl_version);
/* All-reduce n device-resident doubles in place over the communicator: op 0 = MAX, 1 = MIN, 2 = SUM
 * (CLOUDSC_MPI_REDUCE_MAX / MIN / SUM, cloudsc_mpi_mod.F90).  Synchronous; no-op without a communicator. */
int cloudsc2_gpu_allreduce_dev(double *dev, int n, int op);
/* Text of the calling thread's last error. */
const char *cloudsc2_gpu_last_error(void);
/* 1 if a usable CUDA device is visible, else 0 (never falls back to CPU). */
int cloudsc2_gpu_available(void);
/* Number of visible CUDA devices (0 without a GPU): what a multi-rank host uses to map rank -> device,
 * like the reference maps MPI ranks to cores (cloudsc_mpi_mod.F90). */
int cloudsc2_gpu_device_count(void);
/* number of kernels (and NCCL collectives) launched by this library since init, all devices
 * (bench.py: gpu_launches). */
long long cloudsc2_gpu_launch_count(void);

/* ---- nonlinear: replaces the block loop of CLOUDSC_DRIVER -------------------------- */

/* Host-pointer entry: cloudsc2_nl/cloudsc_driver_mod.F90:82-111 (zero PCOVPTOT and
 * TENDENCY_LOC%CLD(:,:,NCLV), SATUR, CLOUDSC2 for every block).  elapsed_kernel_s /
 * elapsed_total_s (may be NULL) return CUDA-event times of the kernel alone and of
 * copies+kernel. */
int cloudsc2_gpu_nl(int nproma, int klev, int ngptot, double ptsphy,
                    const cloudsc2_fields *host, double *elapsed_kernel_s, double *elapsed_total_s);
/* Device-pointer entry, asynchronous on `stream` (NULL = library stream).  If pqs is non-NULL
 * it is used as PQS(NPROMA,KLEV,NBLOCKS) instead of the fused SATUR (CLOUDSC2 call-only
 * semantics, cloudsc2.F90:10-18). */
int cloudsc2_gpu_nl_dev(int nproma, int klev, int ngptot, double ptsphy,
                        const cloudsc2_fields *dev, const double *pqs, void *stream);

/* SATUR alone (satur.F90:10, LDPHYLIN branch :106-123, called with KFLAG=2 at
 * cloudsc_driver_mod.F90:91-92): pqsat[i] = qsat(pt[i], pap[i]) for n points of any layout.
 * HOST pointers; for users of the CLOUDSC2-call semantics who need PQS explicitly. */
int cloudsc2_gpu_satur(long long n, const double *pap, const double *pt, double *pqsat);

/* ---- tangent linear / adjoint on full fields ---------------------------------------- */

/* CLOUDSC2TL (cloudsc2_tl/cloudsc2tl.F90:10-24) preceded by SATUR for the trajectory PQS
 * (cloudsc_driver_tl_mod.F90:135-136).  Trajectory outputs are written to dev->b_loc, pa,
 * fluxes as the reference does (:1079-1111); increments in `din` -> `dout`. */
int cloudsc2_gpu_tl_dev(int nproma, int klev, int ngptot, double ptsphy,
                        const cloudsc2_fields *dev, const cloudsc2_incr_in *din,
                        const cloudsc2_incr_out *dout, void *stream);
/* CLOUDSC2AD (cloudsc2_ad/cloudsc2ad.F90:10-24): input adjoints in `din` are ACCUMULATED
 * (psupsat assigned, :1733), output adjoints in `dout` are consumed and zeroed. */
int cloudsc2_gpu_ad_dev(int nproma, int klev, int ngptot, double ptsphy,
                        const cloudsc2_fields *dev, const cloudsc2_incr_in *din,
                        const cloudsc2_incr_out *dout, void *stream);
/* Host-pointer convenience wrappers (copies inside the call). */
int cloudsc2_gpu_tl(int nproma, int klev, int ngptot, double ptsphy,
                    const cloudsc2_fields *host, const cloudsc2_incr_in *din,
                    const cloudsc2_incr_out *dout);
int cloudsc2_gpu_ad(int nproma, int klev, int ngptot, double ptsphy,
                    const cloudsc2_fields *host, const cloudsc2_incr_in *din,
                    const cloudsc2_incr_out *dout);

/* ---- self tests: replace the block loops of CLOUDSC_DRIVER_TL / CLOUDSC_DRIVER_AD ---- */

/* Taylor test (cloudsc2_tl/cloudsc_driver_tl_mod.F90:126-254): per block 1 NL + 1 TL + 10
 * perturbed NL, ERROR_NORM (:21-31) over 10 fields, max over blocks.  znormg[10] receives the
 * ratios BEFORE the |1-r| redefinition (:278).  ratios_blk (may be NULL) receives the
 * per-block values ZNORM/ZCOUNT, layout [nblocks][10].  Returns 6 if a block is degenerate
 * (ZNORM==0 or ZCOUNT==0; reference STOPs, :247-249) -- distinct from 3 = bad arguments.
 * With a communicator (device set or comm_init_rank) znormg is the MAX over all ranks. Host pointers. */
int cloudsc2_gpu_tl_taylor(int nproma, int klev, int ngptot, double ptsphy,
                           const cloudsc2_fields *host, double znormg[10], double *ratios_blk);
int cloudsc2_gpu_tl_taylor_dev(int nproma, int klev, int ngptot, double ptsphy,
                               const cloudsc2_fields *dev, double znormg[10], double *ratios_blk);
/* Adjoint (dot-product) test (cloudsc2_ad/cloudsc_driver_ad_mod.F90:108-271): dx = 0.01 x
 * (ZSUPSAT=0), y = TL dx, N1 = <y,y>, dx* = AD y, N2 = <dx, dx*>,
 * N3 = |N1-N2|/eps/N2; *znormg = max N3.  norms_col (may be NULL): [ngptot][3] = N1,N2,N3. */
int cloudsc2_gpu_ad_test(int nproma, int klev, int ngptot, double ptsphy,
                         const cloudsc2_fields *host, double *znormg, double *norms_col);
int cloudsc2_gpu_ad_test_dev(int nproma, int klev, int ngptot, double ptsphy,
                             const cloudsc2_fields *dev, double *znormg, double *norms_col);

/* Verdict logic of the reference drivers, host-side, pure functions.
 * cloudsc_driver_tl_mod.F90:273-311: returns the penalty (>=0 passed iff <=5), or
 * -13 for "err 13" (ISTART==0 or >4); -(ITEST) never used otherwise. */
int cloudsc2_taylor_verdict(const double znormg[10], int *istart_out);
/* cloudsc_driver_ad_mod.F90:286-294: 1 = TEST OK (znormg < 10000). */
int cloudsc2_adjoint_verdict(double znormg);

/* ---- device-side expansion (next row 8f-1) ---------------------------------------------- */

/* Replicate `nlon` source columns cyclically into a blocked array, zero-padding the tail
 * (common/module/expand_mod.F90:270-302): column g (0-based) <- source column g mod nlon.
 * src is (nlon, nlev, ndim) column-major = the input.h5 layout; dst is
 * (nproma, nlev, ndim, nblocks).  Both device pointers. */
int cloudsc2_gpu_expand_dev(const double *src, int nlon, int nlev, int ndim,
                            double *dst, int nproma, int ngptot, void *stream);
/* Same for one shard of a block-sharded problem (cloudsc2_nl/dwarf_cloudsc.F90:65-69: each rank
 * owns a contiguous range of the NGPTOTG global columns): local column j <- source column
 * (gcol0 + j) mod nlon, for the ngptot local columns of this shard. */
int cloudsc2_gpu_expand_shard_dev(const double *src, int nlon, int nlev, int ndim, double *dst,
                                  int nproma, int ngptot, long long gcol0, void *stream);

/* ---- device-resident sharded state (SURVEY 8b "Ownership", 8f-1) --------------------------- */

struct cloudsc2_source;      /* include/cloudsc2_host.h: un-expanded columns in the input.h5 layout */
struct cloudsc2_reference;   /* include/cloudsc2_host.h: un-expanded reference columns               */
#define CLOUDSC2_NVALIDATED 10

/* GLOBAL_STATE%LOAD on the devices (cloudsc2_array_state_mod.F90:153-203): upload the KLON
 * un-expanded source columns once per device (about 4 MB), expand every device's block shard of the
 * NGPTOT columns there (expand_mod.F90:270-335, global column g <- source column g mod KLON) and zero
 * the outputs.  Nothing else crosses PCIe.  The state stays resident until cloudsc2_gpu_state_free /
 * the next load / finalize.  In a process-per-GPU job with a job-wide communicator
 * (cloudsc2_gpu_comm_init_rank) NGPTOT is the global NGPTOTG and every process loads the block shard of
 * its rank (dwarf_cloudsc.F90:65-69 on blocks); _state_validate / the test norms are then global, while
 * _state_get fills only this process' blocks of the full-size host array. */
int cloudsc2_gpu_state_load(const struct cloudsc2_source *src, int nproma, int ngptot);
int cloudsc2_gpu_state_free(void);
/* Shard of device `index`: CUDA ordinal, number of blocks, valid columns, first global column. */
int cloudsc2_gpu_state_info(int index, int *device, int *nblocks, int *ngptot, long long *gcol0);
/* Device pointers of the calling thread's current device's shard (for the _dev entry points). */
int cloudsc2_gpu_state_fields(cloudsc2_fields *out);
/* CLOUDSC_DRIVER / CLOUDSC_DRIVER_TL / CLOUDSC_DRIVER_AD on the resident state, all devices
 * concurrently.  elapsed_s: the slowest device (NL: CUDA events around its kernel; tests: wall time of
 * its launches + norm all-reduce); per_device_s: [num_devices] or NULL.  znormg is all-reduced (MAX)
 * on the devices.  _tl_taylor returns 6 if a block is degenerate (reference STOPs, :247-249). */
int cloudsc2_gpu_state_nl(double *elapsed_s, double *per_device_s);
int cloudsc2_gpu_state_tl_taylor(double znormg[10], double *elapsed_s, double *per_device_s);
int cloudsc2_gpu_state_ad_test(double *znormg, double *elapsed_s, double *per_device_s);
/* GLOBAL_STATE%VALIDATE (cloudsc2_array_state_mod.F90:205-252) on the devices: stats[10][5] =
 * min, max, max|err|, sum|err|, sum|ref| of PLUDE, PCOVPTOT, PFPLSL, PFPLSN, PFHPSL, PFHPSN,
 * TENDENCY_LOC%A, %Q, %T, %CLD against the un-expanded reference columns, reduced over the devices
 * with MIN / MAX / SUM all-reduces (validate_mod.F90:197-199). */
int cloudsc2_gpu_state_validate(const struct cloudsc2_reference *ref, double *stats);
/* Copy one array of the state ("pt" ... "pfhpsn", the member names of cloudsc2_fields), all shards in
 * block order, into a host array of the full blocked size (what the Fortran host would own). */
int cloudsc2_gpu_state_get(const char *name, double *host);
/* The NL program's work flow in ONE call (cloudsc2_nl/dwarf_cloudsc.F90:84-122): LOAD with device-side
 * expansion, CLOUDSC_DRIVER, VALIDATE (skipped when ref or stats is NULL).  elapsed_kernel_s: slowest
 * device's kernel; elapsed_total_s: wall time of the call. */
int cloudsc2_gpu_nl_source(const struct cloudsc2_source *src, const struct cloudsc2_reference *ref,
                           int nproma, int ngptot, double *stats, double *elapsed_kernel_s,
                           double *elapsed_total_s);

/* ---- device-side validation (next row 8f-2) ------------------------------------------------ */

/* Error statistics of one computed field against reference columns, "in the L1 norm sense":
 * replaces VALIDATE_R2 / VALIDATE_R3 (common/module/validate_mod.F90:165-261).
 *   ref_src : DEVICE pointer, un-expanded reference columns (nlon, nlev, ndim) column-major (the
 *             layout of reference.h5); global column g is compared with source column
 *             (gcol0 + g) mod nlon, i.e. with what EXPAND of the reference would have produced
 *   field   : DEVICE pointer, blocked (nproma, nlev, ndim, nblocks)
 *   out[5]  : HOST: min(field), max(field) over whole blocks; max|err|, sum|err|, sum|ref| over
 *             the valid columns (same five numbers ERROR_PRINT receives, :263-296).
 * For a block-sharded run reduce out[] over ranks with min, max, max, sum, sum (the reference's
 * CLOUDSC_MPI_REDUCE_MIN/MAX/SUM, :197-199). Synchronous. */
int cloudsc2_gpu_validate_dev(const double *ref_src, int nlon, const double *field, int nproma,
                              int nlev, int ndim, int ngptot, long long gcol0, double out[5]);
/* Same for a slab range of an AOSOA buffer: `field` points at the first wanted slab of block 0 and
 * consecutive blocks are blk_stride doubles apart -- TENDENCY_LOC%T/%A/%Q/%CLD are
 * B_LOC(:,:,1,:), (:,:,2,:), (:,:,3,:), (:,:,4:,:) with blk_stride = 8*NPROMA*KLEV
 * (cloudsc2_array_state_mod.F90:248-251). */
int cloudsc2_gpu_validate_slabs_dev(const double *ref_src, int nlon, const double *field, int nproma,
                                    int nlev, int ndim, long long blk_stride, int ngptot,
                                    long long gcol0, double out[5]);

/* ---- device memory helpers for non-torch hosts ------------------------------------------ */
int cloudsc2_gpu_malloc(void **ptr, unsigned long long bytes);
int cloudsc2_gpu_free(void *ptr);
int cloudsc2_gpu_memcpy_h2d(void *dst, const void *src, unsigned long long bytes);
int cloudsc2_gpu_memcpy_d2h(void *dst, const void *src, unsigned long long bytes);
int cloudsc2_gpu_memset(void *dst, int value, unsigned long long bytes);
int cloudsc2_gpu_sync(void);
/* Page-lock / unlock caller-owned host arrays so that the copies inside the host-pointer entry
 * points run at full PCIe rate and overlap with the kernel (the Fortran host keeps ownership:
 * ALLOCATE in expand_mod.F90:110,127,148).  Optional; pageable memory works, slower. */
int cloudsc2_gpu_host_register(void *ptr, unsigned long long bytes);
int cloudsc2_gpu_host_unregister(void *ptr);
/* Allocate / free page-locked host memory directly (cudaHostAlloc): what a host uses INSTEAD of
 * ALLOCATE + cloudsc2_gpu_host_register when it can (Fortran: C_F_POINTER on the result).  DMA from
 * such memory runs at the full PCIe rate; registered pageable memory was measured ~15 % slower. */
int cloudsc2_gpu_host_alloc(void **ptr, unsigned long long bytes);
int cloudsc2_gpu_host_free(void *ptr);

/* ---- tuning ------------------------------------------------------------------------------- */
/* Integer tuning options (defaults come from the environment variables in parentheses):
 *   "e2e_mode"     (CSC2_E2E_MODE)     0/1 = staged chunked copies in the host-pointer entry points,
 *                                      2 = zero-copy: the kernel reads/writes page-locked, mapped
 *                                      host arrays directly (cloudsc2_gpu_host_register)
 *   "e2e_chunk_mb" (CSC2_E2E_CHUNK_MB) cap of the staging chunk size in MB (default 256)
 *   "e2e_host_derive" (CSC2_E2E_HOST_DERIVE) 1 (default): the host-pointer NL entry does not copy
 *                                      PCOVPTOT, TENDENCY_LOC%CLD(:,:,NCLV) (identically zero) and
 *                                      PFHPSL/PFHPSN (= -PFPLSL*RLVTT, -PFPLSN*RLSTT) back over PCIe
 *                                      but fills them on the host, bit-identically; 0: copy all
 *   "lregcl"                           YRNCL%LREGCL (the regularised TL / AD, cloudsc2tl.F90:575-580 ...) of every
 *                                      context of the device set, without a new init: the reference's programs
 *                                      set this module variable before the driver call
 *                                      (cloudsc2_tl/dwarf_cloudsc.F90:105, cloudsc2_ad/dwarf_cloudsc.F90:105)
 *   "ad_have_trajectory"               1: cloudsc2_gpu_ad_dev trusts that dev->pfplsl / dev->pfplsn
 *                                      already hold the trajectory fluxes of the same inputs (a
 *                                      cloudsc2_gpu_nl_dev / _tl_dev call ran before, as in every
 *                                      4D-Var inner loop) and skips its forward sweep; 0 (default):
 *                                      CLOUDSC2AD as written, trajectory recomputed (cloudsc2ad.F90:364-866)
 * Nothing like this exists in the reference (its only knobs are NUMOMP and NPROMA). */
int cloudsc2_gpu_set_option(const char *name, int value);

/* ---- diagnostics -------------------------------------------------------------------------- */
/* Evaluate one of the kernels' branch-free FP64 elementary functions (csrc/cloudsc2_math.cuh,
 * the replacements of the Fortran intrinsics EXP/TANH/COSH/SQRT and of "/" used by
 * cloudsc2.F90) on n host values: fn 0 = 1/x, 1 = exp, 2 = exp clamped below, 3 = sqrt,
 * 4 = tanh+1, 5 = 1/cosh^2.  Used by the accuracy tests only. */
int cloudsc2_gpu_math_probe(int fn, const double *x, double *y, int n);

#ifdef __cplusplus
}
#endif
#endif /* CLOUDSC2_B200_H */
