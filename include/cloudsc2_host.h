/*
 * cloudsc2_host.h -- host-side helpers of the B200 CLOUDSC2 library (C ABI, no CUDA needed).
 *
 * These mirror the pieces of the reference's Fortran host that sit directly either side of the
 * hot path: the synthetic stand-in for config-files/input.h5 (absent from the reference
 * checkout), the 100 -> NGPTOT column expansion, the blocked-array container, and a minimal
 * reader for the contiguous HDF5 files the reference ships.  Reference interfaces are cited
 * as file:line relative to the reference's src/ directory.
 */
#ifndef CLOUDSC2_HOST_H
#define CLOUDSC2_HOST_H
#include "cloudsc2_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* IFS-standard constants consistent with yoethf.F90:57-69 and with config-files/reference.h5
 * (SURVEY Appendix E).  ASSUMPTION: input.h5 is absent, these are not read from the reference.
 * Switches are set as the NL program sets them (cloudsc2_nl/dwarf_cloudsc.F90:105-107). */
void cloudsc2_default_params(cloudsc2_params *p);

/* Un-expanded source columns in the layout of input.h5: every field is
 * (KLON,KLEV[,NDIM]) column-major, i.e. C-order (NDIM,KLEV,KLON), KLON contiguous. */
typedef struct cloudsc2_source {   /* (tag declared in cloudsc2_b200.h) */
  int klon, klev;
  double ptsphy;
  double *pt, *pq, *pap, *paph /*klev+1*/, *plu, *plude, *pmfu, *pmfd, *pa, *psupsat;
  double *pclv;          /* (KLON,KLEV,5)                                                */
  double *tend_cml;      /* (KLON,KLEV,8): slabs T,A,Q,CLD(5) like B_CML of one block    */
  double *ceta;          /* (KLEV) = PAP(1,JK)/PAPH(1,KLEV+1), dwarf_cloudsc.F90:100-102 */
} cloudsc2_source;

/* Allocate and fill `klon` physically plausible, all-active columns (seeded, deterministic).
 * Stand-in for CLOUDSC2_ARRAY_STATE%LOAD's reads of input.h5
 * (common/module/cloudsc2_array_state_mod.F90:153-203). */
int cloudsc2_source_synth(cloudsc2_source *s, unsigned long long seed, int klon, int klev,
                          const cloudsc2_params *p);
void cloudsc2_source_free(cloudsc2_source *s);

/* Host expansion, common/module/expand_mod.F90:270-335 (EXPAND_R2/R3): column g (0-based) of
 * the blocked field <- source column g mod nlon; tail of last block zero.  The reference reads
 * out of bounds when a block starts at a multiple of nlon other than nlon itself
 * (expand_mod.F90:283); that is not replicated. */
void cloudsc2_expand_host(const double *src, int nlon, int nlev, int ndim, double *dst,
                          int nproma, int ngptot);

/* Blocked arrays of one problem, owned by the library (malloc).  Mirrors the allocatable
 * members of CLOUDSC2_ARRAY_STATE (cloudsc2_array_state_mod.F90:28-60). */
typedef struct cloudsc2_state {
  int nproma, klev, ngptot, nblocks;
  cloudsc2_fields f;     /* all pointers owned by this struct */
} cloudsc2_state;
/* LOAD: allocate, expand inputs from `s`, zero outputs (FIELD_INIT :100-127; B_LOC is
 * allocated but not zeroed by the reference :129-151 -- we zero it for determinism). */
int cloudsc2_state_load(cloudsc2_state *st, const cloudsc2_source *s, int nproma, int ngptot);
void cloudsc2_state_free(cloudsc2_state *st);
int cloudsc2_nblocks(int ngptot, int nproma);

/* Minimal HDF5 reader: superblock v0, contiguous little-endian f8/i4 datasets in the root group
 * -- what config-files/reference.h5 (and the missing input.h5) use.  Replaces the calls the
 * host makes into libhdf5 (common/module/hdf5_file_mod.F90:135-164) for these files only.
 * Returns the number of elements of the dataset (at most max_elems of them are copied into out,
 * which may be NULL to query the size), or <0: -1 cannot open, -2 bad format, -3 not found,
 * -4 unsupported feature, -5 wrong element type. */
long long cloudsc2_h5_read_f8(const char *path, const char *dataset, double *out,
                              long long max_elems, int dims_out[4], int *ndims_out);
long long cloudsc2_h5_read_i4(const char *path, const char *dataset, int *out,
                              long long max_elems);

/* Minimal HDF5 writer, the counterpart of the reader: the same on-disk structures (superblock v0, root
 * group, contiguous little-endian datasets); is_int 0 = f8, 1 = i4; dims in C order.  Replaces the
 * host's writes through libhdf5 (hdf5_file_mod.F90 hdf5_file_write_*).  0 = ok. */
typedef struct cloudsc2_h5_dataset {
  const char *name;
  int is_int, rank;
  long long dims[4];
  const void *data;
} cloudsc2_h5_dataset;
int cloudsc2_h5_write(const char *path, const cloudsc2_h5_dataset *ds, int n);

/* CLOUDSC2_ARRAY_STATE%LOAD's reads of input.h5 (cloudsc2_array_state_mod.F90:153-203): KLON, KLEV,
 * the 100-column fields, PTSPHY and the constants of YOMCST / YOETHF / YRECLDP / YREPHLI that reach
 * the kernels, through the mini reader above.  A missing dataset is an error
 * (the reference aborts).  Switches in *p are set as the NL program sets them; RVTMP2 = 0 because
 * the reference never loads it (yoethf.F90:30 vs :79-99).  0 = ok. */
int cloudsc2_source_load_h5(cloudsc2_source *s, cloudsc2_params *p, const char *path);
/* The un-expanded columns of reference.h5 that VALIDATE compares with
 * (cloudsc2_array_state_mod.F90:225-233); tend_loc is (KLON,KLEV,8): slabs T,A,Q,CLD(5). */
typedef struct cloudsc2_reference {   /* (tag declared in cloudsc2_b200.h) */
  int klon, klev;
  double *plude, *pcovptot, *pfplsl, *pfplsn, *pfhpsl, *pfhpsn; /* fluxes: klev+1 */
  double *tend_loc;
} cloudsc2_reference;
int cloudsc2_reference_load_h5(cloudsc2_reference *r, const char *path);
void cloudsc2_reference_free(cloudsc2_reference *r);
/* Write an input.h5 holding `s` and `p`: every dataset CLOUDSC2_ARRAY_STATE%LOAD reads -- the fields and
 * constants that reach the kernels with their values, all the other scalars the reference's loaders
 * insist on (yoecldp.F90:244-369 etc.; unused by CLOUDSC2) as zeros -- so that the reference's own
 * binaries can be run on the synthetic columns on a machine that has them.  And WRITE_REFERENCE
 * (cloudsc2_array_state_mod.F90:260-287): the ten validated fields of the un-expanded columns. */
int cloudsc2_source_write_h5(const cloudsc2_source *s, const cloudsc2_params *p, const char *path);
int cloudsc2_reference_write_h5(const cloudsc2_reference *r, const char *path);
/* Text of the last error of the two loaders (thread-local). */
const char *cloudsc2_input_last_error(void);

/* Validation statistics of one field, common/module/validate_mod.F90:165-211,263-296
 * (L1 sense): out[0]=min(field) out[1]=max(field) out[2]=max|err| out[3]=sum|err|
 * out[4]=sum|ref| ; relative error % as the reference prints it = see cloudsc2_error_rel. */
void cloudsc2_validate_host(const double *ref, const double *field, int nproma, int nlev,
                            int ngptot, double out[5]);
double cloudsc2_error_rel(const double stats[5], int *flag_out);

#ifdef __cplusplus
}
#endif
#endif
