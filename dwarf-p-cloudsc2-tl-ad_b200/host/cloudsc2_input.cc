// cloudsc2_input.cc -- CLOUDSC2_ARRAY_STATE%LOAD's reads of input.h5 and VALIDATE's reads of
// reference.h5 (common/module/cloudsc2_array_state_mod.F90:153-203, :225-236) through the mini
// HDF5 reader, without libhdf5.  Interfaces and citations: include/cloudsc2_host.h.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "cloudsc2_host.h"

namespace {

thread_local std::string g_input_error;

double *dalloc(size_t n) { return static_cast<double *>(std::calloc(n ? n : 1, sizeof(double))); }

// LOAD_ARRAY (file_io_mod.F90): a missing or mis-shaped dataset aborts the reference; here it is
// an error code + message.
bool read_array(const char *path, const char *name, double *out, size_t n) {
  long long got = cloudsc2_h5_read_f8(path, name, nullptr, 0, nullptr, nullptr);
  if (got < 0) {
    g_input_error = std::string("dataset '") + name + "' in " + path + ": error " + std::to_string(got);
    return false;
  }
  if ((size_t)got != n) {
    g_input_error = std::string("dataset '") + name + "' in " + path + " has " + std::to_string(got) +
                    " elements, expected " + std::to_string(n);
    return false;
  }
  return cloudsc2_h5_read_f8(path, name, out, (long long)n, nullptr, nullptr) == got;
}

bool read_scalar(const char *path, const char *name, double *out) { return read_array(path, name, out, 1); }

bool read_int(const char *path, const char *name, int *out) {
  long long got = cloudsc2_h5_read_i4(path, name, out, 1);
  if (got != 1) {
    g_input_error = std::string("scalar '") + name + "' in " + path + ": error " + std::to_string(got);
    return false;
  }
  return true;
}

}  // namespace

extern "C" {

const char *cloudsc2_input_last_error(void) { return g_input_error.c_str(); }

int cloudsc2_source_load_h5(cloudsc2_source *s, cloudsc2_params *p, const char *path) {
  if (!s || !p || !path) return 1;
  std::memset(s, 0, sizeof(*s));
  int klon = 0, klev = 0;
  if (!read_int(path, "KLON", &klon) || !read_int(path, "KLEV", &klev)) return 2;
  if (klon <= 0 || klev <= 0 || klev > 200) {   // dwarf_cloudsc.F90:88-91 : ZPRES(0:200)
    g_input_error = "KLON/KLEV out of range in " + std::string(path);
    return 2;
  }
  s->klon = klon;
  s->klev = klev;
  const size_t n = (size_t)klon * klev;
  s->pt = dalloc(n); s->pq = dalloc(n); s->pap = dalloc(n); s->paph = dalloc(n + klon);
  s->plu = dalloc(n); s->plude = dalloc(n); s->pmfu = dalloc(n); s->pmfd = dalloc(n);
  s->pa = dalloc(n); s->psupsat = dalloc(n);
  s->pclv = dalloc(n * CLOUDSC2_NCLV);
  s->tend_cml = dalloc(n * CLOUDSC2_NSTATE);
  s->ceta = dalloc(klev);
  bool ok = s->pt && s->pq && s->pap && s->paph && s->plu && s->plude && s->pmfu && s->pmfd &&
            s->pa && s->psupsat && s->pclv && s->tend_cml && s->ceta;
  if (!ok) g_input_error = "out of memory";
  // cloudsc2_array_state_mod.F90:167-183 (fields) and expand_mod.F90:151-154 (STATE_TYPE members)
  ok = ok && read_array(path, "PT", s->pt, n) && read_array(path, "PQ", s->pq, n) &&
       read_array(path, "PAP", s->pap, n) && read_array(path, "PAPH", s->paph, n + klon) &&
       read_array(path, "PLU", s->plu, n) && read_array(path, "PLUDE", s->plude, n) &&
       read_array(path, "PMFU", s->pmfu, n) && read_array(path, "PMFD", s->pmfd, n) &&
       read_array(path, "PA", s->pa, n) && read_array(path, "PSUPSAT", s->psupsat, n) &&
       read_array(path, "PCLV", s->pclv, n * CLOUDSC2_NCLV) &&
       read_array(path, "TENDENCY_CML_T", s->tend_cml, n) &&
       read_array(path, "TENDENCY_CML_A", s->tend_cml + n, n) &&
       read_array(path, "TENDENCY_CML_Q", s->tend_cml + 2 * n, n) &&
       read_array(path, "TENDENCY_CML_CLD", s->tend_cml + 3 * n, n * CLOUDSC2_NCLV);
  // :193 PTSPHY; yomcst.F90:168-176; yoethf.F90:80-98; yoecldp.F90:247-259; yoephli.F90:86
  cloudsc2_default_params(p);   // switches as the NL program sets them; every number below replaced
  ok = ok && read_scalar(path, "PTSPHY", &s->ptsphy) &&
       read_scalar(path, "RG", &p->rg) && read_scalar(path, "RD", &p->rd) &&
       read_scalar(path, "RCPD", &p->rcpd) && read_scalar(path, "RETV", &p->retv) &&
       read_scalar(path, "RLVTT", &p->rlvtt) && read_scalar(path, "RLSTT", &p->rlstt) &&
       read_scalar(path, "RLMLT", &p->rlmlt) && read_scalar(path, "RTT", &p->rtt) &&
       read_scalar(path, "R2ES", &p->r2es) && read_scalar(path, "R3LES", &p->r3les) &&
       read_scalar(path, "R3IES", &p->r3ies) && read_scalar(path, "R4LES", &p->r4les) &&
       read_scalar(path, "R4IES", &p->r4ies) && read_scalar(path, "R5LES", &p->r5les) &&
       read_scalar(path, "R5IES", &p->r5ies) && read_scalar(path, "R5ALVCP", &p->r5alvcp) &&
       read_scalar(path, "R5ALSCP", &p->r5alscp) && read_scalar(path, "RALVDCP", &p->ralvdcp) &&
       read_scalar(path, "RALSDCP", &p->ralsdcp) && read_scalar(path, "RTWAT", &p->rtwat) &&
       read_scalar(path, "RTICE", &p->rtice) && read_scalar(path, "RTWAT_RTICE_R", &p->rtwat_rtice_r) &&
       read_scalar(path, "YRECLDP_RCLCRIT", &p->rclcrit) && read_scalar(path, "YRECLDP_RKCONV", &p->rkconv) &&
       read_scalar(path, "YRECLDP_RLMIN", &p->rlmin) && read_scalar(path, "YRECLDP_RPECONS", &p->rpecons) &&
       read_scalar(path, "YREPHLI_RLPTRC", &p->rlptrc);
  if (!ok) {
    cloudsc2_source_free(s);
    return 3;
  }
  p->rvtmp2 = 0.0;   // declared yoethf.F90:30, absent from YOETHF_LOAD_PARAMETERS :79-99
  // dwarf_cloudsc.F90:100-102 : CETA(JK) = PAP(1,JK,1)/PAPH(1,KLEV+1,1)
  for (int k = 0; k < klev; ++k) s->ceta[k] = s->pap[(size_t)k * klon] / s->paph[(size_t)klev * klon];
  return 0;
}

int cloudsc2_reference_load_h5(cloudsc2_reference *r, const char *path) {
  if (!r || !path) return 1;
  std::memset(r, 0, sizeof(*r));
  int klon = 0, klev = 0;
  if (!read_int(path, "KLON", &klon) || !read_int(path, "KLEV", &klev)) return 2;
  if (klon <= 0 || klev <= 0) return 2;
  r->klon = klon;
  r->klev = klev;
  const size_t n = (size_t)klon * klev, nh = n + klon;
  r->plude = dalloc(n); r->pcovptot = dalloc(n);
  r->pfplsl = dalloc(nh); r->pfplsn = dalloc(nh); r->pfhpsl = dalloc(nh); r->pfhpsn = dalloc(nh);
  r->tend_loc = dalloc(n * CLOUDSC2_NSTATE);
  bool ok = r->plude && r->pcovptot && r->pfplsl && r->pfplsn && r->pfhpsl && r->pfhpsn && r->tend_loc;
  if (!ok) g_input_error = "out of memory";
  // cloudsc2_array_state_mod.F90:225-233
  ok = ok && read_array(path, "PLUDE", r->plude, n) && read_array(path, "PCOVPTOT", r->pcovptot, n) &&
       read_array(path, "PFPLSL", r->pfplsl, nh) && read_array(path, "PFPLSN", r->pfplsn, nh) &&
       read_array(path, "PFHPSL", r->pfhpsl, nh) && read_array(path, "PFHPSN", r->pfhpsn, nh) &&
       read_array(path, "TENDENCY_LOC_T", r->tend_loc, n) &&
       read_array(path, "TENDENCY_LOC_A", r->tend_loc + n, n) &&
       read_array(path, "TENDENCY_LOC_Q", r->tend_loc + 2 * n, n) &&
       read_array(path, "TENDENCY_LOC_CLD", r->tend_loc + 3 * n, n * CLOUDSC2_NCLV);
  if (!ok) {
    cloudsc2_reference_free(r);
    return 3;
  }
  return 0;
}

void cloudsc2_reference_free(cloudsc2_reference *r) {
  if (!r) return;
  std::free(r->plude); std::free(r->pcovptot); std::free(r->pfplsl); std::free(r->pfplsn);
  std::free(r->pfhpsl); std::free(r->pfhpsn); std::free(r->tend_loc);
  std::memset(r, 0, sizeof(*r));
}

}  // extern "C"
