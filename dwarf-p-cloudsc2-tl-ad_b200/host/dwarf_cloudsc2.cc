// dwarf_cloudsc2.cc -- the three dwarf programs on top of the C ABI:
//
//     dwarf-cloudsc2-nl | dwarf-cloudsc2-tl | dwarf-cloudsc2-ad   [NUMOMP [NGPTOT [NPROMA]]]
//
// Host-side mirror of PROGRAM DWARF_CLOUDSC (cloudsc2_nl/dwarf_cloudsc.F90:10-131 and its TL / AD
// twins) for machines without the reference's Fortran toolchain: same command line and defaults
// (:27-29, :49-75), same order of work -- LOAD (input.h5 -> expand to NGPTOT columns in NPROMA
// blocks), CETA from the first column (:100-102), the driver (block loop -> ONE library call), the
// performance table (timer_mod.F90:114-174, on stderr like the reference's unit 0), then VALIDATE
// (NL, cloudsc2_array_state_mod.F90:205-252) or the Taylor / adjoint verdicts
// (cloudsc_driver_tl_mod.F90:272-311, cloudsc_driver_ad_mod.F90:285-294).
//
// All computing happens in libcloudsc2_b200.so (CUDA, sm_100a); there is no CPU fallback: without a
// GPU the program aborts like ABOR1 (abor1.F90:10-14).  The state lives on the device: the
// 100 source columns are uploaded and expanded there; CLOUDSC2_HOST_ARRAYS=1 keeps the
// blocked arrays on the host instead and goes through the host-pointer entry points (copies inside
// the timed call) -- what the unchanged Fortran host would do.
//
// Environment:
//   CLOUDSC2_INPUT      path of input.h5 (default ./input.h5 if it exists; else synthetic columns)
//   CLOUDSC2_REFERENCE  path of reference.h5 for the NL validation (default ./reference.h5 if it
//                       exists; else the un-expanded columns computed as one block are the reference)
//   CLOUDSC2_WRITE_INPUT  path: write the columns and constants in use as an input.h5 (all datasets the
//                       reference's loaders read) before running -- to run the reference's binaries
//                       on the synthetic columns elsewhere
//   CLOUDSC2_WRITE_REFERENCE  1: after the NL run write ./reference.h5 from block 1, like the reference
//                       (dwarf_cloudsc.F90:124-126; needs NPROMA = KLON, cloudsc2_array_state_mod.F90:265-268)
//   CLOUDSC2_SYNTH_SEED / CLOUDSC2_SYNTH_KLON / CLOUDSC2_SYNTH_KLEV   synthetic input (0 / 100 / 137)
//   CLOUDSC2_DEVICE     CUDA device ordinal of rank 0 (0); rank r uses device (CLOUDSC2_DEVICE + r) mod #devices
//   CLOUDSC2_REPEAT     timed repetitions of the driver call, best one reported (1)
//   CLOUDSC2_HOST_ARRAYS  1: see above (pageable arrays, like a Fortran ALLOCATE); 2: page-locked arrays
//                       from cloudsc2_gpu_host_alloc (what INTEGRATION.md recommends to the Fortran host)
//   CLOUDSC2_NUMPROC    number of ranks = GPUs (1).  The reference distributes NGPTOT over MPI ranks
//                       (dwarf_cloudsc.F90:63-67, cloudsc_mpi_mod.F90); here the program forks one process
//                       per GPU, rank r takes the global columns [r*per, (r+1)*per), per =
//                       (NGPTOT-1)/NUMPROC+1, and rank 0 gathers the timings, MAX-reduces the test norms
//                       and MIN/MAX/SUM-reduces the validation statistics -- the same reductions the
//                       reference does with MPI, over pipes because they are a few dozen doubles.  (A rank
//                       expands the source columns from its GLOBAL column offset, so results do not
//                       depend on NUMPROC; the reference restarts every rank at source column 1.)
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include "cloudsc2_host.h"

namespace {

enum Mode { NL, TL, AD };

[[noreturn]] void abor1(const std::string &msg) {   // abor1.F90:10-14
  std::fprintf(stderr, " ABOR1: %s\n", msg.c_str());
  std::fflush(nullptr);
  std::exit(1);
}

void ck(int rc, const char *what) {
  if (rc) abor1(std::string(what) + ": " + cloudsc2_gpu_last_error());
}

bool file_exists(const char *p) {
  if (FILE *f = std::fopen(p, "rb")) { std::fclose(f); return true; }
  return false;
}

int env_int(const char *name, int dflt) {
  const char *v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}

// READ(CLARG,*) of an integer (dwarf_cloudsc.F90:52-55): garbage aborts.
int parse_int_arg(const char *s, const char *what) {
  char *end = nullptr;
  long v = std::strtol(s, &end, 10);
  if (end == s || *end != '\0' || v <= 0 || v > std::numeric_limits<int>::max())
    abor1(std::string("bad value for ") + what + ": '" + s + "'");
  return (int)v;
}

double now_s() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// Fortran E20.13 (validate_mod.F90:293).
std::string fortran_e(double x) {
  char buf[64];
  if (std::isnan(x)) return std::string(17, ' ') + "NaN";
  if (std::isinf(x)) { std::snprintf(buf, sizeof buf, "%20s", x > 0 ? "Infinity" : "-Infinity"); return buf; }
  if (x == 0.0) { std::snprintf(buf, sizeof buf, "%20s", std::signbit(x) ? "-0.0000000000000E+00" : "0.0000000000000E+00"); return buf; }
  char m[40];
  std::snprintf(m, sizeof m, "%.12E", std::fabs(x));          // d.ddddddddddddE+ee
  std::string s(m);
  const size_t e = s.find('E');
  const int exp10 = std::atoi(s.c_str() + e + 1) + 1;
  std::string digits = s.substr(0, 1) + s.substr(2, e - 2);
  char out[64];
  std::snprintf(out, sizeof out, "%s0.%sE%c%02d", x < 0 ? "-" : "", digits.c_str(), exp10 >= 0 ? '+' : '-', std::abs(exp10));
  std::snprintf(buf, sizeof buf, "%20s", out);
  return buf;
}

// ERROR_PRINT, validate_mod.F90:263-296.
int error_print(const char *name, int ndim, const double st[5], long long ngptotg) {
  const double zeps = std::numeric_limits<double>::epsilon();
  double rel; int iopt;
  if (st[3] < zeps) { rel = 0.0; iopt = 1; }
  else if (st[4] < zeps) { rel = st[3] / (1.0 + st[4]); iopt = 2; }
  else { rel = st[3] / st[4]; iopt = 3; }
  const bool warn = rel > 10.0 * zeps;
  std::printf(" %-20s %1dD%1d %s %s %s %s %s%s\n", name, ndim, iopt, fortran_e(st[0]).c_str(),
              fortran_e(st[1]).c_str(), fortran_e(st[2]).c_str(),
              fortran_e(st[3] / (double)ngptotg).c_str(), fortran_e(100.0 * rel).c_str(), warn ? " !!!!" : "");
  return warn ? 1 : 0;
}

struct DevArray {
  double *p = nullptr;
  size_t n = 0;
  void alloc(size_t count) {
    n = count;
    ck(cloudsc2_gpu_malloc(reinterpret_cast<void **>(&p), count * sizeof(double)), "cloudsc2_gpu_malloc");
    ck(cloudsc2_gpu_memset(p, 0, count * sizeof(double)), "cloudsc2_gpu_memset");
  }
  void release() { if (p) cloudsc2_gpu_free(p); p = nullptr; }
};

// Upload `count` doubles and expand them on the device (expand_mod.F90:270-335 -> k_expand).
void load_and_expand(const double *src, int klon, int nlev, int ndim, DevArray &dst, int nproma, int ngptot,
                     long long gcol0) {
  DevArray tmp;
  const size_t count = (size_t)klon * nlev * ndim;
  ck(cloudsc2_gpu_malloc(reinterpret_cast<void **>(&tmp.p), count * sizeof(double)), "cloudsc2_gpu_malloc");
  ck(cloudsc2_gpu_memcpy_h2d(tmp.p, src, count * sizeof(double)), "cloudsc2_gpu_memcpy_h2d");
  ck(cloudsc2_gpu_expand_shard_dev(tmp.p, klon, nlev, ndim, dst.p, nproma, ngptot, gcol0, nullptr),
     "cloudsc2_gpu_expand_shard_dev");
  ck(cloudsc2_gpu_sync(), "cloudsc2_gpu_sync");
  tmp.release();
}

struct DeviceState {   // CLOUDSC2_ARRAY_STATE on the device (cloudsc2_array_state_mod.F90:28-60)
  DevArray pt, pq, pap, paph, plu, plude, pmfu, pmfd, psupsat, pclv, b_cml, b_loc, pa, pcovptot,
      pfplsl, pfplsn, pfhpsl, pfhpsn;
  cloudsc2_fields f{};
  void load(const cloudsc2_source &s, int nproma, int ngptot, long long gcol0) {
    const int nb = cloudsc2_nblocks(ngptot, nproma);
    const size_t n = (size_t)nproma * s.klev * nb, nh = (size_t)nproma * (s.klev + 1) * nb;
    pt.alloc(n); pq.alloc(n); pap.alloc(n); paph.alloc(nh); plu.alloc(n); plude.alloc(n);
    pmfu.alloc(n); pmfd.alloc(n); psupsat.alloc(n); pclv.alloc(n * CLOUDSC2_NCLV);
    b_cml.alloc(n * CLOUDSC2_NSTATE); b_loc.alloc(n * CLOUDSC2_NSTATE); pa.alloc(n);
    pcovptot.alloc(n); pfplsl.alloc(nh); pfplsn.alloc(nh); pfhpsl.alloc(nh); pfhpsn.alloc(nh);
    load_and_expand(s.pt, s.klon, s.klev, 1, pt, nproma, ngptot, gcol0);
    load_and_expand(s.pq, s.klon, s.klev, 1, pq, nproma, ngptot, gcol0);
    load_and_expand(s.pap, s.klon, s.klev, 1, pap, nproma, ngptot, gcol0);
    load_and_expand(s.paph, s.klon, s.klev + 1, 1, paph, nproma, ngptot, gcol0);
    load_and_expand(s.plu, s.klon, s.klev, 1, plu, nproma, ngptot, gcol0);
    load_and_expand(s.plude, s.klon, s.klev, 1, plude, nproma, ngptot, gcol0);
    load_and_expand(s.pmfu, s.klon, s.klev, 1, pmfu, nproma, ngptot, gcol0);
    load_and_expand(s.pmfd, s.klon, s.klev, 1, pmfd, nproma, ngptot, gcol0);
    load_and_expand(s.pa, s.klon, s.klev, 1, pa, nproma, ngptot, gcol0);
    load_and_expand(s.psupsat, s.klon, s.klev, 1, psupsat, nproma, ngptot, gcol0);
    load_and_expand(s.pclv, s.klon, s.klev, CLOUDSC2_NCLV, pclv, nproma, ngptot, gcol0);
    load_and_expand(s.tend_cml, s.klon, s.klev, CLOUDSC2_NSTATE, b_cml, nproma, ngptot, gcol0);
    f.pt = pt.p; f.pq = pq.p; f.pap = pap.p; f.paph = paph.p; f.plu = plu.p; f.plude = plude.p;
    f.pmfu = pmfu.p; f.pmfd = pmfd.p; f.psupsat = psupsat.p; f.pclv = pclv.p; f.b_cml = b_cml.p;
    f.b_loc = b_loc.p; f.pa = pa.p; f.pcovptot = pcovptot.p; f.pfplsl = pfplsl.p; f.pfplsn = pfplsn.p;
    f.pfhpsl = pfhpsl.p; f.pfhpsn = pfhpsn.p;
  }
  void release() {
    for (DevArray *a : {&pt, &pq, &pap, &paph, &plu, &plude, &pmfu, &pmfd, &psupsat, &pclv, &b_cml, &b_loc,
                        &pa, &pcovptot, &pfplsl, &pfplsn, &pfhpsl, &pfhpsn})
      a->release();
  }
};

// One validated field: reference columns (un-expanded, host) against a blocked device field of the
// shard starting at global column gcol0 -> the five numbers ERROR_PRINT needs.
void validate_field(const double *ref_cols, int klon, const double *dev_field, int nproma, int nlev, int ndim,
                    int ngptot, long long gcol0, long long blk_stride, double st[5]) {
  DevArray r;
  const size_t count = (size_t)klon * nlev * ndim;
  ck(cloudsc2_gpu_malloc(reinterpret_cast<void **>(&r.p), count * sizeof(double)), "cloudsc2_gpu_malloc");
  ck(cloudsc2_gpu_memcpy_h2d(r.p, ref_cols, count * sizeof(double)), "cloudsc2_gpu_memcpy_h2d");
  if (blk_stride == 0) blk_stride = (long long)nproma * nlev * ndim;
  ck(cloudsc2_gpu_validate_slabs_dev(r.p, klon, dev_field, nproma, nlev, ndim, blk_stride, ngptot, gcol0, st),
     "cloudsc2_gpu_validate_slabs_dev");
  r.release();
}

// The un-expanded columns run as ONE block of KLON columns: the stand-in for reference.h5.
void self_reference(const cloudsc2_source &s, cloudsc2_reference &r) {
  DeviceState d;
  d.load(s, s.klon, s.klon, 0);
  ck(cloudsc2_gpu_nl_dev(s.klon, s.klev, s.klon, s.ptsphy, &d.f, nullptr, nullptr), "cloudsc2_gpu_nl_dev");
  ck(cloudsc2_gpu_sync(), "cloudsc2_gpu_sync");
  std::memset(&r, 0, sizeof r);
  r.klon = s.klon; r.klev = s.klev;
  const size_t n = (size_t)s.klon * s.klev, nh = n + s.klon;
  auto fetch = [&](const double *dev, size_t count) {
    double *h = static_cast<double *>(std::malloc(count * sizeof(double)));
    if (!h) abor1("out of memory");
    ck(cloudsc2_gpu_memcpy_d2h(h, dev, count * sizeof(double)), "cloudsc2_gpu_memcpy_d2h");
    return h;
  };
  r.plude = fetch(d.plude.p, n); r.pcovptot = fetch(d.pcovptot.p, n);
  r.pfplsl = fetch(d.pfplsl.p, nh); r.pfplsn = fetch(d.pfplsn.p, nh);
  r.pfhpsl = fetch(d.pfhpsl.p, nh); r.pfhpsn = fetch(d.pfhpsn.p, nh);
  r.tend_loc = fetch(d.b_loc.p, n * CLOUDSC2_NSTATE);
  d.release();
}

// Blocked HOST arrays of one shard (what the Fortran host owns, expand_mod.F90:110,127,148): filled from
// the device-expanded state.  pinned: page-locked memory from the library (cloudsc2_gpu_host_alloc).
struct HostState {
  cloudsc2_fields f{};
  std::vector<void *> owned;
  bool pinned = false;
  double *grab(const double *dev, size_t count) {
    void *h = nullptr;
    if (pinned) ck(cloudsc2_gpu_host_alloc(&h, count * sizeof(double)), "cloudsc2_gpu_host_alloc");
    else if (!(h = std::malloc(count * sizeof(double)))) abor1("out of memory");
    owned.push_back(h);
    ck(cloudsc2_gpu_memcpy_d2h(h, dev, count * sizeof(double)), "cloudsc2_gpu_memcpy_d2h");
    return static_cast<double *>(h);
  }
  void from_device(const DeviceState &d, bool pin) {
    pinned = pin;
    f.pt = grab(d.pt.p, d.pt.n); f.pq = grab(d.pq.p, d.pq.n); f.pap = grab(d.pap.p, d.pap.n);
    f.paph = grab(d.paph.p, d.paph.n); f.plu = grab(d.plu.p, d.plu.n); f.plude = grab(d.plude.p, d.plude.n);
    f.pmfu = grab(d.pmfu.p, d.pmfu.n); f.pmfd = grab(d.pmfd.p, d.pmfd.n);
    f.psupsat = grab(d.psupsat.p, d.psupsat.n); f.pclv = grab(d.pclv.p, d.pclv.n);
    f.b_cml = grab(d.b_cml.p, d.b_cml.n); f.b_loc = grab(d.b_loc.p, d.b_loc.n); f.pa = grab(d.pa.p, d.pa.n);
    f.pcovptot = grab(d.pcovptot.p, d.pcovptot.n); f.pfplsl = grab(d.pfplsl.p, d.pfplsl.n);
    f.pfplsn = grab(d.pfplsn.p, d.pfplsn.n); f.pfhpsl = grab(d.pfhpsl.p, d.pfhpsl.n);
    f.pfhpsn = grab(d.pfhpsn.p, d.pfhpsn.n);
  }
  void release() {
    for (void *h : owned) { if (pinned) cloudsc2_gpu_host_free(h); else std::free(h); }
    owned.clear();
  }
};

struct Options {
  Mode mode = NL;
  int numomp = 1, ngptotg = 16384, nproma = 32;   // dwarf_cloudsc.F90:27-29
  int numproc = 1, device0 = 0, repeat = 1, host_arrays = 0;
};

constexpr int NVAL = 10;   // validated fields, in the reference's order (cloudsc2_array_state_mod.F90:239-251)
const char *const VAL_NAME[NVAL] = {"PLUDE", "PCOVPTOT", "PFPLSL", "PFPLSN", "PFHPSL", "PFHPSN", "TENDENCY_LOC%A",
                                    "TENDENCY_LOC%Q", "TENDENCY_LOC%T", "TENDENCY_LOC%CLD"};
const int VAL_NDIM[NVAL] = {2, 2, 2, 2, 2, 2, 2, 2, 2, 3};

// What one rank hands to rank 0 (the reference gathers / reduces the same things over MPI:
// timer_mod.F90:160, validate_mod.F90:197-199, the drivers' reduction(max:znormg)).  Plain data: it
// crosses a pipe when NUMPROC > 1.
struct RankResult {
  int ngptot, nblocks, klon, klev;
  double seconds;
  double znormg_tl[10];
  double znormg_ad;
  double stats[NVAL][5];
  char input_line[200], ref_line[200];
};

// Everything one rank does: LOAD its shard, run the driver call, validate.  rank r owns the global
// columns [r*per, r*per + ngptot) with per = (NGPTOTG-1)/NUMPROC + 1 (dwarf_cloudsc.F90:63-67).
RankResult run_rank(const Options &o, int rank) {
  RankResult res;
  std::memset(&res, 0, sizeof res);
  const int per = (o.ngptotg - 1) / o.numproc + 1;
  const int ngptot = (rank == o.numproc - 1) ? o.ngptotg - (o.numproc - 1) * per : per;
  const long long gcol0 = (long long)rank * per;
  if (ngptot <= 0) abor1("more ranks than columns");
  const int nproma = o.nproma;

  if (!cloudsc2_gpu_available()) abor1("no CUDA device: the CLOUDSC2 GPU path has no CPU fallback");

  // ---- GLOBAL_STATE%LOAD, cloudsc2_array_state_mod.F90:153-203 -------------------------------------
  cloudsc2_source src;
  cloudsc2_params prm;
  const char *in_env = std::getenv("CLOUDSC2_INPUT");
  std::string in_path = (in_env && *in_env) ? in_env : (file_exists("input.h5") ? "input.h5" : "");
  if (!in_path.empty()) {
    if (cloudsc2_source_load_h5(&src, &prm, in_path.c_str()))
      abor1(std::string("cannot load ") + in_path + ": " + cloudsc2_input_last_error());
    std::snprintf(res.input_line, sizeof res.input_line, " input: %s (KLON=%d, KLEV=%d)", in_path.c_str(), src.klon, src.klev);
  } else {
    cloudsc2_default_params(&prm);
    const int seed = env_int("CLOUDSC2_SYNTH_SEED", 0);
    if (cloudsc2_source_synth(&src, (unsigned long long)seed, env_int("CLOUDSC2_SYNTH_KLON", 100),
                              env_int("CLOUDSC2_SYNTH_KLEV", 137), &prm))
      abor1("cannot build the synthetic input");
    std::snprintf(res.input_line, sizeof res.input_line,
                  " input: no input.h5 -- %d synthetic columns x %d levels, seed %d (IFS-standard constants)",
                  src.klon, src.klev, seed);
  }
  if (src.klev > 200) abor1("Dimension of ZPRES/ZPRESF is too short.");   // :88-91
  if (const char *wi = std::getenv("CLOUDSC2_WRITE_INPUT"))
    if (*wi && rank == 0 && cloudsc2_source_write_h5(&src, &prm, wi)) abor1(std::string("cannot write ") + wi);
  // dwarf_cloudsc.F90:105-107 and its twins: LEVAPLS2=.false., LPHYLIN=.true.; LREGCL per program
  prm.levapls2 = 0;
  prm.lphylin = 1;
  prm.ldrain1d = 0;
  prm.lregcl = (o.mode == AD) ? 1 : 0;   // cloudsc2_tl/dwarf_cloudsc.F90:105, cloudsc2_ad/dwarf_cloudsc.F90:105
  // rank -> device, wrapping when there are more ranks than GPUs (ranks then share a device)
  ck(cloudsc2_gpu_init(&prm, src.klev, src.ceta, (o.device0 + rank) % cloudsc2_gpu_device_count()), "cloudsc2_gpu_init");

  const int nblocks = cloudsc2_nblocks(ngptot, nproma);
  res.ngptot = ngptot; res.nblocks = nblocks; res.klon = src.klon; res.klev = src.klev;

  DeviceState dev;
  dev.load(src, nproma, ngptot, gcol0);
  HostState host;
  if (o.host_arrays) host.from_device(dev, o.host_arrays == 2);

  // ---- the driver: ONE library call instead of the OpenMP block loop ----------------------------------
  const int klev = src.klev;
  const double dt = src.ptsphy;
  double best = std::numeric_limits<double>::max();
  for (int it = 0; it < o.repeat + 1; ++it) {   // pass 0 is an untimed warm-up (module load, buffers)
    const double t0 = now_s();
    if (o.mode == NL) {
      if (o.host_arrays) ck(cloudsc2_gpu_nl(nproma, klev, ngptot, dt, &host.f, nullptr, nullptr), "cloudsc2_gpu_nl");
      else ck(cloudsc2_gpu_nl_dev(nproma, klev, ngptot, dt, &dev.f, nullptr, nullptr), "cloudsc2_gpu_nl_dev");
    } else if (o.mode == TL) {
      if (o.host_arrays) ck(cloudsc2_gpu_tl_taylor(nproma, klev, ngptot, dt, &host.f, res.znormg_tl, nullptr), "cloudsc2_gpu_tl_taylor");
      else ck(cloudsc2_gpu_tl_taylor_dev(nproma, klev, ngptot, dt, &dev.f, res.znormg_tl, nullptr), "cloudsc2_gpu_tl_taylor_dev");
    } else {
      if (o.host_arrays) ck(cloudsc2_gpu_ad_test(nproma, klev, ngptot, dt, &host.f, &res.znormg_ad, nullptr), "cloudsc2_gpu_ad_test");
      else ck(cloudsc2_gpu_ad_test_dev(nproma, klev, ngptot, dt, &dev.f, &res.znormg_ad, nullptr), "cloudsc2_gpu_ad_test_dev");
    }
    ck(cloudsc2_gpu_sync(), "cloudsc2_gpu_sync");
    if (it > 0) best = std::min(best, now_s() - t0);
  }
  res.seconds = best;

  if (o.mode == NL) {
    // ---- GLOBAL_STATE%VALIDATE, cloudsc2_array_state_mod.F90:205-252 ---------------------------------
    cloudsc2_reference ref;
    const char *ref_env = std::getenv("CLOUDSC2_REFERENCE");
    std::string ref_path = (ref_env && *ref_env) ? ref_env : (file_exists("reference.h5") && !in_path.empty() ? "reference.h5" : "");
    if (!ref_path.empty()) {
      if (cloudsc2_reference_load_h5(&ref, ref_path.c_str()))
        abor1(std::string("cannot load ") + ref_path + ": " + cloudsc2_input_last_error());
      if (ref.klon != src.klon || ref.klev != src.klev) abor1("reference.h5 and the input differ in KLON/KLEV");
      std::snprintf(res.ref_line, sizeof res.ref_line, " reference: %s", ref_path.c_str());
    } else {
      self_reference(src, ref);
      std::snprintf(res.ref_line, sizeof res.ref_line,
                    " reference: no reference.h5 -- the %d un-expanded columns run as one block", src.klon);
    }
    if (o.host_arrays) {   // bring the host results to the device once for the statistics
      const size_t n = (size_t)nproma * klev * nblocks, nh = (size_t)nproma * (klev + 1) * nblocks;
      ck(cloudsc2_gpu_memcpy_h2d(dev.pcovptot.p, host.f.pcovptot, n * 8), "h2d");
      ck(cloudsc2_gpu_memcpy_h2d(dev.pfplsl.p, host.f.pfplsl, nh * 8), "h2d");
      ck(cloudsc2_gpu_memcpy_h2d(dev.pfplsn.p, host.f.pfplsn, nh * 8), "h2d");
      ck(cloudsc2_gpu_memcpy_h2d(dev.pfhpsl.p, host.f.pfhpsl, nh * 8), "h2d");
      ck(cloudsc2_gpu_memcpy_h2d(dev.pfhpsn.p, host.f.pfhpsn, nh * 8), "h2d");
      ck(cloudsc2_gpu_memcpy_h2d(dev.b_loc.p, host.f.b_loc, n * CLOUDSC2_NSTATE * 8), "h2d");
    }
    const int klon = src.klon;
    const size_t n = (size_t)klon * klev;
    // TENDENCY_LOC%A/%Q/%T/%CLD = B_LOC(:,:,2,:), (:,:,3,:), (:,:,1,:), (:,:,4:,:)  (:248-251): slab
    // ranges of the AOSOA buffer, blocks 8*NPROMA*KLEV doubles apart.
    const long long bstride = (long long)CLOUDSC2_NSTATE * nproma * klev;
    const size_t slab = (size_t)nproma * klev;
    auto V = [&](int i, const double *r, const double *d, int nlev, int ndim, long long bs) {
      validate_field(r, klon, d, nproma, nlev, ndim, ngptot, gcol0, bs, res.stats[i]);
    };
    V(0, ref.plude, dev.plude.p, klev, 1, 0);
    V(1, ref.pcovptot, dev.pcovptot.p, klev, 1, 0);
    V(2, ref.pfplsl, dev.pfplsl.p, klev + 1, 1, 0);
    V(3, ref.pfplsn, dev.pfplsn.p, klev + 1, 1, 0);
    V(4, ref.pfhpsl, dev.pfhpsl.p, klev + 1, 1, 0);
    V(5, ref.pfhpsn, dev.pfhpsn.p, klev + 1, 1, 0);
    V(6, ref.tend_loc + 1 * n, dev.b_loc.p + 1 * slab, klev, 1, bstride);
    V(7, ref.tend_loc + 2 * n, dev.b_loc.p + 2 * slab, klev, 1, bstride);
    V(8, ref.tend_loc + 0 * n, dev.b_loc.p + 0 * slab, klev, 1, bstride);
    V(9, ref.tend_loc + 3 * n, dev.b_loc.p + 3 * slab, klev, CLOUDSC2_NCLV, bstride);
    cloudsc2_reference_free(&ref);
    // ---- GLOBAL_STATE%WRITE_REFERENCE, cloudsc2_array_state_mod.F90:260-287 --------------------------
    if (env_int("CLOUDSC2_WRITE_REFERENCE", 0) == 1 && rank == 0) {
      if (nproma != klon) abor1("[CLOUDSC2] Writing reference requires exactly NPROMA=KLON");   // :265-268
      cloudsc2_reference out;
      std::memset(&out, 0, sizeof out);
      out.klon = klon; out.klev = klev;
      auto fetch = [&](const double *d, size_t count) {
        double *h = static_cast<double *>(std::malloc(count * sizeof(double)));
        if (!h) abor1("out of memory");
        ck(cloudsc2_gpu_memcpy_d2h(h, d, count * sizeof(double)), "cloudsc2_gpu_memcpy_d2h");
        return h;
      };
      out.plude = fetch(dev.plude.p, n); out.pcovptot = fetch(dev.pcovptot.p, n);     // block 1 = first KLON columns
      out.pfplsl = fetch(dev.pfplsl.p, n + klon); out.pfplsn = fetch(dev.pfplsn.p, n + klon);
      out.pfhpsl = fetch(dev.pfhpsl.p, n + klon); out.pfhpsn = fetch(dev.pfhpsn.p, n + klon);
      out.tend_loc = fetch(dev.b_loc.p, n * CLOUDSC2_NSTATE);
      if (cloudsc2_reference_write_h5(&out, "reference.h5")) abor1("cannot write reference.h5");
      cloudsc2_reference_free(&out);
    }
  }
  dev.release();
  host.release();
  cloudsc2_source_free(&src);
  cloudsc2_gpu_finalize();
  return res;
}

// Rank 0's prints: the performance table (timer_mod.F90:114-174, unit 0) and the validation table or
// the test verdict.  Returns the exit status (0 = validated / TEST PASSED / TEST OK).
int report(const Options &o, const std::vector<RankResult> &r) {
  const int np = (int)r.size();
  const double zhpm = 3996006.0;   // cloudsc_driver_mod.F90:58
  std::fprintf(stderr, " %10s%10s%10s%10s%10s %4s : %10s%10s\n", "NUMOMP", "NGPTOT", "#GP-cols", "#BLKS",
               "NPROMA", "tid#", "Time(msec)", "MFlops/s");
  long long sum_cols = 0, sum_blks = 0, sum_mflops = 0, max_msec = 0;
  double max_s = 0.0;
  for (int p = 0; p < np; ++p) {
    const double s = r[p].seconds;
    const long long mflops = s > 0 ? (long long)(1.0e-06 * zhpm * (r[p].ngptot / 100.0) / s) : 0;
    const long long msec = (long long)(s * 1000.0);
    std::fprintf(stderr, " %10d%10d%10d%10d%10d %4d : %10lld%10lld : TOTAL @ rank#%d\n", o.numomp, r[p].ngptot,
                 r[p].ngptot, r[p].nblocks, o.nproma, -1, msec, mflops, p);
    sum_cols += r[p].ngptot; sum_blks += r[p].nblocks; sum_mflops += mflops;
    max_msec = std::max(max_msec, msec); max_s = std::max(max_s, s);
  }
  std::fprintf(stderr, " %6d x%2d%10lld%10lld%10lld%10d %4d : %10lld%10lld : TOTAL\n", np, o.numomp, sum_cols, sum_cols,
               sum_blks, o.nproma, -1, max_msec, sum_mflops);
  std::fprintf(stderr, "     GPU: %.3f ms per driver call = %.4g columns/s on %d GPU(s)\n", max_s * 1e3, sum_cols / max_s, np);

  std::printf("%s\n", r[0].input_line);
  if (o.mode == NL) {
    std::printf("%s\n", r[0].ref_line);
    std::printf(" %-20s %3s %20s %20s %20s %20s %20s\n", "Variable", "Dim", "MinValue", "MaxValue", "AbsMaxErr",
                "AvgAbsErr/GP", "MaxRelErr-%");
    int bad = 0;
    for (int i = 0; i < NVAL; ++i) {
      double st[5] = {r[0].stats[i][0], r[0].stats[i][1], r[0].stats[i][2], r[0].stats[i][3], r[0].stats[i][4]};
      for (int p = 1; p < np; ++p) {   // CLOUDSC_MPI_REDUCE_MIN / MAX / SUM, validate_mod.F90:197-199
        st[0] = std::min(st[0], r[p].stats[i][0]); st[1] = std::max(st[1], r[p].stats[i][1]);
        st[2] = std::max(st[2], r[p].stats[i][2]); st[3] += r[p].stats[i][3]; st[4] += r[p].stats[i][4];
      }
      bad += error_print(VAL_NAME[i], VAL_NDIM[i], st, o.ngptotg);
    }
    return bad ? 3 : 0;   // the reference only prints "!!!!"; a non-zero status makes it scriptable
  }
  if (o.mode == TL) {
    // ---- cloudsc_driver_tl_mod.F90:272-311; ZNORMG = max over blocks, hence over ranks (:125) ---------
    double z[10];
    for (int i = 0; i < 10; ++i) {
      z[i] = r[0].znormg_tl[i];
      for (int p = 1; p < np; ++p) z[i] = std::max(z[i], r[p].znormg_tl[i]);
    }
    std::printf("  TL Taylor test \n");
    std::printf("                 Lambda   Result\n");
    for (int i = 0; i < 10; ++i) std::printf(" %11d   %.16f\n", i + 1, z[i]);
    int istart = 0;
    const int pen = cloudsc2_taylor_verdict(z, &istart);
    std::printf("    ==============================================   \n");
    if (pen == -13) std::printf("        TEST FAILLED, err 13 \n");
    else if (pen > 5) std::printf("        TEST FAILLED, err %12d\n", pen);
    else std::printf("        TEST PASSED, penalty %12d\n", pen);
    std::printf("    ==============================================   \n");
    return (pen >= 0 && pen <= 5) ? 0 : 2;
  }
  // ---- cloudsc_driver_ad_mod.F90:285-294; reduction(max:znormg) :107 -----------------------------------
  double zn = r[0].znormg_ad;
  for (int p = 1; p < np; ++p) zn = std::max(zn, r[p].znormg_ad);
  std::printf("  AD TEST \n");
  std::printf("  The maximum error is %24.16f  times the zero of the machine. \n", zn);
  std::printf("    =============================  \n");
  const int ok = cloudsc2_adjoint_verdict(zn);
  std::printf(ok ? "    =           TEST OK         = \n" : "    =        TEST FAILED        = \n");
  std::printf("    =============================  \n");
  return ok ? 0 : 2;
}

}  // namespace

int main(int argc, char **argv) {
  Options o;
  // ---- which program am I? (three executables in the reference; here one source) ----------------
  std::string exe = argv[0];
  int first = 1;
  if (argc > 1 && std::strncmp(argv[1], "--mode=", 7) == 0) { exe = std::string("-") + (argv[1] + 7); first = 2; }
  const size_t slash = exe.find_last_of('/');
  const std::string base = slash == std::string::npos ? exe : exe.substr(slash + 1);
  if (base.find("-tl") != std::string::npos) o.mode = TL;
  else if (base.find("-ad") != std::string::npos) o.mode = AD;
  else if (base.find("-nl") != std::string::npos) o.mode = NL;
  else abor1("program name must contain -nl, -tl or -ad (or pass --mode=nl|tl|ad first)");
  if (argc > first && (!std::strcmp(argv[first], "-h") || !std::strcmp(argv[first], "--help"))) {
    std::printf("usage: dwarf-cloudsc2-{nl|tl|ad} [NUMOMP [NGPTOT [NPROMA]]]\n"
                "  defaults 1 16384 32 (dwarf_cloudsc.F90:27-29); NUMOMP is accepted and ignored\n"
                "  (one GPU does the whole block loop).  Environment: see the head of host/dwarf_cloudsc2.cc\n");
    return 0;
  }
  // ---- command line, dwarf_cloudsc.F90:49-75 ---------------------------------------------------------
  if (argc > first) o.numomp = parse_int_arg(argv[first], "NUMOMP");
  if (argc > first + 1) o.ngptotg = parse_int_arg(argv[first + 1], "NGPTOT");
  if (argc > first + 2) o.nproma = parse_int_arg(argv[first + 2], "NPROMA");
  o.numproc = std::max(1, env_int("CLOUDSC2_NUMPROC", 1));
  o.device0 = env_int("CLOUDSC2_DEVICE", 0);
  o.repeat = std::max(1, env_int("CLOUDSC2_REPEAT", 1));
  o.host_arrays = env_int("CLOUDSC2_HOST_ARRAYS", 0);

  // cloudsc_driver_mod.F90:63-66 (format 1003, unit 0); NGPBLKS of rank 0 like the reference
  const int per = (o.ngptotg - 1) / o.numproc + 1;
  std::fprintf(stderr, "     NUMPROC=%d, NUMOMP=%d, NGPTOTG=%d, NPROMA=%d, NGPBLKS=%d\n", o.numproc, o.numomp, o.ngptotg,
               o.nproma, cloudsc2_nblocks(o.numproc > 1 ? per : o.ngptotg, o.nproma));

  std::vector<RankResult> results(o.numproc);
  if (o.numproc == 1) {
    results[0] = run_rank(o, 0);
  } else {
    // One process per GPU, like the reference's MPI ranks (cloudsc_mpi_mod.F90).  The parent never
    // touches CUDA (a forked child cannot use a context created before the fork); every child sends its
    // RankResult through a pipe -- the gather / reductions of the reference are the few scalars in it.
    std::fflush(nullptr);
    std::vector<pid_t> pid(o.numproc);
    std::vector<int> fd(o.numproc);
    for (int r = 0; r < o.numproc; ++r) {
      int p[2];
      if (pipe(p) != 0) abor1("pipe() failed");
      pid[r] = fork();
      if (pid[r] < 0) abor1("fork() failed");
      if (pid[r] == 0) {
        close(p[0]);
        const RankResult res = run_rank(o, r);
        const char *b = reinterpret_cast<const char *>(&res);
        size_t left = sizeof res;
        while (left) {
          const ssize_t w = write(p[1], b, left);
          if (w <= 0) _exit(4);
          b += w; left -= (size_t)w;
        }
        close(p[1]);
        std::fflush(nullptr);
        _exit(0);
      }
      close(p[1]);
      fd[r] = p[0];
    }
    bool failed = false;
    for (int r = 0; r < o.numproc; ++r) {
      char *b = reinterpret_cast<char *>(&results[r]);
      size_t left = sizeof(RankResult);
      while (left) {
        const ssize_t g = read(fd[r], b, left);
        if (g <= 0) break;
        b += g; left -= (size_t)g;
      }
      close(fd[r]);
      int st = 0;
      waitpid(pid[r], &st, 0);
      if (left || !WIFEXITED(st) || WEXITSTATUS(st) != 0) failed = true;
    }
    if (failed) abor1("a rank failed (see its ABOR1 message above)");
  }
  return report(o, results);
}
