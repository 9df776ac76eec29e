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
//   CLOUDSC2_REFERENCE  path of reference.h5 for the NL validation (default ./reference.h5 next to an input.h5;
//                       for the default synthetic columns the shipped config-files/reference_synth_seed0.h5 =
//                       the CPU restatement's results; else, or with the value "self", the un-expanded columns
//                       computed as one block on the GPU are the reference: a self-consistency check only)
//   CLOUDSC2_WRITE_INPUT  path: write the columns and constants in use as an input.h5 (all datasets the
//                       reference's loaders read) before running -- to run the reference's binaries
//                       on the synthetic columns elsewhere
//   CLOUDSC2_WRITE_REFERENCE  1: after the NL run write ./reference.h5 from block 1, like the reference
//                       (dwarf_cloudsc.F90:124-126; needs NPROMA = KLON, cloudsc2_array_state_mod.F90:265-268)
//   CLOUDSC2_SYNTH_SEED / CLOUDSC2_SYNTH_KLON / CLOUDSC2_SYNTH_KLEV   synthetic input (0 / 100 / 137)
//   CLOUDSC2_DEVICE     CUDA device ordinal when one GPU is used (0)
//   CLOUDSC2_REPEAT     timed repetitions of the driver call, best one reported (1)
//   CLOUDSC2_HOST_ARRAYS  1: see above (pageable arrays, like a Fortran ALLOCATE); 2: page-locked arrays
//                       from cloudsc2_gpu_host_alloc; 3: malloc + cloudsc2_gpu_host_register (what
//                       INTEGRATION.md recommends to the unchanged Fortran host)
//   CLOUDSC2_NGPUS      number of GPUs driven by this ONE process (1; 0 = all visible; CLOUDSC2_NUMPROC is
//                       accepted as an alias).  The reference distributes NGPTOT over MPI ranks
//                       (dwarf_cloudsc.F90:63-67, cloudsc_mpi_mod.F90); here the library shards the NPROMA
//                       blocks over its device set with the same arithmetic (cloudsc2_gpu_init_multi), runs
//                       one host thread and one stream set per device, and all-reduces the test norms (MAX)
//                       and the validation statistics (MIN / MAX / SUM) on the devices with NCCL over NVLink.
//                       A device expands the source columns from its GLOBAL column offset, so results do not
//                       depend on the number of GPUs (the reference restarts every rank at source column 1).
//   CLOUDSC2_VALIDATE_TOL  exit-status tolerance of the NL validation on the relative L1 error (1e-11); the
//                       "!!!!" marks follow the reference (> 10 eps) regardless
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <utility>
#include <vector>

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

// directory of this executable (bin/), to find the shipped config-files/ next to it
std::string exe_dir() {
  char buf[4096];
  const ssize_t n = readlink("/proc/self/exe", buf, sizeof buf - 1);
  if (n <= 0) return ".";
  buf[n] = 0;
  std::string p(buf);
  const size_t slash = p.find_last_of('/');
  return slash == std::string::npos ? "." : p.substr(0, slash);
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
  bool finite = true;
  for (int i = 0; i < 5; ++i) finite = finite && std::isfinite(st[i]) && std::fabs(st[i]) < 1.0e300;
  const bool warn = !finite || rel > 10.0 * zeps;   // NaN > x is false: a non-finite statistic is always marked
  std::printf(" %-20s %1dD%1d %s %s %s %s %s%s\n", name, ndim, iopt, fortran_e(st[0]).c_str(),
              fortran_e(st[1]).c_str(), fortran_e(st[2]).c_str(),
              fortran_e(st[3] / (double)ngptotg).c_str(), fortran_e(100.0 * rel).c_str(), warn ? " !!!!" : "");
  return warn ? 1 : 0;
}

// The un-expanded columns run as ONE block of KLON columns on one device: the stand-in for
// reference.h5 when none is given.  That is a SELF-consistency check (blocking, expansion, sharding,
// reductions), not a verification against an independent reference, and is reported as such.
void self_reference(const cloudsc2_source &s, cloudsc2_reference &r) {
  ck(cloudsc2_gpu_state_load(&s, s.klon, s.klon), "cloudsc2_gpu_state_load");
  ck(cloudsc2_gpu_state_nl(nullptr, nullptr), "cloudsc2_gpu_state_nl");
  std::memset(&r, 0, sizeof r);
  r.klon = s.klon; r.klev = s.klev;
  const size_t n = (size_t)s.klon * s.klev, nh = n + s.klon;
  auto fetch = [&](const char *name, size_t count) {
    double *h = static_cast<double *>(std::malloc(count * sizeof(double)));
    if (!h) abor1("out of memory");
    ck(cloudsc2_gpu_state_get(name, h), "cloudsc2_gpu_state_get");
    return h;
  };
  r.plude = fetch("plude", n); r.pcovptot = fetch("pcovptot", n);
  r.pfplsl = fetch("pfplsl", nh); r.pfplsn = fetch("pfplsn", nh);
  r.pfhpsl = fetch("pfhpsl", nh); r.pfhpsn = fetch("pfhpsn", nh);
  r.tend_loc = fetch("b_loc", n * CLOUDSC2_NSTATE);
  ck(cloudsc2_gpu_state_free(), "cloudsc2_gpu_state_free");
}

// Blocked HOST arrays of the whole problem (what the Fortran host owns, expand_mod.F90:110,127,148):
// filled from the device-expanded state.  mode 1: malloc (pageable, like ALLOCATE); 2: page-locked
// memory from the library (cloudsc2_gpu_host_alloc); 3: malloc + cloudsc2_gpu_host_register.
struct HostState {
  cloudsc2_fields f{};
  std::vector<std::pair<void *, size_t>> owned;
  int mode = 1;
  double *grab(const char *name, size_t count) {
    void *h = nullptr;
    if (mode == 2) ck(cloudsc2_gpu_host_alloc(&h, count * sizeof(double)), "cloudsc2_gpu_host_alloc");
    else if (!(h = std::malloc(count * sizeof(double)))) abor1("out of memory");
    if (mode == 3) ck(cloudsc2_gpu_host_register(h, count * sizeof(double)), "cloudsc2_gpu_host_register");
    owned.push_back({h, count});
    ck(cloudsc2_gpu_state_get(name, static_cast<double *>(h)), "cloudsc2_gpu_state_get");
    return static_cast<double *>(h);
  }
  void from_state(int nproma, int klev, int nblocks, int m) {
    mode = m;
    const size_t n = (size_t)nproma * klev * nblocks, nh = (size_t)nproma * (klev + 1) * nblocks;
    f.pt = grab("pt", n); f.pq = grab("pq", n); f.pap = grab("pap", n); f.paph = grab("paph", nh);
    f.plu = grab("plu", n); f.plude = grab("plude", n); f.pmfu = grab("pmfu", n); f.pmfd = grab("pmfd", n);
    f.psupsat = grab("psupsat", n); f.pclv = grab("pclv", n * CLOUDSC2_NCLV);
    f.b_cml = grab("b_cml", n * CLOUDSC2_NSTATE); f.b_loc = grab("b_loc", n * CLOUDSC2_NSTATE);
    f.pa = grab("pa", n); f.pcovptot = grab("pcovptot", n); f.pfplsl = grab("pfplsl", nh);
    f.pfplsn = grab("pfplsn", nh); f.pfhpsl = grab("pfhpsl", nh); f.pfhpsn = grab("pfhpsn", nh);
  }
  void release() {
    for (auto &h : owned) {
      if (mode == 3) cloudsc2_gpu_host_unregister(h.first);
      if (mode == 2) cloudsc2_gpu_host_free(h.first); else std::free(h.first);
    }
    owned.clear();
  }
};

struct Options {
  Mode mode = NL;
  int numomp = 1, ngptotg = 16384, nproma = 32;   // dwarf_cloudsc.F90:27-29
  int ngpus = 1, repeat = 1, host_arrays = 0;
  double tol = 1.0e-11;
};

constexpr int NVAL = CLOUDSC2_NVALIDATED;   // in the reference's order (cloudsc2_array_state_mod.F90:239-251)
const char *const VAL_NAME[NVAL] = {"PLUDE", "PCOVPTOT", "PFPLSL", "PFPLSN", "PFHPSL", "PFHPSN", "TENDENCY_LOC%A",
                                    "TENDENCY_LOC%Q", "TENDENCY_LOC%T", "TENDENCY_LOC%CLD"};
const int VAL_NDIM[NVAL] = {2, 2, 2, 2, 2, 2, 2, 2, 2, 3};

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
                "  (the GPUs do the whole block loop).  Environment: see the head of host/dwarf_cloudsc2.cc\n");
    return 0;
  }
  // ---- command line, dwarf_cloudsc.F90:49-75 ---------------------------------------------------------
  if (argc > first) o.numomp = parse_int_arg(argv[first], "NUMOMP");
  if (argc > first + 1) o.ngptotg = parse_int_arg(argv[first + 1], "NGPTOT");
  if (argc > first + 2) o.nproma = parse_int_arg(argv[first + 2], "NPROMA");
  o.ngpus = env_int("CLOUDSC2_NGPUS", env_int("CLOUDSC2_NUMPROC", 1));    // 0: all visible
  o.repeat = std::max(1, env_int("CLOUDSC2_REPEAT", 1));
  o.host_arrays = env_int("CLOUDSC2_HOST_ARRAYS", 0);
  if (const char *t = std::getenv("CLOUDSC2_VALIDATE_TOL")) if (*t) o.tol = std::atof(t);
  const int nproma = o.nproma, ngptot = o.ngptotg;
  const int nblocks = cloudsc2_nblocks(ngptot, nproma);

  if (!cloudsc2_gpu_available()) abor1("no CUDA device: the CLOUDSC2 GPU path has no CPU fallback");
  if (o.ngpus <= 0) o.ngpus = cloudsc2_gpu_device_count();
  // cloudsc_driver_mod.F90:63-66 (format 1003, unit 0); one process, NUMPROC = 1 like a non-MPI build
  std::fprintf(stderr, "     NUMPROC=%d, NUMOMP=%d, NGPTOTG=%d, NPROMA=%d, NGPBLKS=%d   (GPUs: %d)\n", 1, o.numomp,
               ngptot, nproma, nblocks, o.ngpus);

  // ---- GLOBAL_STATE%LOAD, cloudsc2_array_state_mod.F90:153-203 -------------------------------------
  cloudsc2_source src;
  cloudsc2_params prm;
  char input_line[200], ref_line[240];
  const char *in_env = std::getenv("CLOUDSC2_INPUT");
  std::string in_path = (in_env && *in_env) ? in_env : (file_exists("input.h5") ? "input.h5" : "");
  if (!in_path.empty()) {
    if (cloudsc2_source_load_h5(&src, &prm, in_path.c_str()))
      abor1(std::string("cannot load ") + in_path + ": " + cloudsc2_input_last_error());
    std::snprintf(input_line, sizeof input_line, " input: %s (KLON=%d, KLEV=%d)", in_path.c_str(), src.klon, src.klev);
  } else {
    cloudsc2_default_params(&prm);
    const int seed = env_int("CLOUDSC2_SYNTH_SEED", 0);
    if (cloudsc2_source_synth(&src, (unsigned long long)seed, env_int("CLOUDSC2_SYNTH_KLON", 100),
                              env_int("CLOUDSC2_SYNTH_KLEV", 137), &prm))
      abor1("cannot build the synthetic input");
    std::snprintf(input_line, sizeof input_line,
                  " input: no input.h5 -- %d synthetic columns x %d levels, seed %d (IFS-standard constants)",
                  src.klon, src.klev, seed);
  }
  if (src.klev > 200) abor1("Dimension of ZPRES/ZPRESF is too short.");   // :88-91
  if (const char *wi = std::getenv("CLOUDSC2_WRITE_INPUT"))
    if (*wi && cloudsc2_source_write_h5(&src, &prm, wi)) abor1(std::string("cannot write ") + wi);
  // dwarf_cloudsc.F90:105-107 and its twins: LEVAPLS2=.false., LPHYLIN=.true.; LREGCL per program
  prm.levapls2 = 0;
  prm.lphylin = 1;
  prm.ldrain1d = 0;
  prm.lregcl = (o.mode == AD) ? 1 : 0;   // cloudsc2_tl/dwarf_cloudsc.F90:105, cloudsc2_ad/dwarf_cloudsc.F90:105
  // ONE process drives all GPUs: contexts, worker threads and the NCCL communicator live in the library
  if (o.ngpus > 1) ck(cloudsc2_gpu_init_multi(&prm, src.klev, src.ceta, o.ngpus), "cloudsc2_gpu_init_multi");
  else ck(cloudsc2_gpu_init(&prm, src.klev, src.ceta, env_int("CLOUDSC2_DEVICE", 0)), "cloudsc2_gpu_init");
  const int ndev = cloudsc2_gpu_num_devices();
  const int klev = src.klev;
  const double dt = src.ptsphy;

  // ---- the reference columns (NL only) -- before the big state takes the memory ----------------------
  cloudsc2_reference ref;
  std::memset(&ref, 0, sizeof ref);
  bool independent_ref = false;
  if (o.mode == NL) {
    const char *ref_env = std::getenv("CLOUDSC2_REFERENCE");
    std::string ref_path = (ref_env && *ref_env) ? ref_env : (file_exists("reference.h5") && !in_path.empty() ? "reference.h5" : "");
    const bool force_self = ref_path == "self";
    if (force_self) ref_path.clear();
    bool shipped = false;
    if (ref_path.empty() && !force_self && in_path.empty() && src.klon == 100 && src.klev == 137 &&
        env_int("CLOUDSC2_SYNTH_SEED", 0) == 0) {
      // the default synthetic columns have a reference of their own: the CPU oracle's results (pinned to the
      // reference's Fortran text), shipped like the reference ships config-files/reference.h5
      const std::string cand = exe_dir() + "/../config-files/reference_synth_seed0.h5";
      if (file_exists(cand.c_str())) { ref_path = cand; shipped = true; }
    }
    if (!ref_path.empty()) {
      if (cloudsc2_reference_load_h5(&ref, ref_path.c_str()))
        abor1(std::string("cannot load ") + ref_path + ": " + cloudsc2_input_last_error());
      if (ref.klon != src.klon || ref.klev != src.klev) abor1("reference.h5 and the input differ in KLON/KLEV");
      if (shipped)
        std::snprintf(ref_line, sizeof ref_line, " reference: config-files/reference_synth_seed0.h5 (the synthetic columns run by the CPU "
                      "restatement of the reference, NPROMA=KLON)");
      else
        std::snprintf(ref_line, sizeof ref_line, " reference: %s", ref_path.c_str());
      independent_ref = true;
    } else {
      self_reference(src, ref);
      std::snprintf(ref_line, sizeof ref_line,
                    " reference: no reference.h5 -- the %d un-expanded columns run as one block: SELF-CONSISTENCY only, "
                    "results UNVERIFIED against an independent reference", src.klon);
    }
  }

  ck(cloudsc2_gpu_state_load(&src, nproma, ngptot), "cloudsc2_gpu_state_load");
  HostState host;
  if (o.host_arrays) host.from_state(nproma, klev, nblocks, o.host_arrays);

  // ---- the driver: ONE library call instead of the OpenMP block loop ----------------------------------
  double znormg_tl[10] = {0}, znormg_ad = 0.0;
  std::vector<double> per_dev(std::max(1, ndev), 0.0), best_dev(std::max(1, ndev), 0.0);
  double best = std::numeric_limits<double>::max();
  for (int it = 0; it < o.repeat + 1; ++it) {   // pass 0 is an untimed warm-up (module load, buffers)
    const double t0 = now_s();
    double t_lib = 0.0;
    if (o.mode == NL) {
      if (o.host_arrays) ck(cloudsc2_gpu_nl(nproma, klev, ngptot, dt, &host.f, nullptr, nullptr), "cloudsc2_gpu_nl");
      else ck(cloudsc2_gpu_state_nl(&t_lib, per_dev.data()), "cloudsc2_gpu_state_nl");
    } else if (o.mode == TL) {
      if (o.host_arrays) ck(cloudsc2_gpu_tl_taylor(nproma, klev, ngptot, dt, &host.f, znormg_tl, nullptr), "cloudsc2_gpu_tl_taylor");
      else ck(cloudsc2_gpu_state_tl_taylor(znormg_tl, &t_lib, per_dev.data()), "cloudsc2_gpu_state_tl_taylor");
    } else {
      if (o.host_arrays) ck(cloudsc2_gpu_ad_test(nproma, klev, ngptot, dt, &host.f, &znormg_ad, nullptr), "cloudsc2_gpu_ad_test");
      else ck(cloudsc2_gpu_state_ad_test(&znormg_ad, &t_lib, per_dev.data()), "cloudsc2_gpu_state_ad_test");
    }
    const double wall = now_s() - t0;
    // resident state: the library's own clock (CUDA events / slowest device); host arrays: the whole call
    const double t = (o.host_arrays || t_lib <= 0.0) ? wall : t_lib;
    if (it > 0 && t < best) { best = t; best_dev = per_dev; }
  }

  // ---- performance table (timer_mod.F90:114-174, unit 0): one line per GPU, then the total -----------
  const double zhpm = 3996006.0;   // cloudsc_driver_mod.F90:58
  std::fprintf(stderr, " %10s%10s%10s%10s%10s %4s : %10s%10s\n", "NUMOMP", "NGPTOT", "#GP-cols", "#BLKS",
               "NPROMA", "tid#", "Time(msec)", "MFlops/s");
  for (int d = 0; d < ndev; ++d) {
    int ord = 0, nb = 0, ng = 0; long long g0 = 0;
    ck(cloudsc2_gpu_state_info(d, &ord, &nb, &ng, &g0), "cloudsc2_gpu_state_info");
    const double s = (o.host_arrays || best_dev[d] <= 0.0) ? best : best_dev[d];
    const long long mflops = s > 0 ? (long long)(1.0e-06 * zhpm * (ng / 100.0) / s) : 0;
    std::fprintf(stderr, " %10d%10d%10d%10d%10d %4d : %10lld%10lld : GPU %d\n", o.numomp, ngptot, ng, nb, nproma,
                 d, (long long)(s * 1000.0), mflops, ord);
  }
  {
    const long long mflops = best > 0 ? (long long)(1.0e-06 * zhpm * (ngptot / 100.0) / best) : 0;
    std::fprintf(stderr, " %6d x%2d%10d%10d%10d%10d %4d : %10lld%10lld : TOTAL\n", 1, o.numomp, ngptot, ngptot, nblocks,
                 nproma, -1, (long long)(best * 1000.0), mflops);
  }
  {
    int crank = 0, csize = 1, cver = 0;
    cloudsc2_gpu_comm_info(&crank, &csize, &cver);
    std::fprintf(stderr, "     GPU: %.3f ms per driver call = %.4g columns/s on %d GPU(s), one process; NCCL %d ranks=%d\n",
                 best * 1e3, ngptot / best, ndev, cver, csize);
  }

  int status = 0;
  std::printf("%s\n", input_line);
  if (o.mode == NL) {
    // ---- GLOBAL_STATE%VALIDATE, cloudsc2_array_state_mod.F90:205-252 ---------------------------------
    if (o.host_arrays) {
      // the host arrays hold the results: put them into the resident state's outputs, shard by shard, so
      // that the device-side statistics see exactly what the host-pointer entry returned
      const size_t n2 = (size_t)nproma * klev, n2h = (size_t)nproma * (klev + 1);
      for (int d = 0; d < ndev; ++d) {
        int ord = 0, nb = 0, ng = 0; long long g0 = 0;
        ck(cloudsc2_gpu_state_info(d, &ord, &nb, &ng, &g0), "cloudsc2_gpu_state_info");
        if (!nb) continue;
        ck(cloudsc2_gpu_select_device(d), "cloudsc2_gpu_select_device");
        cloudsc2_fields f;
        ck(cloudsc2_gpu_state_fields(&f), "cloudsc2_gpu_state_fields");
        const size_t b0 = (size_t)(g0 / nproma);
        auto put = [&](double *dev, const double *h, size_t per_blk) {
          ck(cloudsc2_gpu_memcpy_h2d(dev, h + per_blk * b0, per_blk * nb * sizeof(double)), "cloudsc2_gpu_memcpy_h2d");
        };
        put(f.b_loc, host.f.b_loc, CLOUDSC2_NSTATE * n2); put(f.pcovptot, host.f.pcovptot, n2);
        put(f.pfplsl, host.f.pfplsl, n2h); put(f.pfplsn, host.f.pfplsn, n2h);
        put(f.pfhpsl, host.f.pfhpsl, n2h); put(f.pfhpsn, host.f.pfhpsn, n2h);
      }
      ck(cloudsc2_gpu_select_device(-1), "cloudsc2_gpu_select_device");
    }
    double stats[NVAL][5];
    ck(cloudsc2_gpu_state_validate(&ref, &stats[0][0]), "cloudsc2_gpu_state_validate");
    std::printf("%s\n", ref_line);
    std::printf(" %-20s %3s %20s %20s %20s %20s %20s\n", "Variable", "Dim", "MinValue", "MaxValue", "AbsMaxErr",
                "AvgAbsErr/GP", "MaxRelErr-%");
    int marks = 0, bad = 0;
    for (int i = 0; i < NVAL; ++i) {
      marks += error_print(VAL_NAME[i], VAL_NDIM[i], stats[i], ngptot);
      // exit status: the reference only prints "!!!!" above 10 eps -- which every field of a GPU run against a
      // CPU-made reference.h5 exceeds (exp / division / FMA differ in the last bits).  The status uses its own,
      // documented tolerance on the relative L1 error (CLOUDSC2_VALIDATE_TOL, default 1e-11), and any
      // non-finite statistic fails.
      const double *st = stats[i];
      bool finite = true;
      for (int k = 0; k < 5; ++k) finite = finite && std::isfinite(st[k]) && std::fabs(st[k]) < 1.0e300;
      const double rel = st[3] / (st[4] > 0.0 ? st[4] : 1.0 + st[4]);
      if (!finite || !(rel <= o.tol)) ++bad;
    }
    std::printf(" validation: %d field(s) above the exit-status tolerance %.1e (relative L1), %d marked '!!!!' (> 10 eps)%s\n",
                bad, o.tol, marks, independent_ref ? "" : "; self-consistency only: UNVERIFIED");
    status = bad ? 3 : 0;
    // ---- GLOBAL_STATE%WRITE_REFERENCE, cloudsc2_array_state_mod.F90:260-287 --------------------------
    if (env_int("CLOUDSC2_WRITE_REFERENCE", 0) == 1) {
      const int klon = src.klon;
      if (nproma != klon) abor1("[CLOUDSC2] Writing reference requires exactly NPROMA=KLON");   // :265-268
      cloudsc2_reference out;
      std::memset(&out, 0, sizeof out);
      out.klon = klon; out.klev = klev;
      auto fetch = [&](const char *name, size_t blk) {     // block 1 = the first KLON columns
        std::vector<double> all(blk * nblocks);
        ck(cloudsc2_gpu_state_get(name, all.data()), "cloudsc2_gpu_state_get");
        double *h = static_cast<double *>(std::malloc(blk * sizeof(double)));
        if (!h) abor1("out of memory");
        std::memcpy(h, all.data(), blk * sizeof(double));
        return h;
      };
      const size_t n = (size_t)klon * klev;
      out.plude = fetch("plude", n); out.pcovptot = fetch("pcovptot", n);
      out.pfplsl = fetch("pfplsl", n + klon); out.pfplsn = fetch("pfplsn", n + klon);
      out.pfhpsl = fetch("pfhpsl", n + klon); out.pfhpsn = fetch("pfhpsn", n + klon);
      out.tend_loc = fetch("b_loc", n * CLOUDSC2_NSTATE);
      if (cloudsc2_reference_write_h5(&out, "reference.h5")) abor1("cannot write reference.h5");
      cloudsc2_reference_free(&out);
    }
    cloudsc2_reference_free(&ref);
  } else if (o.mode == TL) {
    // ---- cloudsc_driver_tl_mod.F90:272-311; ZNORMG = max over blocks, all-reduced on the devices (:125) --
    std::printf("  TL Taylor test \n");
    std::printf("                 Lambda   Result\n");
    for (int i = 0; i < 10; ++i) std::printf(" %11d   %.16f\n", i + 1, znormg_tl[i]);
    int istart = 0;
    const int pen = cloudsc2_taylor_verdict(znormg_tl, &istart);
    std::printf("    ==============================================   \n");
    if (pen == -13) std::printf("        TEST FAILLED, err 13 \n");
    else if (pen > 5) std::printf("        TEST FAILLED, err %12d\n", pen);
    else std::printf("        TEST PASSED, penalty %12d\n", pen);
    std::printf("    ==============================================   \n");
    status = (pen >= 0 && pen <= 5) ? 0 : 2;
  } else {
    // ---- cloudsc_driver_ad_mod.F90:285-294; reduction(max:znormg) :107 ---------------------------------
    std::printf("  AD TEST \n");
    std::printf("  The maximum error is %24.16f  times the zero of the machine. \n", znormg_ad);
    std::printf("    =============================  \n");
    const int ok = cloudsc2_adjoint_verdict(znormg_ad);
    std::printf(ok ? "    =           TEST OK         = \n" : "    =        TEST FAILED        = \n");
    std::printf("    =============================  \n");
    status = ok ? 0 : 2;
  }
  host.release();
  cloudsc2_gpu_state_free();
  cloudsc2_source_free(&src);
  cloudsc2_gpu_finalize();
  return status;
}
