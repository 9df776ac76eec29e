// cloudsc2_multi.cu -- the device set: life cycle, per-device contexts and worker threads, the NCCL
// communicator, block sharding of the host-pointer entry points over the GPUs of one box, and the
// device-resident sharded state built from the un-expanded source columns.
//
// Reference behaviour this replaces: MPI ranks over columns (cloudsc2_nl/dwarf_cloudsc.F90:65-69,
// common/module/cloudsc_mpi_mod.F90) with OpenMP threads over blocks inside a rank
// (cloudsc_driver_mod.F90:73-81), `reduction(max:znormg)` of the test norms
// (cloudsc_driver_tl_mod.F90:125, cloudsc_driver_ad_mod.F90:107) and CLOUDSC_MPI_REDUCE_MIN/MAX/SUM
// of the validation statistics (validate_mod.F90:197-199).  Here ONE process drives all GPUs: one
// persistent host thread and one stream set per device, contiguous block ranges per device, and
// one NCCL all-reduce over NVLink of the device-resident scalars -- no data-path collective.
// NCCL is bound at run time (dlopen of libnccl.so.2): a single-GPU host needs no NCCL at all, and
// a host that already carries its own NCCL (PyTorch) shares that copy.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/cloudsc2_host.h"
#include "cloudsc2_ctx.h"

// single-context implementations (cloudsc2_api.cu)
extern "C" {
int csc2_nl_host_one(int nproma, int klev, int ngptot, double ptsphy, const cloudsc2_fields *h,
                     double *elapsed_kernel_s, double *elapsed_total_s);
int csc2_tlad_host_one(bool is_ad, int nproma, int klev, int ngptot, double ptsphy, const cloudsc2_fields *h,
                       const cloudsc2_incr_in *a, const cloudsc2_incr_out *b);
int csc2_taylor_host_one(int nproma, int klev, int ngptot, double ptsphy, const cloudsc2_fields *h,
                         double znormg[10], double *ratios_blk);
int csc2_adtest_host_one(int nproma, int klev, int ngptot, double ptsphy, const cloudsc2_fields *h,
                         double *znormg, double *norms_col);
}

namespace {

thread_local char tl_err[1024] = "";
thread_local Ctx *tl_ctx = nullptr;
thread_local bool tl_is_worker = false;

// ---- NCCL, bound at run time ---------------------------------------------------------------------
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  char why[256] = "";
};
NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names)
      if ((api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!api.handle) {
      snprintf(api.why, sizeof api.why, "cannot load libnccl.so.2: %s", dlerror());
      return;
    }
    auto sym = [&](const char *s) { return dlsym(api.handle, s); };
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitAll || !api.CommInitRank || !api.CommDestroy || !api.AllReduce ||
        !api.GetErrorString) {
      snprintf(api.why, sizeof api.why, "libnccl.so.2 lacks a required symbol");
      api.handle = nullptr;
    }
  });
  return api.handle ? &api : nullptr;
}
#define NK(call)                                                                                     \
  do {                                                                                               \
    ncclResult_t r_ = (call);                                                                        \
    if (r_ != ncclSuccess)                                                                           \
      return csc2_fail(300 + (int)r_, "%s failed: %s (%s:%d)", #call, nccl_api()->GetErrorString(r_), \
                       __FILE__, __LINE__);                                                          \
  } while (0)

// ---- worker threads --------------------------------------------------------------------------------
struct Worker {
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  int (*fn)(int, void *) = nullptr;
  void *arg = nullptr;
  bool has_job = false, done = false, quit = false;
  int rc = 0;
  char err[1024] = "";
};

struct DeviceSet {
  int n = 0;
  std::unique_ptr<Ctx> ctx[CSC2_MAX_DEVICES];
  std::unique_ptr<Worker> worker[CSC2_MAX_DEVICES];
  bool inproc_comm = false;     // communicator made by ncclCommInitAll over the set
  std::mutex call_mu;           // one multi-device call at a time
  // host barrier of csc2_agree
  std::mutex bar_mu;
  std::condition_variable bar_cv;
  int bar_count = 0, bar_any = 0, bar_result = 0;
  unsigned long long bar_gen = 0;
  void stop_workers() {
    for (int i = 0; i < CSC2_MAX_DEVICES; ++i) {
      if (!worker[i]) continue;
      Worker &w = *worker[i];
      {
        std::lock_guard<std::mutex> lk(w.m);
        w.quit = true;
      }
      w.cv.notify_all();
      if (w.th.joinable()) w.th.join();
      worker[i].reset();
    }
  }
  // A host that exits without cloudsc2_gpu_finalize must not die in ~thread (joinable): stop the workers;
  // device memory and streams go with the process (the CUDA runtime may already be unloading here).
  ~DeviceSet() { stop_workers(); }
};
DeviceSet set_;
Ctx no_ctx;                      // what csc2_ctx() returns before any init (init == false)

void worker_main(int index) {
  Worker &w = *set_.worker[index];
  Ctx &c = *set_.ctx[index];
  tl_ctx = &c;
  tl_is_worker = true;
  cudaSetDevice(c.device);
  std::unique_lock<std::mutex> lk(w.m);
  for (;;) {
    w.cv.wait(lk, [&] { return w.has_job || w.quit; });
    if (w.quit) return;
    w.has_job = false;
    lk.unlock();
    tl_err[0] = 0;
    const int rc = w.fn(index, w.arg);
    lk.lock();
    w.rc = rc;
    std::memcpy(w.err, tl_err, sizeof w.err);
    w.done = true;
    w.cv.notify_all();
  }
}

int init_ctx(Ctx &c, const cloudsc2_params *params, int klev, const double *ceta, int device, int index) {
  CK(cudaSetDevice(device));
  c.device = device;
  c.index = index;
  c.prm = *params;
  c.klev = klev;
  std::memcpy(c.ceta, ceta, sizeof(double) * klev);
  c.kwin0 = 0; c.kwin1 = -1;
  bool any = false;
  for (int jk = 0; jk < klev - 1; ++jk) {        // DO JK=1,KLEV-1 (cloudsc2.F90:318)
    if (ceta[jk] > 0.1 && ceta[jk] < 0.4) {
      if (!any) c.kwin0 = jk;
      c.kwin1 = jk;
      any = true;
    }
  }
  for (int jk = 0; jk < klev; ++jk)              // cloudsc2.F90:266, ZSCAL = 0.9 (:172)
    c.zscalm[jk] = 0.9 * std::pow(std::max(ceta[jk] - 0.2, 1.e-12), 0.2);
  CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  for (int i = 0; i < kStreams; ++i) CK(cudaStreamCreateWithFlags(&c.pipe[i], cudaStreamNonBlocking));
  for (int i = 0; i < 4; ++i) CK(cudaEventCreate(&c.ev[i]));
  CK(cudaEventCreateWithFlags(&c.scratch_done, cudaEventDisableTiming));
  {
    // per-level constants -> __constant__ tables of the three kernel translation units (per device)
    std::vector<double> sq(klev);
    for (int k = 0; k < klev; ++k) sq[k] = std::sqrt(std::max(1.0 - ceta[k], 0.0));   // cloudsc2.F90:398
    CK(csc2_upload_levels_nl(c.ceta, c.zscalm, sq.data(), klev, c.stream));
    CK(csc2_upload_levels_tl(c.ceta, c.zscalm, sq.data(), klev, c.stream));
    CK(csc2_upload_levels_ad(c.ceta, c.zscalm, sq.data(), klev, c.stream));
  }
  c.launches = 0;
  c.comm = nullptr;
  c.comm_rank = 0;
  c.comm_size = 1;
  c.init = true;
  return 0;
}

void destroy_ctx(Ctx &c) {
  if (!c.init) return;
  cudaSetDevice(c.device);
  cudaDeviceSynchronize();
  if (c.comm) {
    if (NcclApi *n = nccl_api()) n->CommDestroy(static_cast<ncclComm_t>(c.comm));
    c.comm = nullptr;
  }
  c.in.release(); c.out.release(); c.work.release(); c.work2.release(); c.res.release();
  c.state.mem.release();
  c.state.loaded = false;
  if (c.stream) cudaStreamDestroy(c.stream);
  for (int i = 0; i < kStreams; ++i) if (c.pipe[i]) cudaStreamDestroy(c.pipe[i]);
  for (int i = 0; i < 4; ++i) if (c.ev[i]) cudaEventDestroy(c.ev[i]);
  if (c.scratch_done) cudaEventDestroy(c.scratch_done);
  c.stream = nullptr;
  c.scratch_done = nullptr;
  for (int i = 0; i < kStreams; ++i) c.pipe[i] = nullptr;
  for (int i = 0; i < 4; ++i) c.ev[i] = nullptr;
  c.init = false;
}

void teardown() {
  set_.stop_workers();
  for (int i = 0; i < set_.n; ++i) {
    if (set_.ctx[i]) destroy_ctx(*set_.ctx[i]);
    set_.ctx[i].reset();
  }
  set_.n = 0;
  set_.inproc_comm = false;
  tl_ctx = nullptr;
}

int check_params(const cloudsc2_params *params, int klev, const double *ceta) {
  if (!params || !ceta) return csc2_fail(3, "cloudsc2_gpu_init: NULL argument");
  if (klev < 2 || klev > CSC2_KLEV_MAX) return csc2_fail(3, "klev=%d outside [2,%d]", klev, CSC2_KLEV_MAX);
  // Only the configuration the three dwarf programs run is implemented on the device
  // (cloudsc2_{nl,tl,ad}/dwarf_cloudsc.F90:105-107; LDRAIN1D = .FALSE. in every driver).
  if (!params->lphylin) return csc2_fail(4, "LPHYLIN=.FALSE. is not supported (the dwarf forces .TRUE.)");
  if (params->levapls2 || params->ldrain1d)
    return csc2_fail(4, "LEVAPLS2/LDRAIN1D=.TRUE. (precipitation evaporation) is not supported");
  if (!cloudsc2_gpu_available()) return csc2_fail(5, "no CUDA device available (there is no CPU fallback)");
  return 0;
}

int init_devices(const cloudsc2_params *params, int klev, const double *ceta, int ndev, const int *devices) {
  if (int rc = check_params(params, klev, ceta)) return rc;
  if (ndev < 1 || ndev > CSC2_MAX_DEVICES) return csc2_fail(3, "bad number of devices %d", ndev);
  teardown();
  for (int i = 0; i < ndev; ++i) {
    set_.ctx[i].reset(new Ctx());
    set_.n = i + 1;
    if (int rc = init_ctx(*set_.ctx[i], params, klev, ceta, devices[i], i)) {
      char keep[1024];
      std::memcpy(keep, tl_err, sizeof keep);
      teardown();
      csc2_set_error_text(keep);
      return rc;
    }
  }
  if (ndev > 1) {
    NcclApi *n = nccl_api();
    if (!n) { teardown(); return csc2_fail(7, "NCCL is needed to drive %d devices and cannot be loaded (libnccl.so.2)", ndev); }
    ncclComm_t comms[CSC2_MAX_DEVICES];
    ncclResult_t r = n->CommInitAll(comms, ndev, devices);
    if (r != ncclSuccess) {
      teardown();
      return csc2_fail(300 + (int)r, "ncclCommInitAll over %d devices failed: %s", ndev, n->GetErrorString(r));
    }
    for (int i = 0; i < ndev; ++i) {
      set_.ctx[i]->comm = comms[i];
      set_.ctx[i]->comm_rank = i;
      set_.ctx[i]->comm_size = ndev;
    }
    set_.inproc_comm = true;
    for (int i = 0; i < ndev; ++i) {
      set_.worker[i].reset(new Worker());
      set_.worker[i]->th = std::thread(worker_main, i);
    }
  }
  cudaSetDevice(set_.ctx[0]->device);
  return 0;
}

cloudsc2_fields offset_fields(const cloudsc2_fields &f, int b0, int nproma, int klev) {
  const size_t n2 = (size_t)nproma * klev * b0, n2h = (size_t)nproma * (klev + 1) * b0;
  cloudsc2_fields o;
  o.pt = f.pt + n2; o.pq = f.pq + n2; o.pap = f.pap + n2; o.paph = f.paph + n2h; o.plu = f.plu + n2;
  o.plude = f.plude + n2; o.pmfu = f.pmfu + n2; o.pmfd = f.pmfd + n2; o.psupsat = f.psupsat + n2;
  o.pclv = f.pclv + CLOUDSC2_NCLV * n2; o.b_cml = f.b_cml + CLOUDSC2_NSTATE * n2;
  o.b_loc = f.b_loc + CLOUDSC2_NSTATE * n2; o.pa = f.pa + n2; o.pcovptot = f.pcovptot + n2;
  o.pfplsl = f.pfplsl + n2h; o.pfplsn = f.pfplsn + n2h; o.pfhpsl = f.pfhpsl + n2h; o.pfhpsn = f.pfhpsn + n2h;
  return o;
}

int check_host_call(int nproma, int klev, int ngptot, const cloudsc2_fields *h) {
  if (int rc = csc2_require_init()) return rc;
  if (nproma <= 0 || ngptot <= 0) return csc2_fail(3, "bad dimensions nproma=%d ngptot=%d", nproma, ngptot);
  if (!h) return csc2_fail(3, "fields pointer is NULL");
  (void)klev;
  return 0;
}

// A rank that owns no block still takes part in the collectives, with the identities.
int reduce_identity(Ctx &c, int nmax, int nsum) {
  if (int rc = csc2_agree(c.res.reserve(32 * sizeof(double)))) return rc;
  double *d = c.res.d();
  CK(csc2_launch_fill(d, nmax, -1.7976931348623157e308, c.stream));
  if (int rc = csc2_allreduce(c, d, nmax, 0, c.stream)) return rc;
  if (nsum) {
    CK(csc2_launch_fill(d + 11, nsum, 0.0, c.stream));
    if (int rc = csc2_allreduce(c, d + 11, nsum, 2, c.stream)) return rc;
  }
  CK(cudaStreamSynchronize(c.stream));
  return 0;
}

}  // namespace

// ---- internal interface (cloudsc2_ctx.h) -----------------------------------------------------------
int csc2_fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_err, sizeof(tl_err), fmt, ap);
  va_end(ap);
  return code;
}
const char *csc2_error_text() { return tl_err; }
void csc2_set_error_text(const char *s) { snprintf(tl_err, sizeof tl_err, "%s", s); }

Ctx &csc2_ctx() {
  if (tl_ctx) return *tl_ctx;
  if (set_.n > 0) return *set_.ctx[0];
  return no_ctx;
}
bool csc2_ctx_ready() { return csc2_ctx().init; }
int csc2_require_init() {
  Ctx &c = csc2_ctx();
  if (!c.init) return csc2_fail(2, "cloudsc2_gpu_init has not been called");
  CK(cudaSetDevice(c.device));      // the calling thread may be new (e.g. an OpenMP worker of the host)
  return 0;
}
int csc2_num_devices() { return set_.n; }
Ctx &csc2_ctx_at(int index) { return *set_.ctx[index]; }

int csc2_on_all_devices(int (*fn)(int, void *), void *arg) {
  if (set_.n <= 1 || tl_ctx) return fn(csc2_ctx().index, arg);     // single device, or already on a worker
  std::lock_guard<std::mutex> call(set_.call_mu);
  for (int i = 0; i < set_.n; ++i) {
    Worker &w = *set_.worker[i];
    std::lock_guard<std::mutex> lk(w.m);
    w.fn = fn; w.arg = arg; w.done = false; w.has_job = true;
    w.cv.notify_all();
  }
  int rc = 0;
  for (int i = 0; i < set_.n; ++i) {
    Worker &w = *set_.worker[i];
    std::unique_lock<std::mutex> lk(w.m);
    w.cv.wait(lk, [&] { return w.done; });
    if (w.rc && !rc) {
      rc = w.rc;
      char msg[1024];
      snprintf(msg, sizeof msg, "device %d: %.1000s", set_.ctx[i]->device, w.err);
      csc2_set_error_text(msg);
    }
  }
  return rc;
}

Shard csc2_shard(int index, int ndev, int nproma, int ngptot) {
  const int nblocks = ngptot / nproma + std::min(ngptot % nproma, 1);
  const int per = (nblocks - 1) / ndev + 1;              // dwarf_cloudsc.F90:65, on blocks
  Shard s;
  s.b0 = std::min(index * per, nblocks);
  s.nb = std::min(nblocks, s.b0 + per) - s.b0;
  s.gcol0 = (long long)s.b0 * nproma;
  s.ngptot = s.nb ? (int)std::min<long long>((long long)s.nb * nproma, (long long)ngptot - s.gcol0) : 0;
  return s;
}

int csc2_agree(int rc) {
  if (!tl_is_worker || set_.n <= 1) return rc;
  std::unique_lock<std::mutex> lk(set_.bar_mu);
  if (rc) set_.bar_any = 1;
  const unsigned long long gen = set_.bar_gen;
  if (++set_.bar_count == set_.n) {
    set_.bar_result = set_.bar_any;
    set_.bar_any = 0;
    set_.bar_count = 0;
    ++set_.bar_gen;
    set_.bar_cv.notify_all();
  } else {
    set_.bar_cv.wait(lk, [&] { return set_.bar_gen != gen; });
  }
  if (rc) return rc;
  if (set_.bar_result) return csc2_fail(8, "another device of the set failed before the all-reduce; collective skipped");
  return 0;
}

int csc2_allreduce(Ctx &c, double *dev, int n, int op, cudaStream_t s) {
  if (!c.comm || c.comm_size <= 1) return 0;
  // the in-process communicator spans the device set: only the set's workers, all together, may enter it; a
  // user thread working on one selected device (cloudsc2_gpu_select_device) gets that device's own values
  if (set_.inproc_comm && !tl_is_worker) return 0;
  NcclApi *api = nccl_api();
  if (!api) return csc2_fail(7, "NCCL is not available");
  const ncclRedOp_t ops[3] = {ncclMax, ncclMin, ncclSum};
  NK(api->AllReduce(dev, dev, (size_t)n, ncclDouble, ops[op], static_cast<ncclComm_t>(c.comm), s));
  c.launches += 1;
  return 0;
}

// =====================================================================================================
extern "C" {

const char *cloudsc2_gpu_last_error(void) { return tl_err; }

int cloudsc2_gpu_available(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  return n > 0 ? 1 : 0;
}

int cloudsc2_gpu_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

long long cloudsc2_gpu_launch_count(void) {
  long long n = 0;
  for (int i = 0; i < set_.n; ++i) n += set_.ctx[i]->launches;
  return n;
}

int cloudsc2_gpu_init(const cloudsc2_params *params, int klev, const double *ceta, int device) {
  return init_devices(params, klev, ceta, 1, &device);
}

int cloudsc2_gpu_init_multi(const cloudsc2_params *params, int klev, const double *ceta, int ngpus) {
  const int avail = cloudsc2_gpu_device_count();
  if (ngpus <= 0) ngpus = avail;
  if (ngpus > avail) return csc2_fail(5, "%d GPUs requested, %d visible (there is no CPU fallback)", ngpus, avail);
  int devs[CSC2_MAX_DEVICES];
  for (int i = 0; i < ngpus && i < CSC2_MAX_DEVICES; ++i) devs[i] = i;
  return init_devices(params, klev, ceta, ngpus, devs);
}

int cloudsc2_gpu_init_devices(const cloudsc2_params *params, int klev, const double *ceta, int ngpus,
                              const int *devices) {
  if (!devices) return csc2_fail(3, "device list is NULL");
  return init_devices(params, klev, ceta, ngpus, devices);
}

int cloudsc2_gpu_finalize(void) {
  teardown();
  return 0;
}

int cloudsc2_gpu_num_devices(void) { return set_.n; }

int cloudsc2_shard_blocks(int index, int nshards, int nproma, int ngptot, int *block0, int *nblocks,
                          int *ngptot_local, long long *gcol0) {
  if (nshards < 1 || index < 0 || index >= nshards || nproma < 1 || ngptot < 1)
    return csc2_fail(3, "bad arguments to cloudsc2_shard_blocks");
  const Shard s = csc2_shard(index, nshards, nproma, ngptot);
  if (block0) *block0 = s.b0;
  if (nblocks) *nblocks = s.nb;
  if (ngptot_local) *ngptot_local = s.ngptot;
  if (gcol0) *gcol0 = s.gcol0;
  return 0;
}

int cloudsc2_gpu_select_device(int index) {
  if (index < 0) { tl_ctx = nullptr; return 0; }
  if (index >= set_.n) return csc2_fail(3, "device index %d outside the set of %d", index, set_.n);
  tl_ctx = set_.ctx[index].get();
  CK(cudaSetDevice(tl_ctx->device));
  return 0;
}

int cloudsc2_gpu_comm_info(int *rank, int *size, int *nccl_version) {
  Ctx &c = csc2_ctx();
  if (rank) *rank = c.comm ? c.comm_rank : 0;
  if (size) *size = c.comm ? c.comm_size : 1;
  if (nccl_version) {
    *nccl_version = 0;
    if (NcclApi *n = nccl_api()) if (n->GetVersion) n->GetVersion(nccl_version);
  }
  return 0;
}

int cloudsc2_gpu_comm_unique_id(void *id, int bytes) {
  if (!id || bytes < (int)sizeof(ncclUniqueId)) return csc2_fail(3, "id buffer must hold %zu bytes", sizeof(ncclUniqueId));
  NcclApi *n = nccl_api();
  if (!n) return csc2_fail(7, "NCCL is not available");
  ncclUniqueId u;
  NK(n->GetUniqueId(&u));
  std::memcpy(id, &u, sizeof u);
  return 0;
}

int cloudsc2_gpu_comm_init_rank(int rank, int nranks, const void *id, int bytes) {
  if (int rc = csc2_require_init()) return rc;
  if (set_.n != 1) return csc2_fail(3, "cloudsc2_gpu_comm_init_rank needs a single-device context (one process per GPU)");
  if (!id || bytes < (int)sizeof(ncclUniqueId) || rank < 0 || rank >= nranks) return csc2_fail(3, "bad arguments");
  NcclApi *n = nccl_api();
  if (!n) return csc2_fail(7, "NCCL is not available");
  Ctx &c = *set_.ctx[0];
  if (c.comm) { n->CommDestroy(static_cast<ncclComm_t>(c.comm)); c.comm = nullptr; }
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof u);
  ncclComm_t comm;
  NK(n->CommInitRank(&comm, nranks, u, rank));
  c.comm = comm;
  c.comm_rank = rank;
  c.comm_size = nranks;
  return 0;
}

int cloudsc2_gpu_allreduce_dev(double *dev, int n, int op) {
  if (int rc = csc2_require_init()) return rc;
  if (!dev || n <= 0 || op < 0 || op > 2) return csc2_fail(3, "bad arguments to cloudsc2_gpu_allreduce_dev");
  Ctx &c = csc2_ctx();
  if (int rc = csc2_allreduce(c, dev, n, op, c.stream)) return rc;
  CK(cudaStreamSynchronize(c.stream));
  return 0;
}

/* ---- host-pointer entry points: block-sharded over the device set --------------------------------- */

struct NlJob {
  int nproma, klev, ngptot;
  double ptsphy;
  const cloudsc2_fields *h;
  double ek[CSC2_MAX_DEVICES], et[CSC2_MAX_DEVICES];
};
static int nl_job(int idx, void *a) {
  NlJob &j = *static_cast<NlJob *>(a);
  j.ek[idx] = j.et[idx] = 0.0;
  const Shard sh = csc2_shard(idx, std::max(1, set_.n), j.nproma, j.ngptot);
  if (!sh.nb) return 0;
  const cloudsc2_fields f = offset_fields(*j.h, sh.b0, j.nproma, j.klev);
  return csc2_nl_host_one(j.nproma, j.klev, sh.ngptot, j.ptsphy, &f, &j.ek[idx], &j.et[idx]);
}
int cloudsc2_gpu_nl(int nproma, int klev, int ngptot, double ptsphy, const cloudsc2_fields *h,
                    double *elapsed_kernel_s, double *elapsed_total_s) {
  if (int rc = check_host_call(nproma, klev, ngptot, h)) return rc;
  if (set_.n <= 1) return csc2_nl_host_one(nproma, klev, ngptot, ptsphy, h, elapsed_kernel_s, elapsed_total_s);
  NlJob j{nproma, klev, ngptot, ptsphy, h, {}, {}};
  const auto t0 = std::chrono::steady_clock::now();
  if (int rc = csc2_on_all_devices(nl_job, &j)) return rc;
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  // devices run concurrently: the job took as long as the slowest device
  if (elapsed_kernel_s) *elapsed_kernel_s = *std::max_element(j.ek, j.ek + set_.n);
  if (elapsed_total_s) *elapsed_total_s = std::max(wall, *std::max_element(j.et, j.et + set_.n));
  return 0;
}

struct TlAdJob {
  bool is_ad;
  int nproma, klev, ngptot;
  double ptsphy;
  const cloudsc2_fields *h;
  const cloudsc2_incr_in *a;
  const cloudsc2_incr_out *b;
};
static int tlad_job(int idx, void *p) {
  TlAdJob &j = *static_cast<TlAdJob *>(p);
  const Shard sh = csc2_shard(idx, std::max(1, set_.n), j.nproma, j.ngptot);
  if (!sh.nb) return 0;
  const cloudsc2_fields f = offset_fields(*j.h, sh.b0, j.nproma, j.klev);
  const size_t n2 = (size_t)j.nproma * j.klev * sh.b0, n2h = (size_t)j.nproma * (j.klev + 1) * sh.b0;
  cloudsc2_incr_in a = *j.a;
  cloudsc2_incr_out b = *j.b;
  a.paph += n2h; a.pap += n2; a.pq += n2; a.pqs += n2; a.pt += n2; a.pl += n2; a.pi += n2; a.plude += n2;
  a.plu += n2; a.pmfu += n2; a.pmfd += n2; a.gtent += n2; a.gtenq += n2; a.gtenl += n2; a.gteni += n2;
  a.psupsat += n2;
  b.tent += n2; b.tenq += n2; b.tenl += n2; b.teni += n2; b.pclc += n2; b.pcovptot += n2;
  b.pfplsl += n2h; b.pfplsn += n2h; b.pfhpsl += n2h; b.pfhpsn += n2h;
  return csc2_tlad_host_one(j.is_ad, j.nproma, j.klev, sh.ngptot, j.ptsphy, &f, &a, &b);
}
static int tlad(bool is_ad, int nproma, int klev, int ngptot, double ptsphy, const cloudsc2_fields *h,
                const cloudsc2_incr_in *a, const cloudsc2_incr_out *b) {
  if (int rc = check_host_call(nproma, klev, ngptot, h)) return rc;
  if (!a || !b) return csc2_fail(3, "increment struct pointer is NULL");
  if (set_.n <= 1) return csc2_tlad_host_one(is_ad, nproma, klev, ngptot, ptsphy, h, a, b);
  const void *pa[] = {a->paph, a->pap, a->pq, a->pqs, a->pt, a->pl, a->pi, a->plude, a->plu, a->pmfu, a->pmfd,
                      a->gtent, a->gtenq, a->gtenl, a->gteni, a->psupsat, b->tent, b->tenq, b->tenl, b->teni,
                      b->pclc, b->pfplsl, b->pfplsn, b->pfhpsl, b->pfhpsn, b->pcovptot};
  for (const void *q : pa) if (!q) return csc2_fail(3, "a pointer in the increment structs is NULL");
  TlAdJob j{is_ad, nproma, klev, ngptot, ptsphy, h, a, b};
  return csc2_on_all_devices(tlad_job, &j);
}
int cloudsc2_gpu_tl(int nproma, int klev, int ngptot, double ptsphy, const cloudsc2_fields *h,
                    const cloudsc2_incr_in *a, const cloudsc2_incr_out *b) {
  return tlad(false, nproma, klev, ngptot, ptsphy, h, a, b);
}
int cloudsc2_gpu_ad(int nproma, int klev, int ngptot, double ptsphy, const cloudsc2_fields *h,
                    const cloudsc2_incr_in *a, const cloudsc2_incr_out *b) {
  return tlad(true, nproma, klev, ngptot, ptsphy, h, a, b);
}

struct TestJob {
  int nproma, klev, ngptot;
  double ptsphy;
  const cloudsc2_fields *h;
  double *out_blk;                    // ratios_blk [nblocks][10] or norms_col [ngptot][3] or NULL
  double z[CSC2_MAX_DEVICES][10];
};
static int taylor_job(int idx, void *p) {
  TestJob &j = *static_cast<TestJob *>(p);
  const Shard sh = csc2_shard(idx, std::max(1, set_.n), j.nproma, j.ngptot);
  if (!sh.nb) return reduce_identity(csc2_ctx(), 10, 1);
  const cloudsc2_fields f = offset_fields(*j.h, sh.b0, j.nproma, j.klev);
  return csc2_taylor_host_one(j.nproma, j.klev, sh.ngptot, j.ptsphy, &f, j.z[idx],
                              j.out_blk ? j.out_blk + (size_t)10 * sh.b0 : nullptr);
}
int cloudsc2_gpu_tl_taylor(int nproma, int klev, int ngptot, double ptsphy, const cloudsc2_fields *h,
                           double znormg[10], double *ratios_blk) {
  if (int rc = check_host_call(nproma, klev, ngptot, h)) return rc;
  if (!znormg) return csc2_fail(3, "znormg is NULL");
  if (set_.n <= 1) return csc2_taylor_host_one(nproma, klev, ngptot, ptsphy, h, znormg, ratios_blk);
  TestJob j{nproma, klev, ngptot, ptsphy, h, ratios_blk, {}};
  const int rc = csc2_on_all_devices(taylor_job, &j);
  for (int i = 0; i < 10; ++i) znormg[i] = j.z[0][i];     // all-reduced on the devices: identical on every rank
  return rc;
}
static int adtest_job(int idx, void *p) {
  TestJob &j = *static_cast<TestJob *>(p);
  const Shard sh = csc2_shard(idx, std::max(1, set_.n), j.nproma, j.ngptot);
  if (!sh.nb) return reduce_identity(csc2_ctx(), 1, 0);
  const cloudsc2_fields f = offset_fields(*j.h, sh.b0, j.nproma, j.klev);
  return csc2_adtest_host_one(j.nproma, j.klev, sh.ngptot, j.ptsphy, &f, &j.z[idx][0],
                              j.out_blk ? j.out_blk + (size_t)3 * sh.gcol0 : nullptr);
}
int cloudsc2_gpu_ad_test(int nproma, int klev, int ngptot, double ptsphy, const cloudsc2_fields *h,
                         double *znormg, double *norms_col) {
  if (int rc = check_host_call(nproma, klev, ngptot, h)) return rc;
  if (!znormg) return csc2_fail(3, "znormg is NULL");
  if (set_.n <= 1) return csc2_adtest_host_one(nproma, klev, ngptot, ptsphy, h, znormg, norms_col);
  TestJob j{nproma, klev, ngptot, ptsphy, h, norms_col, {}};
  const int rc = csc2_on_all_devices(adtest_job, &j);
  *znormg = j.z[0][0];
  return rc;
}

/* ---- device-resident sharded state (SURVEY 8b "Ownership", 8f-1) ------------------------------------ */

struct LoadJob {
  const cloudsc2_source *src;
  int nproma, ngptot;
};
static int load_job(int idx, void *p) {
  LoadJob &j = *static_cast<LoadJob *>(p);
  Ctx &c = csc2_ctx();
  DevState &st = c.state;
  const cloudsc2_source &s = *j.src;
  // shards: the devices of an in-process set, or -- one process per GPU with a job-wide communicator
  // (cloudsc2_gpu_comm_init_rank) -- the ranks of that communicator; NGPTOT is the global NGPTOTG then
  const bool by_rank = set_.n <= 1 && c.comm && c.comm_size > 1;
  const Shard sh = by_rank ? csc2_shard(c.comm_rank, c.comm_size, j.nproma, j.ngptot)
                           : csc2_shard(idx, std::max(1, set_.n), j.nproma, j.ngptot);
  st.loaded = false;
  st.nproma = j.nproma; st.klev = s.klev; st.ngptot = sh.ngptot; st.nblocks = sh.nb; st.gcol0 = sh.gcol0;
  st.ptsphy = s.ptsphy;
  if (s.klev != c.klev) return csc2_fail(3, "source klev=%d differs from the klev=%d given at init", s.klev, c.klev);
  const size_t nb = std::max(sh.nb, 1);
  const size_t n2 = (size_t)j.nproma * s.klev * nb, n2h = (size_t)j.nproma * (s.klev + 1) * nb;
  const size_t total = 11 * n2 + 5 * n2h + CLOUDSC2_NCLV * n2 + 2 * CLOUDSC2_NSTATE * n2;
  if (int rc = st.mem.reserve(total * sizeof(double))) return rc;
  double *p0 = st.mem.d();
  auto take = [&](size_t n) { double *q = p0; p0 += n; return q; };
  double *pt = take(n2), *pq = take(n2), *pap = take(n2), *paph = take(n2h), *plu = take(n2), *plude = take(n2),
         *pmfu = take(n2), *pmfd = take(n2), *psupsat = take(n2), *pclv = take(CLOUDSC2_NCLV * n2),
         *b_cml = take(CLOUDSC2_NSTATE * n2);
  double *out0 = p0;
  double *b_loc = take(CLOUDSC2_NSTATE * n2), *pa = take(n2), *pcovptot = take(n2), *pfplsl = take(n2h),
         *pfplsn = take(n2h), *pfhpsl = take(n2h), *pfhpsn = take(n2h);
  st.f.pt = pt; st.f.pq = pq; st.f.pap = pap; st.f.paph = paph; st.f.plu = plu; st.f.plude = plude;
  st.f.pmfu = pmfu; st.f.pmfd = pmfd; st.f.psupsat = psupsat; st.f.pclv = pclv; st.f.b_cml = b_cml;
  st.f.b_loc = b_loc; st.f.pa = pa; st.f.pcovptot = pcovptot; st.f.pfplsl = pfplsl; st.f.pfplsn = pfplsn;
  st.f.pfhpsl = pfhpsl; st.f.pfhpsn = pfhpsn;
  if (!sh.nb) { st.loaded = true; return 0; }
  cudaStream_t q = c.stream;
  // FIELD_INIT of the outputs (cloudsc2_array_state_mod.F90:100-127; B_LOC zeroed for determinism)
  CK(cudaMemsetAsync(out0, 0, (size_t)(p0 - out0) * sizeof(double), q));
  // the un-expanded columns go up once (about 4 MB), the expansion happens on the device
  const size_t k2 = (size_t)s.klon * s.klev, k2h = (size_t)s.klon * (s.klev + 1);
  const size_t src_total = 10 * k2 + k2h + CLOUDSC2_NCLV * k2 + CLOUDSC2_NSTATE * k2;
  if (int rc = c.work2.reserve(src_total * sizeof(double))) return rc;
  double *d = c.work2.d();
  struct Item { const double *h; double *dst; int nlev, ndim; };
  const Item items[] = {{s.pt, pt, s.klev, 1}, {s.pq, pq, s.klev, 1}, {s.pap, pap, s.klev, 1},
                        {s.paph, paph, s.klev + 1, 1}, {s.plu, plu, s.klev, 1}, {s.plude, plude, s.klev, 1},
                        {s.pmfu, pmfu, s.klev, 1}, {s.pmfd, pmfd, s.klev, 1}, {s.pa, pa, s.klev, 1},
                        {s.psupsat, psupsat, s.klev, 1}, {s.pclv, pclv, s.klev, CLOUDSC2_NCLV},
                        {s.tend_cml, b_cml, s.klev, CLOUDSC2_NSTATE}};
  for (const Item &it : items) {
    if (!it.h) return csc2_fail(3, "a field of the source columns is NULL");
    const size_t n = (size_t)s.klon * it.nlev * it.ndim;
    CK(cudaMemcpyAsync(d, it.h, n * sizeof(double), cudaMemcpyHostToDevice, q));
    CK(csc2_launch_expand(d, s.klon, (long long)it.nlev * it.ndim, it.dst, j.nproma, sh.ngptot, sh.nb, sh.gcol0, q));
    c.launches += 1;
    d += n;
  }
  CK(cudaStreamSynchronize(q));
  st.loaded = true;
  return 0;
}

int cloudsc2_gpu_state_load(const cloudsc2_source *src, int nproma, int ngptot) {
  if (int rc = csc2_require_init()) return rc;
  if (!src || nproma <= 0 || ngptot <= 0 || src->klon <= 0) return csc2_fail(3, "bad arguments to cloudsc2_gpu_state_load");
  LoadJob j{src, nproma, ngptot};
  return csc2_on_all_devices(load_job, &j);
}

static int free_job(int, void *) {
  DevState &st = csc2_ctx().state;
  st.mem.release();
  st.loaded = false;
  return 0;
}
int cloudsc2_gpu_state_free(void) {
  if (!csc2_ctx_ready()) return 0;
  return csc2_on_all_devices(free_job, nullptr);
}

int cloudsc2_gpu_state_info(int index, int *device, int *nblocks, int *ngptot, long long *gcol0) {
  if (index < 0 || index >= set_.n) return csc2_fail(3, "device index %d outside the set of %d", index, set_.n);
  const Ctx &c = *set_.ctx[index];
  if (!c.state.loaded) return csc2_fail(2, "cloudsc2_gpu_state_load has not been called");
  if (device) *device = c.device;
  if (nblocks) *nblocks = c.state.nblocks;
  if (ngptot) *ngptot = c.state.ngptot;
  if (gcol0) *gcol0 = c.state.gcol0;
  return 0;
}

struct RunJob {
  int mode;                       // 0 NL, 1 Taylor test, 2 adjoint test, 3 TL (dx = 0.01 x), 4 AD of that
  double seconds[CSC2_MAX_DEVICES];
  double z[CSC2_MAX_DEVICES][10];
  int rc_degenerate;
};
static int run_job(int idx, void *p) {
  RunJob &j = *static_cast<RunJob *>(p);
  Ctx &c = csc2_ctx();
  DevState &st = c.state;
  j.seconds[idx] = 0.0;
  if (!st.loaded) return csc2_fail(2, "cloudsc2_gpu_state_load has not been called");
  if (!st.nblocks) {
    if (j.mode == 1) return reduce_identity(c, 10, 1);
    if (j.mode == 2) return reduce_identity(c, 1, 0);
    return 0;
  }
  const auto t0 = std::chrono::steady_clock::now();
  int rc = 0;
  if (j.mode == 0) {
    CK(cudaEventRecord(c.ev[0], c.stream));
    rc = cloudsc2_gpu_nl_dev(st.nproma, st.klev, st.ngptot, st.ptsphy, &st.f, nullptr, nullptr);
    if (rc) return rc;
    CK(cudaEventRecord(c.ev[1], c.stream));
    CK(cudaEventSynchronize(c.ev[1]));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c.ev[0], c.ev[1]));
    j.seconds[idx] = ms * 1e-3;
    return 0;
  }
  if (j.mode == 1) rc = cloudsc2_gpu_tl_taylor_dev(st.nproma, st.klev, st.ngptot, st.ptsphy, &st.f, j.z[idx], nullptr);
  else if (j.mode == 2) rc = cloudsc2_gpu_ad_test_dev(st.nproma, st.klev, st.ngptot, st.ptsphy, &st.f, &j.z[idx][0], nullptr);
  else return csc2_fail(3, "unknown state run mode %d", j.mode);
  // the test entries end with a stream synchronisation: wall time of this device's share
  j.seconds[idx] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return rc;
}
static int state_run(int mode, double *z, int nz, double *elapsed_s, double *per_device_s) {
  if (int rc = csc2_require_init()) return rc;
  RunJob j;
  std::memset(&j, 0, sizeof j);
  j.mode = mode;
  const int rc = csc2_on_all_devices(run_job, &j);
  const int n = std::max(1, set_.n);
  if (elapsed_s) *elapsed_s = *std::max_element(j.seconds, j.seconds + n);
  if (per_device_s) for (int i = 0; i < n; ++i) per_device_s[i] = j.seconds[i];
  // with a communicator the norms are all-reduced on the devices; take the first non-empty shard
  for (int i = 0; i < n; ++i)
    if (set_.ctx[i]->state.nblocks) {
      for (int k = 0; k < nz; ++k) z[k] = j.z[i][k];
      break;
    }
  return rc;
}
int cloudsc2_gpu_state_nl(double *elapsed_s, double *per_device_s) {
  return state_run(0, nullptr, 0, elapsed_s, per_device_s);
}
int cloudsc2_gpu_state_tl_taylor(double znormg[10], double *elapsed_s, double *per_device_s) {
  if (!znormg) return csc2_fail(3, "znormg is NULL");
  return state_run(1, znormg, 10, elapsed_s, per_device_s);
}
int cloudsc2_gpu_state_ad_test(double *znormg, double *elapsed_s, double *per_device_s) {
  if (!znormg) return csc2_fail(3, "znormg is NULL");
  return state_run(2, znormg, 1, elapsed_s, per_device_s);
}

// The ten fields GLOBAL_STATE%VALIDATE compares, in its order (cloudsc2_array_state_mod.F90:239-251).
struct ValJob {
  const cloudsc2_reference *ref;
  double stats[CSC2_MAX_DEVICES][CLOUDSC2_NVALIDATED][5];
};
static int val_job(int idx, void *p) {
  ValJob &j = *static_cast<ValJob *>(p);
  Ctx &c = csc2_ctx();
  DevState &st = c.state;
  cudaStream_t q = c.stream;
  double *d_out = nullptr;
  // everything before the collectives; its status is agreed between the devices first
  auto local_part = [&]() -> int {
    if (!st.loaded) return csc2_fail(2, "cloudsc2_gpu_state_load has not been called");
    const cloudsc2_reference &r = *j.ref;
    const int klon = r.klon, klev = st.klev, nproma = st.nproma;
    const size_t n = (size_t)klon * klev, nh = n + klon;
    // result layout on the device: [10][8] doubles (5 used), reduced in place
    const size_t ref_total = 2 * n + 4 * nh + CLOUDSC2_NSTATE * n;
    if (int rc = c.work2.reserve(ref_total * sizeof(double))) return rc;
    if (int rc = c.res.reserve(csc2_validate_scratch_bytes() + 128 * sizeof(double))) return rc;
    d_out = c.res.d();
    void *scratch = d_out + 128;
    double *d = c.work2.d();
    auto up = [&](const double *h, size_t cnt) -> double * {
      double *dst = d;
      cudaMemcpyAsync(dst, h, cnt * sizeof(double), cudaMemcpyHostToDevice, q);
      d += cnt;
      return dst;
    };
    const double *r_plude = up(r.plude, n), *r_pcov = up(r.pcovptot, n), *r_fl = up(r.pfplsl, nh),
                 *r_fn = up(r.pfplsn, nh), *r_hl = up(r.pfhpsl, nh), *r_hn = up(r.pfhpsn, nh),
                 *r_loc = up(r.tend_loc, CLOUDSC2_NSTATE * n);
    CK(cudaGetLastError());
    const long long bstride = (long long)CLOUDSC2_NSTATE * nproma * klev;
    const size_t slab = (size_t)nproma * klev;
    struct V { const double *ref, *field; int nlev, ndim; long long bs; };
    // TENDENCY_LOC%A/%Q/%T/%CLD = B_LOC(:,:,2,:), (:,:,3,:), (:,:,1,:), (:,:,4:,:)  (:248-251)
    const V v[CLOUDSC2_NVALIDATED] = {
        {r_plude, st.f.plude, klev, 1, 0}, {r_pcov, st.f.pcovptot, klev, 1, 0}, {r_fl, st.f.pfplsl, klev + 1, 1, 0},
        {r_fn, st.f.pfplsn, klev + 1, 1, 0}, {r_hl, st.f.pfhpsl, klev + 1, 1, 0}, {r_hn, st.f.pfhpsn, klev + 1, 1, 0},
        {r_loc + 1 * n, st.f.b_loc + 1 * slab, klev, 1, bstride}, {r_loc + 2 * n, st.f.b_loc + 2 * slab, klev, 1, bstride},
        {r_loc + 0 * n, st.f.b_loc + 0 * slab, klev, 1, bstride},
        {r_loc + 3 * n, st.f.b_loc + 3 * slab, klev, CLOUDSC2_NCLV, bstride}};
    // device layout for the collectives: mins [0..10) | maxs [16..36): vmax, maxerr | sums [48..68)
    double *d_min = d_out, *d_max = d_out + 16, *d_sum = d_out + 48;
    CK(csc2_launch_fill(d_min, 10, 1.7976931348623157e308, q));
    CK(csc2_launch_fill(d_max, 20, -1.7976931348623157e308, q));
    CK(csc2_launch_fill(d_sum, 20, 0.0, q));
    for (int i = 0; i < CLOUDSC2_NVALIDATED && st.nblocks; ++i) {
      const long long bs = v[i].bs ? v[i].bs : (long long)nproma * v[i].nlev * v[i].ndim;
      // results land where the collectives want them (stream order keeps the shared scratch safe)
      CK(csc2_launch_validate_split(v[i].ref, klon, v[i].field, nproma, (long long)v[i].nlev * v[i].ndim, bs, st.ngptot,
                                    st.nblocks, st.gcol0, scratch, d_min + i, d_max + 2 * i, d_sum + 2 * i, q));
      c.launches += 2;
    }
    return 0;
  };
  if (int rc = csc2_agree(local_part())) return rc;
  // CLOUDSC_MPI_REDUCE_MIN / MAX / SUM (validate_mod.F90:197-199) over NVLink, on the device
  if (int rc = csc2_allreduce(c, d_out, 10, 1, q)) return rc;
  if (int rc = csc2_allreduce(c, d_out + 16, 20, 0, q)) return rc;
  if (int rc = csc2_allreduce(c, d_out + 48, 20, 2, q)) return rc;
  double h[80];
  CK(cudaMemcpyAsync(h, d_out, sizeof h, cudaMemcpyDeviceToHost, q));
  CK(cudaStreamSynchronize(q));
  for (int i = 0; i < CLOUDSC2_NVALIDATED; ++i) {
    double *o = j.stats[idx][i];
    o[0] = h[i]; o[1] = h[16 + 2 * i]; o[2] = h[16 + 2 * i + 1]; o[3] = h[48 + 2 * i]; o[4] = h[48 + 2 * i + 1];
  }
  return 0;
}
int cloudsc2_gpu_state_validate(const cloudsc2_reference *ref, double *stats) {
  if (int rc = csc2_require_init()) return rc;
  if (!ref || !stats) return csc2_fail(3, "NULL argument to cloudsc2_gpu_state_validate");
  if (ref->klev != csc2_ctx().klev) return csc2_fail(3, "reference klev differs from the state's");
  std::unique_ptr<ValJob> j(new ValJob());
  j->ref = ref;
  if (int rc = csc2_on_all_devices(val_job, j.get())) return rc;
  std::memcpy(stats, j->stats[csc2_ctx().index], sizeof(double) * CLOUDSC2_NVALIDATED * 5);
  return 0;
}

// Copy one array of the state, all shards in block order, to the host.
struct GetJob {
  int which;
  double *host;
};
static int get_job(int, void *p) {
  GetJob &j = *static_cast<GetJob *>(p);
  Ctx &c = csc2_ctx();
  DevState &st = c.state;
  if (!st.loaded) return csc2_fail(2, "cloudsc2_gpu_state_load has not been called");
  if (!st.nblocks) return 0;
  const size_t n2 = (size_t)st.nproma * st.klev, n2h = (size_t)st.nproma * (st.klev + 1);
  const double *src[18] = {st.f.pt, st.f.pq, st.f.pap, st.f.paph, st.f.plu, st.f.plude, st.f.pmfu, st.f.pmfd,
                           st.f.psupsat, st.f.pclv, st.f.b_cml, st.f.b_loc, st.f.pa, st.f.pcovptot, st.f.pfplsl,
                           st.f.pfplsn, st.f.pfhpsl, st.f.pfhpsn};
  const size_t per_blk[18] = {n2, n2, n2, n2h, n2, n2, n2, n2, n2, CLOUDSC2_NCLV * n2, CLOUDSC2_NSTATE * n2,
                              CLOUDSC2_NSTATE * n2, n2, n2, n2h, n2h, n2h, n2h};
  const size_t b0 = (size_t)(st.gcol0 / st.nproma);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpyAsync(j.host + per_blk[j.which] * b0, src[j.which], per_blk[j.which] * st.nblocks * sizeof(double),
                     cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
  return 0;
}
int cloudsc2_gpu_state_get(const char *name, double *host) {
  static const char *names[18] = {"pt", "pq", "pap", "paph", "plu", "plude", "pmfu", "pmfd", "psupsat", "pclv",
                                  "b_cml", "b_loc", "pa", "pcovptot", "pfplsl", "pfplsn", "pfhpsl", "pfhpsn"};
  if (int rc = csc2_require_init()) return rc;
  if (!name || !host) return csc2_fail(3, "NULL argument to cloudsc2_gpu_state_get");
  GetJob j{-1, host};
  for (int i = 0; i < 18; ++i) if (!std::strcmp(name, names[i])) j.which = i;
  if (j.which < 0) return csc2_fail(3, "unknown state array '%s'", name);
  return csc2_on_all_devices(get_job, &j);
}

// Device pointers of this thread's current device's shard (for callers that drive the _dev entries).
int cloudsc2_gpu_state_fields(cloudsc2_fields *out) {
  if (int rc = csc2_require_init()) return rc;
  if (!out) return csc2_fail(3, "NULL argument");
  if (!csc2_ctx().state.loaded) return csc2_fail(2, "cloudsc2_gpu_state_load has not been called");
  *out = csc2_ctx().state.f;
  return 0;
}

// The dwarf's own NL work flow in one call (dwarf_cloudsc.F90:84-122): LOAD (expand on the device),
// CLOUDSC_DRIVER, VALIDATE.  Only the un-expanded columns cross PCIe.
int cloudsc2_gpu_nl_source(const cloudsc2_source *src, const cloudsc2_reference *ref, int nproma, int ngptot,
                           double *stats, double *elapsed_kernel_s, double *elapsed_total_s) {
  const auto t0 = std::chrono::steady_clock::now();
  if (int rc = cloudsc2_gpu_state_load(src, nproma, ngptot)) return rc;
  if (int rc = cloudsc2_gpu_state_nl(elapsed_kernel_s, nullptr)) return rc;
  if (ref && stats)
    if (int rc = cloudsc2_gpu_state_validate(ref, stats)) return rc;
  if (elapsed_total_s) *elapsed_total_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return 0;
}

}  // extern "C"
