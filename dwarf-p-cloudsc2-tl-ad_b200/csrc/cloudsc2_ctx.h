// cloudsc2_ctx.h -- internal: the per-device context of the library and the device set.
//
// One process drives 1..N GPUs (SURVEY 8b "Threading", 8e): the library owns one Ctx per device
// (streams, staging buffers, scratch, launch counter, NCCL communicator handle) and one persistent
// host worker thread per device.  Every entry point works on "the calling thread's current
// context": the primary context (device 0 of the set) for user threads, the worker's own context
// for worker threads, or whatever cloudsc2_gpu_select_device chose.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "cloudsc2_launch.h"

constexpr int kStreams = 3;
constexpr int CSC2_MAX_DEVICES = 16;

int csc2_fail(int code, const char *fmt, ...);     // sets the calling thread's error text
const char *csc2_error_text();
void csc2_set_error_text(const char *s);

#define CK(call)                                                                             \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess)                                                                   \
      return csc2_fail(100 + (int)e_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                       __FILE__, __LINE__);                                                  \
  } while (0)

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);        // cudaFree synchronises the device: nothing still reads the old buffer
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return csc2_fail(100 + (int)e, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    cap = bytes;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  double *d() const { return static_cast<double *>(p); }
};

// Device-resident CLOUDSC2_ARRAY_STATE of one shard (cloudsc2_array_state_mod.F90:28-60),
// built by cloudsc2_gpu_state_load from the un-expanded source columns.
struct DevState {
  bool loaded = false;
  int nproma = 0, klev = 0, ngptot = 0, nblocks = 0;   // this shard
  long long gcol0 = 0;                                 // global index of the shard's first column
  double ptsphy = 0.0;
  DevBuf mem;                                          // one allocation for all 18 arrays
  cloudsc2_fields f{};
};

struct Ctx {
  bool init = false;
  int device = 0;          // CUDA ordinal
  int index = 0;           // position in the device set
  cloudsc2_params prm;
  int klev = 0;
  double ceta[CSC2_KLEV_MAX];
  double zscalm[CSC2_KLEV_MAX];
  int kwin0 = 0, kwin1 = -1;
  cudaStream_t stream = nullptr;
  cudaStream_t pipe[kStreams] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t scratch_done = nullptr;   // last use of work/res by a kernel (possibly on a caller-supplied stream)
  bool scratch_used = false;
  long long launches = 0;
  DevBuf in, out, work, work2, res;     // staging for the host-pointer entry points + scratch
  DevState state;
  // NCCL communicator of this device (ncclComm_t; void* keeps nccl.h out of the other files):
  // rank comm_rank of comm_size, either of the in-process device set (ncclCommInitAll) or of a
  // multi-process job (cloudsc2_gpu_comm_init_rank).  nullptr: no reduction across devices.
  void *comm = nullptr;
  int comm_rank = 0, comm_size = 1;
};

Ctx &csc2_ctx();                  // the calling thread's current context (never null once initialised)
bool csc2_ctx_ready();
int csc2_require_init();          // 0, or error 2 with text; also makes the context's device current
int csc2_num_devices();           // size of the device set (0 before init)
Ctx &csc2_ctx_at(int index);

// Run fn(index) on the worker thread of every device of the set (its context current, its CUDA
// device selected) and wait.  Returns the first non-zero return code; the error text of that worker
// becomes the caller's.  With one device fn runs inline on the calling thread.
int csc2_on_all_devices(int (*fn)(int index, void *arg), void *arg);

// Block range of device `index` when `nblocks` NPROMA blocks are spread over `ndev` devices: the
// arithmetic the reference uses for MPI ranks on columns (cloudsc2_nl/dwarf_cloudsc.F90:65-69),
// applied to blocks: per = (nblocks-1)/ndev + 1, device r owns [r*per, min(nblocks, (r+1)*per)).
struct Shard {
  int b0, nb;           // first block, number of blocks (0: nothing to do)
  int ngptot;           // valid columns of the shard
  long long gcol0;      // = b0 * nproma
};
Shard csc2_shard(int index, int ndev, int nproma, int ngptot);

// Called by every device's worker exactly once per multi-device job, BEFORE the job's first collective, with
// the status of its work so far: returns rc if it is non-zero, 8 if another device of the set failed, else 0 --
// and then nobody enters the collective (an all-reduce entered by only some ranks never returns).  Outside a
// worker thread (single device, process-per-GPU jobs) it returns rc unchanged.
int csc2_agree(int rc);

// MAX / MIN / SUM all-reduce of n device-resident doubles over the context's communicator, in
// place, enqueued on `s` (no-op without a communicator).  op: 0 max, 1 min, 2 sum.
int csc2_allreduce(Ctx &c, double *dev, int n, int op, cudaStream_t s);
