// cloudsc2_stage.cuh -- per-thread asynchronous staging of the level slabs in shared memory.
//
// Every CLOUDSC2 kernel streams, per level, 15-42 doubles per column from HBM.  Holding the
// next level in registers (the first version of the NL kernel) costs 2 registers per value and
// caps occupancy; here each thread copies ITS OWN column's values of level jk+D-1 into a
// shared-memory ring with cp.async (LDGSTS, 8 bytes per thread per field, coalesced along
// NPROMA exactly like the direct loads) while level jk is computed.  A thread only ever reads
// the slots it wrote itself, so no CTA barrier is needed: cp.async.wait_group orders the
// thread's own copies.  Ring layout: [stage][field][thread] -> bank-conflict-free LDS.64.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cloudsc2_nl.cuh"

__device__ __forceinline__ void csc2_cp_async8(double *smem_dst, const double *gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void csc2_cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void csc2_cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Offsets (in doubles) of one column inside the blocked arrays: element (jl, jk, ibl) of a field
// with block stride bs lives at ibl*bs + jk*nproma + jl.
struct ColOffsets {
  size_t o1;     // plain (NPROMA,KLEV,NBLOCKS) arrays
  size_t oh;     // half-level (NPROMA,KLEV+1,NBLOCKS) arrays
  size_t ocld;   // PL / PI inside PCLV
  size_t ocml;   // the four slabs of TENDENCY_CML
  size_t oloc;   // the slabs of TENDENCY_LOC
};
__device__ __forceinline__ ColOffsets csc2_col_offsets(int ibl, int jl, int nproma, int klev,
                                                       long long bs_cld, long long bs_cml,
                                                       long long bs_loc) {
  const size_t n2 = (size_t)nproma * klev;
  ColOffsets o;
  o.o1 = (size_t)ibl * n2 + jl;
  o.oh = (size_t)ibl * (n2 + nproma) + jl;
  o.ocld = (size_t)ibl * bs_cld + jl;
  o.ocml = (size_t)ibl * bs_cml + jl;
  o.oloc = (size_t)ibl * bs_loc + jl;
  return o;
}

// ---- the 15 (+1) trajectory inputs of one level ------------------------------------------------
// Ring fields 0..15 of a slot (d points at this thread's element of field 0; fields are NT apart):
//   0 PAPHP1(JK+1) [or PAPHP1(JK) when LOWER]  1 PAPP1  2 PTM1  3 PQM1  4 PL  5 PI  6 PLUDE
//   7 PLU(JK+1)  8 PMFU  9 PMFD  10-13 PGTENT/Q/L/I  14 PSUPSAT  15 PQS (only if given)
constexpr int CSC2_NTRAJ = 16;

// PQS_MODE: 0 = PQS never staged (fused SATUR), 1 = always, 2 = decided at run time by in.pqs
template <int NT, bool LOWER, int PQS_MODE = 2>
__device__ __forceinline__ void csc2_stage_traj(double *d, const TrajIn &in, const ColOffsets &o,
                                                int jk, int klev, int nproma) {
  const size_t l = (size_t)jk * nproma;
  csc2_cp_async8(d + 0 * NT, in.paph + o.oh + l + (LOWER ? 0 : nproma));
  csc2_cp_async8(d + 1 * NT, in.pap + o.o1 + l);
  csc2_cp_async8(d + 2 * NT, in.pt + o.o1 + l);
  csc2_cp_async8(d + 3 * NT, in.pq + o.o1 + l);
  csc2_cp_async8(d + 4 * NT, in.pl + o.ocld + l);
  csc2_cp_async8(d + 5 * NT, in.pi + o.ocld + l);
  csc2_cp_async8(d + 6 * NT, in.plude + o.o1 + l);
  if (jk < klev - 1) csc2_cp_async8(d + 7 * NT, in.plu + o.o1 + l + nproma);   // PLU(JK+1), :434-438
  csc2_cp_async8(d + 8 * NT, in.pmfu + o.o1 + l);
  csc2_cp_async8(d + 9 * NT, in.pmfd + o.o1 + l);
  csc2_cp_async8(d + 10 * NT, in.gt + o.ocml + l);
  csc2_cp_async8(d + 11 * NT, in.gq + o.ocml + l);
  csc2_cp_async8(d + 12 * NT, in.gl + o.ocml + l);
  csc2_cp_async8(d + 13 * NT, in.gi + o.ocml + l);
  csc2_cp_async8(d + 14 * NT, in.psupsat + o.o1 + l);
  if (PQS_MODE == 1 || (PQS_MODE == 2 && in.pqs)) csc2_cp_async8(d + 15 * NT, in.pqs + o.o1 + l);
}
// 15 consecutive ring fields starting at d as a LevIn (field 0 -> paph1, whichever half level it is)
template <int NT>
__device__ __forceinline__ LevIn csc2_read_level(const double *d, int jk, int klev) {
  LevIn x;
  x.paph1 = d[0 * NT]; x.pap = d[1 * NT]; x.pt = d[2 * NT]; x.pq = d[3 * NT]; x.pl = d[4 * NT];
  x.pi = d[5 * NT]; x.plude = d[6 * NT];
  x.plu1 = (jk < klev - 1) ? d[7 * NT] : 0.0;
  x.pmfu = d[8 * NT]; x.pmfd = d[9 * NT]; x.gt = d[10 * NT]; x.gq = d[11 * NT]; x.gl = d[12 * NT];
  x.gi = d[13 * NT]; x.psupsat = d[14 * NT];
  return x;
}

// Opt a kernel in to more than 48 kB of dynamic shared memory, once per (kernel, device): the
// attribute is per device, and one process may drive several devices from several threads.
// `flags` is a static of the calling launcher (one per kernel instantiation): bit d = done on device d.
#include <atomic>
typedef std::atomic<unsigned long long> CSC2_SMEM_FLAGS;
template <typename K>
static inline cudaError_t csc2_allow_smem(K kern, size_t smem, CSC2_SMEM_FLAGS &flags) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (flags.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) flags.fetch_or(bit, std::memory_order_release);
  return e;
}
