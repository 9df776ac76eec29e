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
