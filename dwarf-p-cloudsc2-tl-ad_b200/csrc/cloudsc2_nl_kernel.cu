// cloudsc2_nl_kernel.cu -- the nonlinear kernel: fused SATUR + CLOUDSC2 for every column of every
// NPROMA block in one launch.  Replaces the OpenMP block loop of CLOUDSC_DRIVER
// (reference src/cloudsc2_nl/cloudsc_driver_mod.F90:82-111).
//
// Mapping: thread <-> column, CTA = 128 consecutive columns, the KLEV loop keeps the column state
// in registers while the next level's 15 inputs are copied by cp.async into a per-thread slot of a
// shared-memory ring (cloudsc2_stage.cuh).  Loads/stores are coalesced along NPROMA (JL); every
// input is read exactly once (plus the ~40-level tropopause pre-pass over PT/PGTENT), PQSAT never
// touches memory.  The same kernel, with flux check-points, is the forward sweep of the adjoint.
#include <cstdint>
#include <cstdlib>

#include "cloudsc2_nl.cuh"
#include "cloudsc2_stage.cuh"
#include "cloudsc2_launch.h"

namespace {

__device__ __forceinline__ double ldin(const double *p) { return __ldg(p); }
__device__ __forceinline__ void stout(double *p, double v) { __stcs(p, v); }

constexpr int NL_NF = CSC2_NTRAJ;   // staged fields per level: 15 inputs + optional PQS

// Template parameters: HAS_PQS (PQS supplied by the caller instead of the fused SATUR), STAGES =
// depth of the shared-memory ring (levels in flight + the one being computed), NT = threads per
// CTA, MAXREG = register cap (-> CTAs per SM), RV = (RVTMP2 != 0).
// CKPT: forward (trajectory) sweep of the adjoint -- additionally check-points the rain / snow flux
// ENTERING every level ([2][klev][ncol_pad]) and writes the trajectory outputs only on request.
struct NLCkpt {
  double *ckpt;
  long long ncol_pad;
  int write_traj;
};

// PROBE (tools/probes, never the default): extra dummy instructions per level to measure what the
// kernel is sensitive to -- 1: 64 integer-ALU ops, 2: 32 independent FP64 FMAs with a constant operand,
// 3: 64 FP32 FMAs.
template <bool HAS_PQS, int STAGES, int NT, int MAXREG, bool RV, bool CKPT, int PROBE = 0>
__global__ void __maxnreg__(MAXREG)
k_cloudsc2_nl(const __grid_constant__ KConst c, const Geom g, const TrajIn in, const TrajOut out,
              const NLCkpt ck) {
  extern __shared__ double ring_all[];
  double *ring = ring_all + threadIdx.x;
  csc2_math_init();
  const int gcol = blockIdx.x * blockDim.x + threadIdx.x;
  const int ibl = gcol / g.nproma;
  if (ibl >= g.nblocks) return;
  const int jl = gcol - ibl * g.nproma;
  const int klev = g.klev, nproma = g.nproma;
  const ColOffsets o = csc2_col_offsets(ibl, jl, nproma, klev, in.bs_cld, in.bs_cml, out.bs_loc);

  if (CKPT && gcol >= g.ngptot) return;
  if (gcol >= g.ngptot) {
    // padding column of the last block: the DRIVER zeroes PCOVPTOT and TENDENCY_LOC%CLD(:,:,NCLV)
    // of whole blocks (driver_mod.F90:87-88); CLOUDSC2 itself never touches columns beyond ICEND.
    // loc_last != NULL marks the driver-level call (NULL: plain CLOUDSC2 semantics, e.g. cloudsc2_).
    if (out.loc_last)
      for (int jk = 0; jk < klev; ++jk) {
        stout(out.pcovptot + o.o1 + (size_t)jk * nproma, 0.0);
        stout(out.loc_last + o.oloc + (size_t)jk * nproma, 0.0);
      }
    return;
  }

  // start the pipeline before the tropopause pre-pass so that its latency is covered
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < klev) csc2_stage_traj<NT, false, HAS_PQS ? 1 : 0>(ring + s * (NL_NF * NT), in, o, s, klev, nproma);
    csc2_cp_async_commit();
  }

  const CritRH crh = make_critrh(tropopause_eta(c, in.pt, in.gt, o.o1, o.ocml, nproma));

  Carry st;
  st.paph0 = ldin(in.paph + o.oh);
  st.rfl = 0.0;
  st.sfl = 0.0;
  const bool wr = !CKPT || ck.write_traj != 0;
  double *ck_r = nullptr, *ck_s = nullptr;
  if (CKPT) {
    ck_r = ck.ckpt + gcol;
    ck_s = ck.ckpt + (size_t)klev * ck.ncol_pad + gcol;
  }
  // flux rows at the model top (cloudsc2.F90:308-309, :732-733 -> -0.0)
  if (wr) {
    stout(out.pfplsl + o.oh, 0.0);
    stout(out.pfplsn + o.oh, 0.0);
    stout(out.pfhpsl + o.oh, -0.0 * c.rlvtt);
    stout(out.pfhpsn + o.oh, -0.0 * c.rlstt);
  }

  int slot = 0, pslot = STAGES - 1;
  unsigned probe_i = threadIdx.x;
  double probe_d[4] = {1.0, 2.0, 3.0, 4.0};
  float probe_f[4] = {1.f, 2.f, 3.f, 4.f};
  for (int jk = 0; jk < klev; ++jk) {
    if (PROBE == 1) {
#pragma unroll
      for (int i = 0; i < 64; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(probe_i) : "r"(jk), "r"(i));
    } else if (PROBE == 2) {
#pragma unroll
      for (int i = 0; i < 32; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(probe_d[i & 3]) : "d"(1.0000001));
    } else if (PROBE == 3) {
#pragma unroll
      for (int i = 0; i < 64; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(probe_f[i & 3]) : "f"(1.0000001f));
    }
    const int pf = jk + STAGES - 1;
    if (pf < klev) csc2_stage_traj<NT, false, HAS_PQS ? 1 : 0>(ring + pslot * (NL_NF * NT), in, o, pf, klev, nproma);
    csc2_cp_async_commit();
    csc2_cp_async_wait<STAGES - 1>();
    const LevIn cur = csc2_read_level<NT>(ring + slot * (NL_NF * NT), jk, klev);
    const double pqs = HAS_PQS ? ring[(size_t)slot * (NL_NF * NT) + 15 * NT]
                               : satur_point(c, cur.pt, csc2_rcp(cur.pap));
    if (CKPT && !wr) {   // with the trajectory outputs written, PFPLSL/PFPLSN are the check-points
      ck_r[(size_t)jk * ck.ncol_pad] = st.rfl;
      ck_s[(size_t)jk * ck.ncol_pad] = st.sfl;
    }
    LevOut y;
    nl_level<RV>(c, crh, jk, cur, pqs, st, y);

    if (wr) {
      const size_t l = (size_t)jk * nproma;
      stout(out.tent + o.oloc + l, y.tent);
      stout(out.tenq + o.oloc + l, y.tenq);
      stout(out.tenl + o.oloc + l, y.tenl);
      stout(out.teni + o.oloc + l, y.teni);
      if (out.loc_last) stout(out.loc_last + o.oloc + l, 0.0);
      stout(out.pclc + o.o1 + l, y.pclc);
      stout(out.pcovptot + o.o1 + l, 0.0);
      stout(out.pfplsl + o.oh + l + nproma, y.rfln);
      stout(out.pfplsn + o.oh + l + nproma, y.sfln);
      stout(out.pfhpsl + o.oh + l + nproma, -y.rfln * c.rlvtt);   // cloudsc2.F90:730-735
      stout(out.pfhpsn + o.oh + l + nproma, -y.sfln * c.rlstt);
    }
    slot = (slot + 1 == STAGES) ? 0 : slot + 1;
    pslot = (pslot + 1 == STAGES) ? 0 : pslot + 1;
  }
  if (PROBE != 0 && (probe_i == 0xdeadbeefu || probe_d[0] + probe_d[1] + probe_d[2] + probe_d[3] == 0.5 ||
                     probe_f[0] + probe_f[1] + probe_f[2] + probe_f[3] == 0.5f))
    stout(out.pcovptot + o.o1, 1.0);   // never true: keeps the dummy chains observable
}

// ---- experimental variant (CSC2_NL_VARIANT=20): level slabs fetched by ONE DMA warp with TMA bulk copies --
// CTA = 16 compute warps (512 consecutive columns, one thread per column as above) + 1 DMA warp.  Per
// level the DMA warp's lanes each issue one `cp.async.bulk` of a contiguous segment (min(NPROMA,512)
// columns of one field) into the shared-memory ring and the bytes arrive on an mbarrier; the compute
// warps wait on that barrier, read their own column's 15 values and release the stage through a second
// mbarrier.  The compute warps then carry no load instructions, no address arithmetic and no array base
// pointers for the inputs (~60 of the 876 warp instructions per level of the cp.async kernel).
// Needs: no padding columns (NGPTOT = NBLOCKS*NPROMA), NPROMA a divisor or a multiple of 512 and >= 32,
// 16-byte aligned arrays, fused SATUR; anything else runs the cp.async kernel.
constexpr int TMA_NF = 15;
struct TmaTable {
  const double *base[TMA_NF];
  long long blk_stride[TMA_NF];   // doubles between consecutive blocks of the field
  int lvl_off[TMA_NF];            // 1 for PAPHP1(JK+1) and PLU(JK+1)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "CSC2_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra CSC2_DONE;\n\t"
      "bra CSC2_WAIT;\n\t"
      "CSC2_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// TMA_CW compute warps per CTA (+ 1 DMA warp), TMA_ST ring stages (levels in flight + the one being read)
template <bool RV, int TMA_CW, int TMA_ST>
__global__ void __launch_bounds__((TMA_CW + 1) * 32, 16 / TMA_CW)
k_cloudsc2_nl_tma(const __grid_constant__ KConst c, const Geom g, const TrajIn in, const TrajOut out,
                  const __grid_constant__ TmaTable tab) {
  constexpr int TMA_COLS = TMA_CW * 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *ring = reinterpret_cast<double *>(smem_raw);                               // [TMA_ST][15][TMA_COLS]
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(ring + TMA_ST * TMA_NF * TMA_COLS);
  __shared__ TmaTable stab;   // a shared copy: the DMA lanes index it with a run-time field number
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + TMA_ST);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int klev = g.klev, nproma = g.nproma;
  const long long ncol = (long long)g.nblocks * nproma;     // == NGPTOT (launch condition)
  const long long col0 = (long long)blockIdx.x * TMA_COLS;
  const int valid = (int)(ncol - col0 < TMA_COLS ? ncol - col0 : TMA_COLS);   // a multiple of 32
  if (threadIdx.x == 0) {
#pragma unroll
    for (int f = 0; f < TMA_NF; ++f) {
      stab.base[f] = tab.base[f]; stab.blk_stride[f] = tab.blk_stride[f]; stab.lvl_off[f] = tab.lvl_off[f];
    }
#pragma unroll
    for (int st = 0; st < TMA_ST; ++st) { mbar_init(full0 + 8 * st, 1); mbar_init(empty0 + 8 * st, valid >> 5); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  csc2_math_init();   // __syncthreads inside: table, barriers and the exp table are visible to everyone

  if (warp == TMA_CW) {
    // ---------------- DMA warp ----------------
    const int seg = nproma < TMA_COLS ? nproma : TMA_COLS;   // contiguous columns of one field
    const int nseg = valid / seg;
    const int ncopy = TMA_NF * nseg;
    int s = 0, round = 0;                     // stage of level jk, number of times the ring has wrapped
    for (int jk = 0; jk < klev; ++jk) {
      if (round > 0) mbar_wait(empty0 + 8 * s, (round - 1) & 1);     // every compute warp has read the stage
      const bool last = jk == klev - 1;                              // PLU(JK+1) does not exist there (:434-438)
      if (lane == 0) mbar_expect_tx(full0 + 8 * s, (uint32_t)((TMA_NF - (last ? 1 : 0)) * valid * 8));
      __syncwarp();
      for (int idx = lane; idx < ncopy; idx += 32) {
        const int f = idx / nseg, sg = idx - f * nseg;
        if (last && f == 7) continue;
        const long long gc = col0 + (long long)sg * seg;
        const long long ibl = gc / nproma;
        const int jl0 = (int)(gc - ibl * nproma);
        const double *src = stab.base[f] + ibl * stab.blk_stride[f] +
                            (long long)(jk + stab.lvl_off[f]) * nproma + jl0;
        tma_load_1d(smem_u32(ring + ((size_t)s * TMA_NF + f) * TMA_COLS + sg * seg), src,
                    (uint32_t)seg * 8u, full0 + 8 * s);
      }
      if (++s == TMA_ST) { s = 0; ++round; }
    }
    return;
  }

  // ---------------- compute warps ----------------
  const int tcol = warp * 32 + lane;
  if (tcol >= valid) return;                 // whole warps only (valid is a multiple of 32)
  const long long gcol = col0 + tcol;
  const int ibl = (int)(gcol / nproma);
  const int jl = (int)(gcol - (long long)ibl * nproma);
  const ColOffsets o = csc2_col_offsets(ibl, jl, nproma, klev, in.bs_cld, in.bs_cml, out.bs_loc);
  const CritRH crh = make_critrh(tropopause_eta(c, in.pt, in.gt, o.o1, o.ocml, nproma));
  Carry st;
  st.paph0 = ldin(in.paph + o.oh);
  st.rfl = 0.0;
  st.sfl = 0.0;
  stout(out.pfplsl + o.oh, 0.0);
  stout(out.pfplsn + o.oh, 0.0);
  stout(out.pfhpsl + o.oh, -0.0 * c.rlvtt);
  stout(out.pfhpsn + o.oh, -0.0 * c.rlstt);
  const double *mine = ring + tcol;
  int s = 0, round = 0;
  for (int jk = 0; jk < klev; ++jk) {
    mbar_wait(full0 + 8 * s, round & 1);
    const LevIn cur = csc2_read_level<TMA_COLS>(mine + (size_t)s * TMA_NF * TMA_COLS, jk, klev);
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * s);
    const double pqs = satur_point(c, cur.pt, csc2_rcp(cur.pap));
    LevOut y;
    nl_level<RV>(c, crh, jk, cur, pqs, st, y);
    const size_t l = (size_t)jk * nproma;
    stout(out.tent + o.oloc + l, y.tent);
    stout(out.tenq + o.oloc + l, y.tenq);
    stout(out.tenl + o.oloc + l, y.tenl);
    stout(out.teni + o.oloc + l, y.teni);
    if (out.loc_last) stout(out.loc_last + o.oloc + l, 0.0);
    stout(out.pclc + o.o1 + l, y.pclc);
    stout(out.pcovptot + o.o1 + l, 0.0);
    stout(out.pfplsl + o.oh + l + nproma, y.rfln);
    stout(out.pfplsn + o.oh + l + nproma, y.sfln);
    stout(out.pfhpsl + o.oh + l + nproma, -y.rfln * c.rlvtt);
    stout(out.pfhpsn + o.oh + l + nproma, -y.sfln * c.rlstt);
    if (++s == TMA_ST) { s = 0; ++round; }
  }
}

// ---- experimental variant (CSC2_NL_VARIANT=25): warp-private TMA staging --------------------------------
// No DMA warp and no CTA-wide coupling: every warp fetches ITS OWN 32 columns.  Lane f < 15 owns field f
// and issues, per level, ONE cp.async.bulk of the warp's 256 contiguous bytes of that field into the
// warp's ring slots; the bytes arrive on a per-warp, per-stage mbarrier.  One warp instruction replaces the
// 15 LDGSTS + 30 address adds + 15 base-pointer loads of the cp.async kernel; the source address of a lane
// advances by NPROMA doubles per level.  Needs NPROMA % 32 == 0 and no padding columns (else cp.async kernel).
template <bool RV>
__global__ void __maxnreg__(128)
k_cloudsc2_nl_wtma(const __grid_constant__ KConst c, const Geom g, const TrajIn in, const TrajOut out,
                   const __grid_constant__ TmaTable tab) {
  constexpr int NT = 128, NW = NT / 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *ring = reinterpret_cast<double *>(smem_raw);                               // [2][15][128]
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(ring + 2 * TMA_NF * NT);   // [NW][2]
  __shared__ TmaTable stab;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int klev = g.klev, nproma = g.nproma;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int f = 0; f < TMA_NF; ++f) {
      stab.base[f] = tab.base[f]; stab.blk_stride[f] = tab.blk_stride[f]; stab.lvl_off[f] = tab.lvl_off[f];
    }
#pragma unroll
    for (int i = 0; i < 2 * NW; ++i) mbar_init(smem_u32(bars + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  csc2_math_init();   // __syncthreads inside
  const long long ncol = (long long)g.nblocks * nproma;          // == NGPTOT (launch condition)
  const long long wcol0 = (long long)blockIdx.x * NT + warp * 32; // first column of this warp
  if (wcol0 >= ncol) return;                                      // whole warps only
  const long long gcol = wcol0 + lane;
  const int ibl = (int)(gcol / nproma);
  const int jl = (int)(gcol - (long long)ibl * nproma);
  const ColOffsets o = csc2_col_offsets(ibl, jl, nproma, klev, in.bs_cld, in.bs_cml, out.bs_loc);

  // this lane's field: source of level 0 (lanes >= 15 own nothing) and destination slots of the warp
  const int f = lane < TMA_NF ? lane : 0;
  const int wbl = (int)(wcol0 / nproma);
  const double *src0 = stab.base[f] + (long long)wbl * stab.blk_stride[f] +
                       (long long)stab.lvl_off[f] * nproma + (wcol0 - (long long)wbl * nproma);
  const uint32_t dst0 = smem_u32(ring + (size_t)f * NT + warp * 32);
  const uint32_t bar0 = smem_u32(bars + 2 * warp);
  auto issue = [&](int lev, int s) {
    const bool last = lev == klev - 1;                            // PLU(JK+1) does not exist at the last level
    if (lane == 0) mbar_expect_tx(bar0 + 8 * s, (uint32_t)((TMA_NF - (last ? 1 : 0)) * 256));
    __syncwarp();
    if (lane < TMA_NF && !(last && lane == 7))
      tma_load_1d(dst0 + (uint32_t)(s * TMA_NF * NT * 8), src0 + (size_t)lev * nproma, 256u, bar0 + 8 * s);
  };
  issue(0, 0);

  const CritRH crh = make_critrh(tropopause_eta(c, in.pt, in.gt, o.o1, o.ocml, nproma));
  Carry st;
  st.paph0 = ldin(in.paph + o.oh);
  st.rfl = 0.0;
  st.sfl = 0.0;
  stout(out.pfplsl + o.oh, 0.0);
  stout(out.pfplsn + o.oh, 0.0);
  stout(out.pfhpsl + o.oh, -0.0 * c.rlvtt);
  stout(out.pfhpsn + o.oh, -0.0 * c.rlstt);
  const double *mine = ring + threadIdx.x;
  for (int jk = 0; jk < klev; ++jk) {
    const int s = jk & 1;
    // the other stage was read during the previous iteration and its values have been consumed
    if (jk + 1 < klev) issue(jk + 1, s ^ 1);
    mbar_wait(bar0 + 8 * s, (jk >> 1) & 1);
    const LevIn cur = csc2_read_level<NT>(mine + (size_t)s * TMA_NF * NT, jk, klev);
    const double pqs = satur_point(c, cur.pt, csc2_rcp(cur.pap));
    LevOut y;
    nl_level<RV>(c, crh, jk, cur, pqs, st, y);
    const size_t l = (size_t)jk * nproma;
    stout(out.tent + o.oloc + l, y.tent);
    stout(out.tenq + o.oloc + l, y.tenq);
    stout(out.tenl + o.oloc + l, y.tenl);
    stout(out.teni + o.oloc + l, y.teni);
    if (out.loc_last) stout(out.loc_last + o.oloc + l, 0.0);
    stout(out.pclc + o.o1 + l, y.pclc);
    stout(out.pcovptot + o.o1 + l, 0.0);
    stout(out.pfplsl + o.oh + l + nproma, y.rfln);
    stout(out.pfplsn + o.oh + l + nproma, y.sfln);
    stout(out.pfhpsl + o.oh + l + nproma, -y.rfln * c.rlvtt);
    stout(out.pfhpsn + o.oh + l + nproma, -y.sfln * c.rlstt);
  }
}

// ---- experimental variant (CSC2_NL_VARIANT=30): two adjacent columns per thread --------------------------
// A thread owns columns 2t and 2t+1 of the CTA: every staged copy, ring read and output store is 16 bytes
// wide, so the per-level overhead that does not depend on the column (array base pointers, address adds,
// constant loads, loop control: ~150 of the 876 warp instructions) is paid once for two columns, and the two
// independent level evaluations sit in one basic block.  255 registers, 8 warps per SM (= 16 column-warps).
// Needs even NPROMA, 16-byte aligned arrays and no padding columns (else the cp.async kernel).
__device__ __forceinline__ void csc2_cp_async16(double2 *smem_dst, const double *gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void stout2(double *p, double a, double b) { __stcs(reinterpret_cast<double2 *>(p), make_double2(a, b)); }

template <bool RV, int MAXREG>
__global__ void __maxnreg__(MAXREG)
k_cloudsc2_nl_x2(const __grid_constant__ KConst c, const Geom g, const TrajIn in, const TrajOut out) {
  constexpr int NT = 128;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *ring = reinterpret_cast<double2 *>(smem_raw) + threadIdx.x;    // [2][15][NT] double2
  csc2_math_init();
  const int klev = g.klev, nproma = g.nproma;
  const long long ncol = (long long)g.nblocks * nproma;                  // == NGPTOT (launch condition)
  const long long gcol = ((long long)blockIdx.x * NT + threadIdx.x) * 2;
  if (gcol >= ncol) return;
  const int ibl = (int)(gcol / nproma);
  const int jl = (int)(gcol - (long long)ibl * nproma);                  // even; jl + 1 is in the same block
  const ColOffsets o = csc2_col_offsets(ibl, jl, nproma, klev, in.bs_cld, in.bs_cml, out.bs_loc);

  auto stage = [&](double2 *d, int jk) {
    const size_t l = (size_t)jk * nproma;
    csc2_cp_async16(d + 0 * NT, in.paph + o.oh + l + nproma);
    csc2_cp_async16(d + 1 * NT, in.pap + o.o1 + l);
    csc2_cp_async16(d + 2 * NT, in.pt + o.o1 + l);
    csc2_cp_async16(d + 3 * NT, in.pq + o.o1 + l);
    csc2_cp_async16(d + 4 * NT, in.pl + o.ocld + l);
    csc2_cp_async16(d + 5 * NT, in.pi + o.ocld + l);
    csc2_cp_async16(d + 6 * NT, in.plude + o.o1 + l);
    if (jk < klev - 1) csc2_cp_async16(d + 7 * NT, in.plu + o.o1 + l + nproma);
    csc2_cp_async16(d + 8 * NT, in.pmfu + o.o1 + l);
    csc2_cp_async16(d + 9 * NT, in.pmfd + o.o1 + l);
    csc2_cp_async16(d + 10 * NT, in.gt + o.ocml + l);
    csc2_cp_async16(d + 11 * NT, in.gq + o.ocml + l);
    csc2_cp_async16(d + 12 * NT, in.gl + o.ocml + l);
    csc2_cp_async16(d + 13 * NT, in.gi + o.ocml + l);
    csc2_cp_async16(d + 14 * NT, in.psupsat + o.o1 + l);
  };
  stage(ring, 0);
  csc2_cp_async_commit();

  const CritRH crh0 = make_critrh(tropopause_eta(c, in.pt, in.gt, o.o1, o.ocml, nproma));
  const CritRH crh1 = make_critrh(tropopause_eta(c, in.pt, in.gt, o.o1 + 1, o.ocml + 1, nproma));
  Carry st0, st1;
  {
    const double2 p0 = *reinterpret_cast<const double2 *>(in.paph + o.oh);
    st0.paph0 = p0.x; st1.paph0 = p0.y;
  }
  st0.rfl = st0.sfl = st1.rfl = st1.sfl = 0.0;
  stout2(out.pfplsl + o.oh, 0.0, 0.0);
  stout2(out.pfplsn + o.oh, 0.0, 0.0);
  stout2(out.pfhpsl + o.oh, -0.0 * c.rlvtt, -0.0 * c.rlvtt);
  stout2(out.pfhpsn + o.oh, -0.0 * c.rlstt, -0.0 * c.rlstt);

  for (int jk = 0; jk < klev; ++jk) {
    const int s = jk & 1;
    if (jk + 1 < klev) stage(ring + (s ^ 1) * (TMA_NF * NT), jk + 1);
    csc2_cp_async_commit();
    csc2_cp_async_wait<1>();
    const double2 *d = ring + s * (TMA_NF * NT);
    LevIn a, b;
    {
      double2 v;
      v = d[0 * NT]; a.paph1 = v.x; b.paph1 = v.y;
      v = d[1 * NT]; a.pap = v.x; b.pap = v.y;
      v = d[2 * NT]; a.pt = v.x; b.pt = v.y;
      v = d[3 * NT]; a.pq = v.x; b.pq = v.y;
      v = d[4 * NT]; a.pl = v.x; b.pl = v.y;
      v = d[5 * NT]; a.pi = v.x; b.pi = v.y;
      v = d[6 * NT]; a.plude = v.x; b.plude = v.y;
      v = (jk < klev - 1) ? d[7 * NT] : make_double2(0.0, 0.0); a.plu1 = v.x; b.plu1 = v.y;
      v = d[8 * NT]; a.pmfu = v.x; b.pmfu = v.y;
      v = d[9 * NT]; a.pmfd = v.x; b.pmfd = v.y;
      v = d[10 * NT]; a.gt = v.x; b.gt = v.y;
      v = d[11 * NT]; a.gq = v.x; b.gq = v.y;
      v = d[12 * NT]; a.gl = v.x; b.gl = v.y;
      v = d[13 * NT]; a.gi = v.x; b.gi = v.y;
      v = d[14 * NT]; a.psupsat = v.x; b.psupsat = v.y;
    }
    const double pqs0 = satur_point(c, a.pt, csc2_rcp(a.pap));
    const double pqs1 = satur_point(c, b.pt, csc2_rcp(b.pap));
    LevOut y0, y1;
    nl_level<RV>(c, crh0, jk, a, pqs0, st0, y0);
    nl_level<RV>(c, crh1, jk, b, pqs1, st1, y1);
    const size_t l = (size_t)jk * nproma;
    stout2(out.tent + o.oloc + l, y0.tent, y1.tent);
    stout2(out.tenq + o.oloc + l, y0.tenq, y1.tenq);
    stout2(out.tenl + o.oloc + l, y0.tenl, y1.tenl);
    stout2(out.teni + o.oloc + l, y0.teni, y1.teni);
    if (out.loc_last) stout2(out.loc_last + o.oloc + l, 0.0, 0.0);
    stout2(out.pclc + o.o1 + l, y0.pclc, y1.pclc);
    stout2(out.pcovptot + o.o1 + l, 0.0, 0.0);
    stout2(out.pfplsl + o.oh + l + nproma, y0.rfln, y1.rfln);
    stout2(out.pfplsn + o.oh + l + nproma, y0.sfln, y1.sfln);
    stout2(out.pfhpsl + o.oh + l + nproma, -y0.rfln * c.rlvtt, -y1.rfln * c.rlvtt);
    stout2(out.pfhpsn + o.oh + l + nproma, -y0.sfln * c.rlstt, -y1.sfln * c.rlstt);
  }
}

// expand_mod.F90:270-302 on the device: dst(nproma, rows, nblocks) <- src(nlon, rows), local
// column j <- source column (gcol0 + j) mod nlon, zero beyond ngptot.
__global__ void k_expand(const double *__restrict__ src, int nlon, long long rows,
                         double *__restrict__ dst, int nproma, int ngptot, long long gcol0,
                         long long total) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; idx < total; idx += stride) {
    const int jl = (int)(idx % nproma);
    const long long t = idx / nproma;
    const long long r = t % rows;
    const long long b = t / rows;
    const long long gcol = b * nproma + jl;
    dst[idx] = (gcol < ngptot) ? __ldg(src + r * nlon + ((gcol0 + gcol) % nlon)) : 0.0;
  }
}

// SATUR on its own (satur.F90:106-123, LDPHYLIN branch): elementwise over n points
__global__ void k_satur(const __grid_constant__ KConst c, const double *__restrict__ pap,
                        const double *__restrict__ pt, double *__restrict__ pqsat, long long n) {
  csc2_math_init();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pqsat[i] = satur_point(c, pt[i], csc2_rcp(pap[i]));
}

// accuracy probe of the branch-free elementary functions (tests/test_gpu_math.py)
__global__ void k_math_probe(int fn, const double *__restrict__ x, double *__restrict__ y, int n) {
  csc2_math_init();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = x[i];
  double r, t;
  switch (fn) {
    case 0: r = csc2_rcp(v); break;
    case 1: r = csc2_exp(v); break;
    case 2: r = csc2_expn(v); break;
    case 3: r = csc2_sqrt(v); break;
    case 4: r = csc2_tanh_p1(v); break;
    default: csc2_tanh_p1_sech2(v, t, r); break;
  }
  y[i] = r;
}

}  // namespace

cudaError_t csc2_launch_satur(const KConst &c, const double *pap, const double *pt, double *pqsat,
                              long long n, cudaStream_t s) {
  k_satur<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(c, pap, pt, pqsat, n);
  return cudaGetLastError();
}

cudaError_t csc2_launch_math_probe(int fn, const double *x, double *y, int n, cudaStream_t s) {
  k_math_probe<<<(n + 127) / 128, 128, 0, s>>>(fn, x, y, n);
  return cudaGetLastError();
}

template <bool HAS_PQS, int STAGES, int NT, int MAXREG, bool RV, bool CKPT = false, int PROBE = 0>
static cudaError_t launch_nl_rv(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
                                cudaStream_t s, NLCkpt ck = NLCkpt{nullptr, 0, 1}) {
  const long long ncol = (long long)g.nblocks * g.nproma;
  const int grid = (int)((ncol + NT - 1) / NT);
  // CSC2_NL_EXTRA_SMEM_KB (probe): unused extra dynamic shared memory per CTA, to measure how the size of the
  // shared-memory carve-out (= what is left for L1) affects the kernel at unchanged occupancy
  static const size_t extra = [] { const char *e = getenv("CSC2_NL_EXTRA_SMEM_KB"); return e ? (size_t)atoi(e) * 1024 : 0; }();
  const size_t smem = (size_t)STAGES * NL_NF * NT * sizeof(double) + extra;
  auto kern = k_cloudsc2_nl<HAS_PQS, STAGES, NT, MAXREG, RV, CKPT, PROBE>;
  static int smem_ok_on_device = -1;
  if (cudaError_t e0 = csc2_allow_smem(kern, smem, smem_ok_on_device)) return e0;
  kern<<<grid, NT, smem, s>>>(c, g, in, out, ck);
  return cudaGetLastError();
}
template <bool HAS_PQS, int STAGES, int NT, int MAXREG>
static cudaError_t launch_nl_variant(const KConst &c, const Geom &g, const TrajIn &in,
                                     const TrajOut &out, cudaStream_t s) {
  // RVTMP2 != 0 (never the case in this dwarf) runs the default shape only
  if (c.rvtmp2 != 0.0) return launch_nl_rv<HAS_PQS, 2, 128, 128, true>(c, g, in, out, s);
  return launch_nl_rv<HAS_PQS, STAGES, NT, MAXREG, false>(c, g, in, out, s);
}

// CSC2_NL_VARIANT (tuning knob, read once): CTA size / CTAs per SM / ring depth
static int g_nl_variant = -1;
static int nl_variant() {
  if (g_nl_variant < 0) {
    const char *e = getenv("CSC2_NL_VARIANT");
    g_nl_variant = e ? atoi(e) : 0;
  }
  return g_nl_variant;
}
void csc2_set_nl_variant(int v) { g_nl_variant = v < 0 ? 0 : v; }

// The TMA variant when the geometry allows it (see k_cloudsc2_nl_tma), else cudaErrorNotSupported.
template <int TMA_CW, int TMA_ST>
static cudaError_t launch_nl_tma(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
                                 cudaStream_t s) {
  constexpr int TMA_COLS = TMA_CW * 32;
  const long long n2 = (long long)g.nproma * g.klev;
  const long long ncol = (long long)g.nblocks * g.nproma;
  const bool shape_ok = g.nproma >= 32 && g.nproma % 2 == 0 &&
                        ((g.nproma <= TMA_COLS && TMA_COLS % g.nproma == 0) || g.nproma % TMA_COLS == 0);
  if (in.pqs || !shape_ok || ncol != g.ngptot) return cudaErrorNotSupported;
  TmaTable t;
  const double *base[TMA_NF] = {in.paph, in.pap, in.pt, in.pq, in.pl, in.pi, in.plude, in.plu, in.pmfu, in.pmfd,
                                in.gt, in.gq, in.gl, in.gi, in.psupsat};
  const long long bs[TMA_NF] = {n2 + g.nproma, n2, n2, n2, in.bs_cld, in.bs_cld, n2, n2, n2, n2,
                                in.bs_cml, in.bs_cml, in.bs_cml, in.bs_cml, n2};
  for (int f = 0; f < TMA_NF; ++f) {
    if (reinterpret_cast<uintptr_t>(base[f]) % 16 != 0) return cudaErrorNotSupported;
    t.base[f] = base[f]; t.blk_stride[f] = bs[f]; t.lvl_off[f] = (f == 0 || f == 7) ? 1 : 0;
  }
  const int grid = (int)((ncol + TMA_COLS - 1) / TMA_COLS);
  const size_t smem = (size_t)TMA_ST * TMA_NF * TMA_COLS * sizeof(double) + 64;
  const bool rv = c.rvtmp2 != 0.0;
  auto k0 = k_cloudsc2_nl_tma<false, TMA_CW, TMA_ST>;
  auto k1 = k_cloudsc2_nl_tma<true, TMA_CW, TMA_ST>;
  static int ok0 = -1, ok1 = -1;
  if (cudaError_t e = rv ? csc2_allow_smem(k1, smem, ok1) : csc2_allow_smem(k0, smem, ok0)) return e;
  if (rv) k1<<<grid, (TMA_CW + 1) * 32, smem, s>>>(c, g, in, out, t);
  else k0<<<grid, (TMA_CW + 1) * 32, smem, s>>>(c, g, in, out, t);
  return cudaGetLastError();
}

static cudaError_t launch_nl_wtma(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
                                  cudaStream_t s) {
  const long long n2 = (long long)g.nproma * g.klev;
  const long long ncol = (long long)g.nblocks * g.nproma;
  if (in.pqs || g.nproma % 32 != 0 || ncol != g.ngptot) return cudaErrorNotSupported;
  TmaTable t;
  const double *base[TMA_NF] = {in.paph, in.pap, in.pt, in.pq, in.pl, in.pi, in.plude, in.plu, in.pmfu, in.pmfd,
                                in.gt, in.gq, in.gl, in.gi, in.psupsat};
  const long long bs[TMA_NF] = {n2 + g.nproma, n2, n2, n2, in.bs_cld, in.bs_cld, n2, n2, n2, n2,
                                in.bs_cml, in.bs_cml, in.bs_cml, in.bs_cml, n2};
  for (int f = 0; f < TMA_NF; ++f) {
    if (reinterpret_cast<uintptr_t>(base[f]) % 16 != 0) return cudaErrorNotSupported;
    t.base[f] = base[f]; t.blk_stride[f] = bs[f]; t.lvl_off[f] = (f == 0 || f == 7) ? 1 : 0;
  }
  const int grid = (int)((ncol + 127) / 128);
  const size_t smem = (size_t)2 * TMA_NF * 128 * sizeof(double) + 64;
  if (c.rvtmp2 != 0.0) k_cloudsc2_nl_wtma<true><<<grid, 128, smem, s>>>(c, g, in, out, t);
  else k_cloudsc2_nl_wtma<false><<<grid, 128, smem, s>>>(c, g, in, out, t);
  return cudaGetLastError();
}

template <int MAXREG>
static cudaError_t launch_nl_x2(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
                                cudaStream_t s) {
  const long long ncol = (long long)g.nblocks * g.nproma;
  if (in.pqs || g.nproma % 2 != 0 || ncol != g.ngptot) return cudaErrorNotSupported;
  const void *ptrs[] = {in.paph, in.pap, in.pt, in.pq, in.pl, in.pi, in.plude, in.plu, in.pmfu, in.pmfd, in.gt,
                        in.gq, in.gl, in.gi, in.psupsat, out.tent, out.tenq, out.tenl, out.teni, out.pclc,
                        out.pcovptot, out.pfplsl, out.pfplsn, out.pfhpsl, out.pfhpsn, out.loc_last};
  for (const void *p : ptrs)
    if (reinterpret_cast<uintptr_t>(p) % 16 != 0) return cudaErrorNotSupported;
  if ((in.bs_cld % 2) || (in.bs_cml % 2) || (out.bs_loc % 2)) return cudaErrorNotSupported;
  const int grid = (int)((ncol / 2 + 127) / 128);
  const size_t smem = (size_t)2 * TMA_NF * 128 * sizeof(double2);
  auto k0 = k_cloudsc2_nl_x2<false, MAXREG>;
  auto k1 = k_cloudsc2_nl_x2<true, MAXREG>;
  static int ok0 = -1, ok1 = -1;
  const bool rv = c.rvtmp2 != 0.0;
  if (cudaError_t e = rv ? csc2_allow_smem(k1, smem, ok1) : csc2_allow_smem(k0, smem, ok0)) return e;
  if (rv) k1<<<grid, 128, smem, s>>>(c, g, in, out);
  else k0<<<grid, 128, smem, s>>>(c, g, in, out);
  return cudaGetLastError();
}

cudaError_t csc2_launch_nl(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
                           cudaStream_t s) {
  if (nl_variant() >= 30 && nl_variant() <= 32) {   // two columns per thread at 255 / 168 / 128 registers
    const cudaError_t e = nl_variant() == 30   ? launch_nl_x2<255>(c, g, in, out, s)
                          : nl_variant() == 31 ? launch_nl_x2<168>(c, g, in, out, s)
                                               : launch_nl_x2<128>(c, g, in, out, s);
    if (e != cudaErrorNotSupported) return e;
  }
  if (nl_variant() == 25) {
    const cudaError_t e = launch_nl_wtma(c, g, in, out, s);
    if (e != cudaErrorNotSupported) return e;
  }
  if (nl_variant() >= 20 && nl_variant() <= 24) {   // experimental TMA variants: (compute warps, stages)
    cudaError_t e;
    switch (nl_variant()) {
      case 20: e = launch_nl_tma<16, 3>(c, g, in, out, s); break;
      case 21: e = launch_nl_tma<8, 3>(c, g, in, out, s); break;
      case 22: e = launch_nl_tma<4, 3>(c, g, in, out, s); break;
      case 23: e = launch_nl_tma<4, 2>(c, g, in, out, s); break;
      default: e = launch_nl_tma<8, 2>(c, g, in, out, s); break;
    }
    if (e != cudaErrorNotSupported) return e;
  }
  if (in.pqs) return launch_nl_variant<true, 2, 128, 128>(c, g, in, out, s);
  switch (nl_variant()) {   //                     stages, threads/CTA, registers -> warps per SM
    case 1: return launch_nl_variant<false, 3, 128, 168>(c, g, in, out, s);   // 12
    case 2: return launch_nl_variant<false, 2, 128, 168>(c, g, in, out, s);   // 12 (the default until r1c)
    case 3: return launch_nl_variant<false, 2, 64, 144>(c, g, in, out, s);    // 14
    case 4: return launch_nl_variant<false, 2, 128, 96>(c, g, in, out, s);    // 20
    case 5: return launch_nl_variant<false, 2, 96, 112>(c, g, in, out, s);    // 18
    case 6: return launch_nl_variant<false, 2, 64, 112>(c, g, in, out, s);    // 18
    case 7: return launch_nl_variant<false, 2, 64, 104>(c, g, in, out, s);    // 18 (19 by regs)
    case 8: return launch_nl_variant<false, 3, 128, 128>(c, g, in, out, s);   // 16, deeper ring
    case 9: return launch_nl_variant<false, 2, 256, 128>(c, g, in, out, s);   // 16, 2 CTAs of 8 warps
    case 10: return launch_nl_variant<false, 2, 64, 128>(c, g, in, out, s);   // 16, 8 CTAs of 2 warps
    case 14: return launch_nl_variant<false, 2, 32, 128>(c, g, in, out, s);   // 16, 16 CTAs of 1 warp
    case 11: return launch_nl_rv<false, 2, 128, 128, false, false, 1>(c, g, in, out, s);   // probes
    case 12: return launch_nl_rv<false, 2, 128, 128, false, false, 2>(c, g, in, out, s);
    case 13: return launch_nl_rv<false, 2, 128, 128, false, false, 3>(c, g, in, out, s);
    // measured at 163 840 columns: 16 warps 0.849 ms, 12 warps 0.885, 14 warps 0.878, 18 warps 0.916-0.994,
    // 20 warps 0.995 (more warps than 16 cost registers -> instructions, and the kernel is issue-bound)
    default: return launch_nl_variant<false, 2, 128, 128>(c, g, in, out, s);  // 16
  }
}

// Forward (trajectory) sweep of the adjoint: the NL kernel + flux check-points.
cudaError_t csc2_launch_nl_ckpt(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
                                double *ckpt, long long ncol_pad, int write_traj, cudaStream_t s) {
  const NLCkpt ck{ckpt, ncol_pad, write_traj};
  TrajOut o = out;
  o.loc_last = nullptr;
  if (c.rvtmp2 != 0.0) {
    if (in.pqs) return launch_nl_rv<true, 2, 128, 128, true, true>(c, g, in, o, s, ck);
    return launch_nl_rv<false, 2, 128, 128, true, true>(c, g, in, o, s, ck);
  }
  if (in.pqs) return launch_nl_rv<true, 2, 128, 128, false, true>(c, g, in, o, s, ck);
  return launch_nl_rv<false, 2, 128, 128, false, true>(c, g, in, o, s, ck);
}

cudaError_t csc2_launch_expand(const double *src, int nlon, long long rows, double *dst, int nproma,
                               int ngptot, int nblocks, long long gcol0, cudaStream_t s) {
  const long long total = (long long)nproma * rows * nblocks;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  k_expand<<<(int)blocks, 256, 0, s>>>(src, nlon, rows, dst, nproma, ngptot, gcol0, total);
  return cudaGetLastError();
}

cudaError_t csc2_upload_levels_nl(const double *ceta, const double *zscalm, const double *sq1mceta,
                                  int klev, cudaStream_t s) {
  return csc2_upload_levels_impl(ceta, zscalm, sq1mceta, klev, s);
}
