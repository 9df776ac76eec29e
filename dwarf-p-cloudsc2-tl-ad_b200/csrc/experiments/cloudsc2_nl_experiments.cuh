// cloudsc2_nl_experiments.cuh -- NOT PART OF THE PRODUCT LIBRARY (see cloudsc2_nl_probe.cuh).
// Measured-and-rejected variants of the nonlinear kernel: TMA bulk copies by a DMA warp,
// warp-private TMA staging, two columns per thread.  Included by cloudsc2_nl_kernel.cu inside its
// anonymous namespace, only with -DCSC2_EXPERIMENTS.
#pragma once


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

