// cloudsc2_nl_experiments_launch.cuh -- NOT PART OF THE PRODUCT LIBRARY (see cloudsc2_nl_experiments.cuh).
// Launchers of the experimental NL variants and the CSC2_NL_VARIANT / "nl_variant" dispatch.
// Included by cloudsc2_nl_kernel.cu at file scope, only with -DCSC2_EXPERIMENTS.
#pragma once

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
  static CSC2_SMEM_FLAGS ok0{0}, ok1{0};
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
  static CSC2_SMEM_FLAGS ok0{0}, ok1{0};
  const bool rv = c.rvtmp2 != 0.0;
  if (cudaError_t e = rv ? csc2_allow_smem(k1, smem, ok1) : csc2_allow_smem(k0, smem, ok0)) return e;
  if (rv) k1<<<grid, 128, smem, s>>>(c, g, in, out);
  else k0<<<grid, 128, smem, s>>>(c, g, in, out);
  return cudaGetLastError();
}

// cudaErrorNotSupported = "run the product kernel"
static cudaError_t csc2_launch_nl_experiment(const KConst &c, const Geom &g, const TrajIn &in, const TrajOut &out,
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
  if (in.pqs) return cudaErrorNotSupported;
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
    case 11: return launch_nl_rv<false, 2, 128, 128, false, 1>(c, g, in, out, s);   // probes
    case 12: return launch_nl_rv<false, 2, 128, 128, false, 2>(c, g, in, out, s);
    case 13: return launch_nl_rv<false, 2, 128, 128, false, 3>(c, g, in, out, s);
    // measured at 163 840 columns: 16 warps 0.849 ms, 12 warps 0.885, 14 warps 0.878, 18 warps 0.916-0.994,
    // 20 warps 0.995 (more warps than 16 cost registers -> instructions, and the kernel is issue-bound)
    default: return cudaErrorNotSupported;                                    // 16: the product shape
  }
}

