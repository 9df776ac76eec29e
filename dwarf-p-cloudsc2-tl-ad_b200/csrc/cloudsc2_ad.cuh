// cloudsc2_ad.cuh -- adjoint of one CLOUDSC2 level for one column: the level's nonlinear
// trajectory is RECOMPUTED from its inputs and the rain/snow flux that entered it (check-pointed
// by the forward sweep), then the adjoint statements of that level are executed, finishing the
// level's 16 input adjoints in the same step.  Follows reference
// src/cloudsc2_ad/cloudsc2ad.F90:449-856 (trajectory), :934-1668 (reverse sweep) and the two
// epilogue loops :1701-1740 (which only touch index (JL,JK) and are therefore folded into the
// level), with CUADJTQSAD (cuadjtqsad.F90:314-367 trajectory, :542-641 adjoint, KCALL==0).
//
// The reference stores 116 (KLON,KLEV) trajectory arrays (:228-292, 127 kB per column); here the
// only state that crosses levels is
//   forward : ZRFL5, ZSFL5 entering each level (2 doubles per level in HBM) and ZTRPAUS,
//   reverse : the adjoints of ZRFL / ZSFL and the pending contribution to PAPHP1(JK).
// Statically dead in this dwarf (LLO2 false, see cloudsc2_nl.cuh): the evaporation block
// (:724-773, :1152-1267) and with it the adjoints of ZEVAPR/ZEVAPS, ZCOVPCLR/ZCOVPTOT (:1407-1420),
// ZCORQS, ZQLIM, ZDTGDP, which stay identically zero.
#pragma once
#include "cloudsc2_nl.cuh"

#define SQA_(x) ((x) * (x))

// Adjoint increments produced by one level (to be accumulated into the caller's arrays).
struct LevAdj {
  double paph_hi, paph_lo;   // contributions to PAPHP1(JK+1) and PAPHP1(JK)
  double pap, pq, pqs, pt, pl, pi, plude, plu1 /*PLU(JK+1)*/, pmfu, pmfd, gt, gq, gl, gi;
  double psupsat;            // ASSIGNED by the reference (:1733), not accumulated
};
// Output adjoints of one level (consumed).
struct LevAdjIn {
  double tent, tenq, tenl, teni, pclc, fl /*PFPLSL(JK+1) incl. folded PFHPSL*/, fn;
};
struct CarryAD {
  double rfl, sfl;           // adjoints of the rain / snow flux entering level JK+1
};

// CUADJTQSAD, KCALL==0, split in two so that the level adjoint can use the adjusted trajectory
// state (needed by the statements that precede the adjustment in reverse order) without running
// the two iterations twice.
struct AdjIter {            // stored trajectory of one iteration (cuadjtqsad.F90:314-367)
  double r, foeew, qs_raw, cor, qs, z2s, q, den;
  bool cap;                 // LLTEST1 / LLTEST2
};
struct AdjTraj {
  AdjIter B, A;             // B = first iteration, A = second
  double z3es, z4es, z5alcp, zaldcp;
};
__device__ __forceinline__ void adj_iter_fwd(const KConst &c, const AdjTraj &tr, double zqp5,
                                             double &t5, double &q5, AdjIter &it) {
  it.r = csc2_rcp(t5 - tr.z4es);
  it.foeew = c.r2es * csc2_exp(tr.z3es * (t5 - c.rtt) * it.r);
  it.qs_raw = zqp5 * it.foeew;
  it.cap = csc2_gt_pos(it.qs_raw, CSC2_ZQMAX);
  it.qs_raw = it.cap ? CSC2_ZQMAX : it.qs_raw;
  it.cor = csc2_rcp(1.0 - c.retv * it.qs_raw);
  it.qs = it.qs_raw * it.cor;
  it.z2s = tr.z5alcp * (it.r * it.r);
  it.q = q5;
  it.den = csc2_rcp(1.0 + it.qs * it.cor * it.z2s);
  const double cond = (q5 - it.qs) * it.den;
  t5 += tr.zaldcp * cond;
  q5 -= cond;
}
// t5/q5: pre-adjustment trajectory (ZTPB5, ZQPB5) -> adjusted trajectory (same arithmetic as
// cuadjtqs_point, so the forward sweep and the recomputation agree bit for bit).
__device__ __forceinline__ void cuadjtqs_traj(const KConst &c, double zqp5 /*1/p5*/, double &t5,
                                              double &q5, AdjTraj &tr) {
  const bool liq = csc2_gt_pos(t5, c.rtt);                              // cuadjtqsad.F90:150-162
  tr.z3es = liq ? c.r3les : c.r3ies;
  tr.z4es = liq ? c.r4les : c.r4ies;
  tr.z5alcp = liq ? c.r5alvcp : c.r5alscp;
  tr.zaldcp = liq ? c.ralvdcp : c.ralsdcp;
  adj_iter_fwd(c, tr, zqp5, t5, q5, tr.B);
  adj_iter_fwd(c, tr, zqp5, t5, q5, tr.A);
}
__device__ __forceinline__ void adj_iter_bwd(const KConst &c, const AdjTraj &tr, const AdjIter &it,
                                             double zqp5, double &zqp_, double &pt_, double &pq_) {
  const double zcond1 = -pq_ + tr.zaldcp * pt_;
  pq_ += zcond1 * it.den;
  const double w = zcond1 * (it.q - it.qs) * (it.den * it.den);
  double zqsat = -zcond1 * it.den - w * it.cor * it.z2s;
  double zcor = -w * it.qs * it.z2s;
  const double z2s = -w * it.qs * it.cor;
  double ztarg = -2.0 * z2s * it.z2s * it.r;                // 2*Z2S*Z5ALCP/(T-Z4ES)**3
  zcor += zqsat * it.qs_raw;
  zqsat = zqsat * it.cor + zcor * c.retv * (it.cor * it.cor);
  zqsat = it.cap ? 0.0 : zqsat;
  const double zfoeew = zqsat * zqp5;
  zqp_ += zqsat * it.foeew;
  // R2ES*EXP(..) of :581/:629 is the stored ZFOEEW5 of the iteration
  ztarg += zfoeew * tr.z3es * (c.rtt - tr.z4es) * it.foeew * (it.r * it.r);
  pt_ += ztarg;
}
// adjoint of both iterations (:546-641); psp_ receives the pressure adjoint (:638)
__device__ __forceinline__ void cuadjtqsad_adj(const KConst &c, const AdjTraj &tr, double zqp5,
                                               double &psp_, double &pt_, double &pq_) {
  double zqp_ = 0.0;
  adj_iter_bwd(c, tr, tr.A, zqp5, zqp_, pt_, pq_);
  adj_iter_bwd(c, tr, tr.B, zqp5, zqp_, pt_, pq_);
  psp_ -= zqp_ * (zqp5 * zqp5);                             // PSP = PSP - ZQP/PSP5**2
}

// One level of CLOUDSC2AD.
//   x5, pqs5      : trajectory inputs of the level (as LevIn, see cloudsc2_nl.cuh)
//   paph0_5       : PAPHP15(JK)
//   rfl5, sfl5    : trajectory rain / snow flux ENTERING the level (forward-sweep checkpoint)
//   ya            : output adjoints of the level
//   ca            : carried flux adjoints (in: of the flux leaving the level, out: entering it)
//   g             : the level's input adjoints
template <bool RV /* RVTMP2 != 0 */, bool LREG /* YRNCL%LREGCL */>
__device__ __forceinline__ void ad_level(const KConst &c, const CritRH &crh, int jk, const LevIn &x5,
                                         double pqs5, double paph0_5, double rfl5, double sfl5,
                                         const LevAdjIn &ya, CarryAD &ca, LevAdj &g) {
  const double dt = c.ptsphy;
  constexpr bool lreg = LREG;
  // ================= trajectory of the level (same arithmetic as nl_level) =================
  const double ztp25 = x5.pt + dt * x5.gt;                 // pre-melt T (ZTP25)
  const double zqp25 = x5.pq + dt * x5.gq + x5.psupsat;    // first-guess q (ZQP25)
  const double zl5 = x5.pl + dt * x5.gl;
  const double zi5 = x5.pi + dt * x5.gi;
  const double zdp5 = x5.paph1 - paph0_5;
  double zzz5 = c.rcpd_inv;
  if (RV) zzz5 = csc2_rcp(c.rcpd + c.rcpd * c.rvtmp2 * zqp25);
  const double zlfdcp5 = c.rlmlt * zzz5, zlsdcp5 = c.rlstt * zzz5, zlvdcp5 = c.rlvtt * zzz5;
  const double pap5_inv = csc2_rcp(x5.pap);

  const double rw = csc2_rcp(ztp25 - c.r4les), ri = csc2_rcp(ztp25 - c.r4ies);
  const bool cold = csc2_lt_pos(ztp25, c.rtt);
  const double targ = 0.17 * (ztp25 - c.rlptrc);
  double tanh_p1, sech2;
  csc2_tanh_p1_sech2(targ, tanh_p1, sech2);
  const double zfwat5 = cold ? 0.545 * tanh_p1 : 1.0;
  const double zfoeew5 = c.r2es * csc2_exp((cold ? c.r3ies * ri : c.r3les * rw) * (ztp25 - c.rtt));
  const double zesdp15 = zfoeew5 * pap5_inv;
  const double zesdp5 = csc2_min_pos(zesdp15, CSC2_ZQMAX);
  const double zfacw5 = c.r5les * (rw * rw), zfaci5 = c.r5ies * (ri * ri);
  const double zfac5 = zfwat5 * zfacw5 + (1.0 - zfwat5) * zfaci5;
  const double zcor5 = csc2_rcp(1.0 - c.retv * zesdp5);
  const double zdqsdtemp5 = zfac5 * zcor5 * pqs5;

  const double zcrh2 = crit_rh(crh, CSC2_CETA(jk), CSC2_SQ1MCETA(jk));
  const bool vcold = csc2_lt_pos(ztp25, c.rtice);
  const double zsupsat5 = vcold ? (1.8 - 3.e-03 * ztp25) : 1.0;
  const double zqsat5 = pqs5 * zsupsat5;
  const double zqcrit5 = zcrh2 * zqsat5;

  const double zscalm = CSC2_ZSCALM(jk);
  const double zqt5 = zqp25 + zl5 + zi5;
  // three-way branch of :543-593 as selects; outside the partial branch the operands of the square
  // root and the reciprocals are replaced by 1 (their results are not used there)
  const bool overcast5 = zqt5 >= zqsat5;
  const bool partial5 = !(zqt5 <= zqcrit5) && !overcast5;
  const double zqpd5 = zqsat5 - zqt5, zqcd5 = zqsat5 - zqcrit5;
  const double den5_inv = csc2_rcp(partial5 ? zqcd5 - zscalm * (zqt5 - zqcrit5) : 1.0);
  const double zsqrt5 = csc2_sqrt((partial5 ? zqpd5 : 1.0) * den5_inv);
  const double zclc5 = partial5 ? 1.0 - zsqrt5 : (overcast5 ? 1.0 : 0.0);
  const double zqc15 = partial5 ? (zscalm * zqpd5 + (1.0 - zscalm) * zqcd5) * (zclc5 * zclc5)
                                : (overcast5 ? (1.0 - zscalm) * zqcd5 : 0.0);

  const double zdp5_inv = csc2_rcp(zdp5);
  const double zgdp5 = c.rg * zdp5_inv;
  const double zlude5 = x5.plude * dt * zgdp5;
  const bool llo1 = (jk < c.klev - 1) && csc2_ge_pos(zlude5, c.rlmin) && csc2_ge_pos(x5.plu1, CSC2_ZEPS2);
  const double plu_inv = csc2_rcp(llo1 ? x5.plu1 : 1.0);
  const double econv = csc2_expn(-zlude5 * plu_inv);
  const double pclc5 = llo1 ? zclc5 + (1.0 - zclc5) * (1.0 - econv) : zclc5;
  const double zqc25 = llo1 ? zqc15 + zlude5 : zqc15;

  const double zfac1 = csc2_rcp(c.rd * ztp25);
  const double zrho5 = x5.pap * zfac1;
  const double zfac2 = csc2_rcp(x5.pap - c.retv * zfoeew5);
  const double zrodqsdp5 = -zrho5 * pqs5 * zfac2;
  const double zldcp5 = zfwat5 * zlvdcp5 + (1.0 - zfwat5) * zlsdcp5;
  const double zfac3 = csc2_rcp(1.0 + zldcp5 * zdqsdtemp5);
  const double dtdzmo5 = c.rg * (c.rcpd_inv - zldcp5 * zrodqsdp5) * zfac3;
  const double zdqsdz5 = zdqsdtemp5 * dtdzmo5 - c.rg * zrodqsdp5;
  const double zfac4 = c.rd * ztp25 * pap5_inv;            // 1/ZRHO5
  const double mf5 = x5.pmfu + x5.pmfd;
  const double zdqc5t = zdqsdz5 * mf5 * dt * zfac4;
  const bool llo3 = zdqc5t < zqc25;
  const double zdqc5 = llo3 ? zdqc5t : zqc25;
  const double zqc35 = zqc25 - zdqc5;

  const double zqlwc15 = zqc35 * zfwat5;
  const double zqiwc15 = zqc35 * (1.0 - zfwat5);
  const double zcondl15 = (zqlwc15 - zl5) * c.zqtmst;
  const double zcondi15 = (zqiwc15 - zi5) * c.zqtmst;

  // melting (:633-651)
  const bool melt = sfl5 != 0.0;
  const bool warm2 = (ztp25 - c.zmeltp2) > 0.0;
  // ZCONS5 = ZCONS2*ZDP5/ZLFDCP5 and its inverse without divisions: 1/ZLFDCP5 = (1/RLMLT)/ZZZ5
  const double lf5_inv = c.rlmlt_inv * (RV ? c.rcpd + c.rcpd * c.rvtmp2 * zqp25 : c.rcpd);
  const double zcons5 = c.zcons2 * zdp5 * lf5_inv;
  const double zcons5_inv = c.zcons2_inv * zdp5_inv * zlfdcp5;
  const double zz2s5 = warm2 ? zcons5 * (ztp25 - c.zmeltp2) : 0.0;
  const bool melt_all = sfl5 <= zz2s5;
  const double zsnmlt5 = melt ? (melt_all ? sfl5 : zz2s5) : 0.0;
  const double ztp15 = ztp25 - zsnmlt5 * zcons5_inv;

  // autoconversion (:655-722)
  const bool cloudy = csc2_gt_pos(pclc5, CSC2_ZEPS2);
  const double pclc5_inv = csc2_rcp(cloudy ? pclc5 : 1.0);
  const double zcldl5 = zqlwc15 * pclc5_inv;
  const double zexp35 = csc2_expn(-SQA_(zcldl5 * c.rlcrit_inv));
  const double zexpdl5 = csc2_exp(-(c.zckcodtl * (1.0 - zexp35)));
  const double zprr5 = cloudy ? zqlwc15 - pclc5 * zcldl5 * zexpdl5 : 0.0;
  const double zcldi5 = zqiwc15 * pclc5_inv;
  const double zexp15 = csc2_exp(0.025 * (ztp15 - c.rtt));
  const double zexp25 = csc2_expn(-SQA_(zcldi5 * c.rlcrit_inv));
  const double zexpdi5 = csc2_exp(-(c.zckcodti * zexp15 * (1.0 - zexp25)));
  const double zprs5 = cloudy ? zqiwc15 - pclc5 * zcldi5 * zexpdi5 : 0.0;
  const double zc2dp5 = c.zcons2 * zdp5;
  const bool frz1 = csc2_lt_pos(ztp15, c.rtt);
  const double zrfreeze15 = frz1 ? zc2dp5 * zprr5 : 0.0;
  const double zfwatr15 = frz1 ? 0.0 : 1.0;

  // first-guess T, q after the tendencies (:777-797)
  const double zldw5 = zldcp5;                             // ZFWAT5*ZLVDCP5+(1-ZFWAT5)*ZLSDCP5
  const double zdqdt5a = -(zcondl15 + zcondi15) + x5.plude * zgdp5;
  const double zdtdt5a = zlvdcp5 * zcondl15 + zlsdcp5 * zcondi15 -
                         (x5.plude * zldw5 - (zlsdcp5 - zlvdcp5) * zrfreeze15) * zgdp5;
  const double ztp35 = ztp15 + dt * zdtdt5a;               // pre-adjust T (ZTPB5)
  const double zqp15 = zqp25 + dt * zdqdt5a;               // pre-adjust q (ZQPB5 = ZQOLD5)
  const double zqold5 = zqp15;

  // ================================ adjoint of the level ====================================
  // carried flux adjoints and this level's output adjoints (:939-957)
  double zsfln = ca.sfl + ya.fn;
  double zrfln = ca.rfl + ya.fl;
  double pclc_ = ya.pclc;
  // final tendencies (:959-1013) -- the trajectory values ZCONDL25.. are needed first, which
  // requires the adjusted state: run the adjustment trajectory + its adjoint below, but the
  // statements that precede it in the reverse order only need ZDQ5 / ZFWATR25 / ZRFREEZE35.
  double t5adj = ztp35, q5adj = zqp15;
  AdjTraj atr;
  cuadjtqs_traj(c, pap5_inv, t5adj, q5adj, atr);           // cloudsc2ad.F90:803-804
  const bool exc = (zqold5 - q5adj) >= 0.0;
  const double zdq5 = exc ? (zqold5 - q5adj) : 0.0;
  const double zdr25 = zc2dp5 * zdq5;
  const bool frz2 = csc2_lt_pos(t5adj, c.rtt);
  const double zfwatr25 = frz2 ? 0.0 : 1.0;
  const double zrfreeze35 = zrfreeze15 + (frz2 ? zfwat5 * zdr25 : 0.0);
  const double zcondl25 = zcondl15 + zfwatr25 * zdq5 * c.zqtmst;
  const double zcondi25 = zcondi15 + (1.0 - zfwatr25) * zdq5 * c.zqtmst;

  double zi_ = -c.zqtmst * ya.teni, zqiwc_ = c.zqtmst * ya.teni;
  double zl_ = -c.zqtmst * ya.tenl, zqlwc_ = c.zqtmst * ya.tenl;
  double zlvdcp_ = 0.0, zlsdcp_ = 0.0, zlfdcp_ = 0.0;
  double zgdp_, zcondl_, zcondi_, plude_, zfwat_, zrfreeze_;
  {
    const double zdtdt = ya.tent, zdqdt = ya.tenq;
    zgdp_ = -zdtdt * (x5.plude * zldw5 - (zlsdcp5 - zlvdcp5) * zrfreeze35);
    zcondl_ = zdtdt * zlvdcp5;
    zcondi_ = zdtdt * zlsdcp5;
    plude_ = -zdtdt * zgdp5 * zldw5;
    zfwat_ = -zdtdt * x5.plude * zgdp5 * (zlvdcp5 - zlsdcp5);
    zrfreeze_ = zdtdt * (zlsdcp5 - zlvdcp5) * zgdp5;
    if (RV) {
      zlvdcp_ += zdtdt * zcondl25 - zdtdt * x5.plude * zgdp5 * zfwat5 - zdtdt * zrfreeze35 * zgdp5;
      zlsdcp_ += zdtdt * zcondi25 - zdtdt * x5.plude * zgdp5 * (1.0 - zfwat5) + zdtdt * zrfreeze35 * zgdp5;
    }
    zgdp_ += zdqdt * x5.plude;
    plude_ += zdqdt * zgdp5;
    zcondl_ -= zdqdt;
    zcondi_ -= zdqdt;
  }

  // excess water to precipitation (:1017-1067)
  double zdp_, zqold_, zqp1_, ztp1_ = 0.0;
  {
    const double zrfreeze2 = zrfreeze_;
    double zdq = (zcondi_ * (1.0 - zfwatr25) + zcondl_ * zfwatr25) * c.zqtmst;
    double zdr2 = (1.0 - zfwatr25) * zsfln + zfwatr25 * zrfln;
    zfwat_ += frz2 ? zdr25 * zrfreeze2 : 0.0;
    zdr2 += frz2 ? zfwat5 * zrfreeze2 : 0.0;
    zdq += zc2dp5 * zdr2;
    zdp_ = c.zcons2 * zdq5 * zdr2;
    if (lreg) zdq *= 0.7;
    zqold_ = exc ? zdq : 0.0;
    zqp1_ = exc ? -zdq : 0.0;
  }

  // saturation adjustment (:1069-1072), on the stored two-iteration trajectory
  double zpp_ = 0.0;
  cuadjtqsad_adj(c, atr, pap5_inv, zpp_, ztp1_, zqp1_);

  // first-guess T and q (:1074-1126)
  double pap_;
  {
    zqp1_ += zqold_;
    pap_ = zpp_;
    const double zdqdt = dt * zqp1_, zdtdt = dt * ztp1_;
    zgdp_ -= zdtdt * (x5.plude * zldw5 - (zlsdcp5 - zlvdcp5) * zrfreeze15);
    zcondl_ += zdtdt * zlvdcp5;
    zcondi_ += zdtdt * zlsdcp5;
    plude_ -= zdtdt * zgdp5 * zldw5;
    zfwat_ -= zdtdt * x5.plude * zgdp5 * (zlvdcp5 - zlsdcp5);
    zrfreeze_ += zdtdt * (zlsdcp5 - zlvdcp5) * zgdp5;
    if (RV) {
      zlvdcp_ += zdtdt * zcondl15 - zdtdt * x5.plude * zgdp5 * zfwat5 - zdtdt * zrfreeze15 * zgdp5;
      zlsdcp_ += zdtdt * zcondi15 - zdtdt * x5.plude * zgdp5 * (1.0 - zfwat5) + zdtdt * zrfreeze15 * zgdp5;
    }
    zgdp_ += zdqdt * x5.plude;
    plude_ += zdqdt * zgdp5;
    zcondl_ -= zdqdt;
    zcondi_ -= zdqdt;
  }

  // new precipitation and autoconversion (:1128-1358)
  {
    const double zdr = (1.0 - zfwatr15) * zsfln + zfwatr15 * zrfln;
    // :1284-1288 (freezing rain source) as selects
    zdp_ += frz1 ? zrfreeze_ * c.zcons2 * zprr5 : 0.0;
    double zprr = frz1 ? zrfreeze_ * zc2dp5 : 0.0;
    double zprs = 0.0;
    zrfreeze_ = frz1 ? 0.0 : zrfreeze_;
    zprr += zc2dp5 * zdr;
    zprs += zc2dp5 * zdr;
    zdp_ += c.zcons2 * (zprr5 + zprs5) * zdr;
    {
      // the adjoint of the autoconversion (:1298-1356) acts only where PCLC5 > ZEPS2; it is
      // evaluated unconditionally (all its trajectory factors are finite) and masked at the end
      const double rl2 = c.rlcrit_inv * c.rlcrit_inv;
      // ice :1298-1327
      const double zprs_c = zprs - zqiwc_;
      const double zinew = -zprs_c;
      const double zdi = -zinew * pclc5 * zcldi5 * zexpdi5;
      const double ki = lreg ? c.zckcodtia : c.zckcodti;
      const double zcldi = zinew * pclc5 * zexpdi5 + (ki * zexp15 * zexp25 * 2.0 * zcldi5 * rl2) * zdi;
      const double zqiwc_c = zqiwc_ + zprs_c + zcldi * pclc5_inv;
      const double ztp1_c = ztp1_ + ki * zexp15 * (1.0 - zexp25) * 0.025 * zdi;
      double pclc_c = pclc_ + zinew * zcldi5 * zexpdi5 - zqiwc15 * zcldi * (pclc5_inv * pclc5_inv);
      // liquid :1332-1356
      const double zprr_c = zprr - zqlwc_;
      const double zlnew = -zprr_c;
      const double zdl = -zlnew * pclc5 * zcldl5 * zexpdl5;
      const double kl = lreg ? c.zckcodtla : c.zckcodtl;
      const double zcldl = zlnew * pclc5 * zexpdl5 + (2.0 * kl * rl2) * zexp35 * zcldl5 * zdl;
      const double zqlwc_c = zqlwc_ + zprr_c + zcldl * pclc5_inv;
      pclc_c += zlnew * zcldl5 * zexpdl5 - zqlwc15 * zcldl * (pclc5_inv * pclc5_inv);
      zqiwc_ = cloudy ? zqiwc_c : zqiwc_;
      zqlwc_ = cloudy ? zqlwc_c : zqlwc_;
      ztp1_ = cloudy ? ztp1_c : ztp1_;
      pclc_ = cloudy ? pclc_c : pclc_;
    }
  }

  // melting of incoming snow (:1362-1400)
  double zsfl_ = zsfln;
  const double zrfl_ = zrfln;
  {
    double zsnmlt = -ztp1_ * zcons5_inv;
    double zcons = ztp1_ * zsnmlt5 * (zcons5_inv * zcons5_inv);
    zsnmlt = zsnmlt - zsfln + zrfln;
    const double zz2s = melt_all ? 0.0 : zsnmlt;
    zcons += warm2 ? (ztp25 - c.zmeltp2) * zz2s : 0.0;
    zsfl_ += (melt && melt_all) ? zsnmlt : 0.0;
    ztp1_ += (melt && warm2) ? zcons5 * zz2s : 0.0;
    zdp_ += melt ? c.zcons2 * zcons * lf5_inv : 0.0;
    if (RV) zlfdcp_ -= melt ? zc2dp5 * zcons * (lf5_inv * lf5_inv) : 0.0;
  }
  ca.sfl = zsfl_;
  ca.rfl = zrfl_;

  // condensate split (:1424-1442)
  double zqc_;
  {
    zqiwc_ += zcondi_ * c.zqtmst;
    zi_ -= zcondi_ * c.zqtmst;
    zqlwc_ += zcondl_ * c.zqtmst;
    zl_ -= zcondl_ * c.zqtmst;
    zqc_ = zqiwc_ * (1.0 - zfwat5) + zqlwc_ * zfwat5;
    zfwat_ += (zqlwc_ - zqiwc_) * zqc35;
  }

  // compensating subsidence (:1446-1496)
  double zfoeew_, zdqsdtemp_, pqs_, pmf_ = 0.0;
  {
    double zdqc = -zqc_;
    zqc_ = llo3 ? zqc_ : 0.0;                              // LLO3 false: ZQC = ZQC + ZDQC = 0
    if (lreg) zdqc *= 0.1;
    const double zdqsdz = llo3 ? zdqc * dt * mf5 * zfac4 : 0.0;
    pmf_ = llo3 ? zdqc * dt * zdqsdz5 * zfac4 : 0.0;
    double zrho = llo3 ? -zdqc * zdqc5 * zfac4 : 0.0;
    const double dtdzmo = zdqsdz * zdqsdtemp5;
    zdqsdtemp_ = zdqsdz * dtdzmo5;
    double zrodqsdp = -zdqsdz * c.rg;
    const double zldcp = -dtdzmo * (c.rg * zrodqsdp5 + dtdzmo5 * zdqsdtemp5) * zfac3;
    zrodqsdp -= dtdzmo * c.rg * zldcp5 * zfac3;
    zdqsdtemp_ -= dtdzmo * dtdzmo5 * zldcp5 * zfac3;
    zfwat_ += zldcp * (zlvdcp5 - zlsdcp5);
    if (RV) {
      zlvdcp_ += zldcp * zfwat5;
      zlsdcp_ += zldcp * (1.0 - zfwat5);
    }
    zrho -= zrodqsdp * pqs5 * zfac2;
    pqs_ = -zrodqsdp * zrho5 * zfac2;
    const double t2 = zrodqsdp * zrho5 * pqs5 * (zfac2 * zfac2);
    pap_ += t2;
    zfoeew_ = -t2 * c.retv;
    pap_ += zrho * zfac1;
    ztp1_ -= zrho * x5.pap * (c.rd * zfac1) * zfac1;       // ZRHO*PAPP15/ZTP25*ZFAC1
  }

  // convective component (:1500-1527)
  double plu1_ = 0.0, paph_g;
  {
    const double zlude = llo1 ? zqc_ + ((1.0 - zclc5) * plu_inv) * econv * pclc_ : 0.0;
    plu1_ = llo1 ? -((1.0 - zclc5) * zlude5 * (plu_inv * plu_inv)) * econv * pclc_ : 0.0;
    pclc_ = llo1 ? pclc_ * (1.0 - (1.0 - econv)) : pclc_;
    plude_ += dt * zgdp5 * zlude;
    zgdp_ += dt * x5.plude * zlude;
    paph_g = c.rg * zgdp_ * (zdp5_inv * zdp5_inv);         // -> -PAPHP1(JK+1), +PAPHP1(JK)
  }

  // uniform total-water distribution (:1531-1583)
  double zqsat_, zqcrit_, zqt_;
  {
    // partial branch (:1545-1583), evaluated with the guarded trajectory values and selected
    double zqpd = zscalm * zqc_ * (zclc5 * zclc5);
    double zqcd = (1.0 - zscalm) * zqc_ * (zclc5 * zclc5);
    double pc_ = pclc_ + (zscalm * zqpd5 + (1.0 - zscalm) * zqcd5) * 2.0 * zclc5 * zqc_;
    if (lreg) {                                            // :1554-1559
      const double zrat = (partial5 ? zqpd5 : 1.0) * csc2_rcp(partial5 ? zqcd5 : 1.0);
      const double b = 1.0 - zscalm * (1.0 - zrat);
      const double zyyy = csc2_min_pos(3.5 * csc2_sqrt(zrat * (b * b * b)) * csc2_rcp(1.0 - zscalm), 0.3);
      pc_ = zyyy * pc_;
    }
    const double h = (0.5 * csc2_rcp(zsqrt5)) * pc_ * den5_inv;
    zqpd -= h;
    const double h2 = h * zqpd5 * den5_inv;
    zqcd += h2;
    const double ov = (1.0 - zscalm) * zqc_;               // overcast branch (:1536-1540)
    zqt_ = partial5 ? -h2 * zscalm - zqpd : 0.0;
    zqcrit_ = partial5 ? h2 * zscalm - zqcd : (overcast5 ? -ov : 0.0);
    zqsat_ = partial5 ? zqcd + zqpd : (overcast5 ? ov : 0.0);
    pclc_ = partial5 ? pc_ : pclc_;
  }
  zqp1_ += zqt_;
  zl_ += zqt_;
  zi_ += zqt_;

  // critical humidity, supersaturation, dqs/dT factor (:1585-1666)
  {
    zqsat_ += zqcrit_ * zcrh2;
    pqs_ += zqsat_ * zsupsat5;
    const double zsupsat = zqsat_ * pqs5;
    ztp1_ -= vcold ? zsupsat * 3.e-03 : 0.0;
    pqs_ += zfac5 * zcor5 * zdqsdtemp_;
    const double zcor = zfac5 * pqs5 * zdqsdtemp_;
    const double zfac = zcor5 * pqs5 * zdqsdtemp_;
    double zesdp = c.retv * zcor * (zcor5 * zcor5);
    const double zfacw = zfwat5 * zfac;
    const double zfaci = (1.0 - zfwat5) * zfac;
    zfwat_ += (zfacw5 - zfaci5) * zfac;
    ztp1_ -= 2.0 * zfaci5 * ri * zfaci;                    // 2*R5IES*ZFACI/(T-R4IES)**3
    ztp1_ -= 2.0 * zfacw5 * rw * zfacw;
    zesdp = csc2_gt_pos(zesdp15, CSC2_ZQMAX) ? 0.0 : zesdp;
    zfoeew_ += zesdp * pap5_inv;
    pap_ -= zesdp * zfoeew5 * (pap5_inv * pap5_inv);
    const double rsel = cold ? ri : rw;
    ztp1_ += (cold ? c.r3ies * (c.rtt - c.r4ies) : c.r3les * (c.rtt - c.r4les)) * zfoeew_ * zfoeew5 * (rsel * rsel);
    ztp1_ += cold ? (0.545 * 0.17) * zfwat_ * sech2 : 0.0;
  }

  // epilogue of the level (:1701-1740)
  if (RV) {
    const double zzz = c.rlvtt * zlvdcp_ + c.rlstt * zlsdcp_ + c.rlmlt * zlfdcp_;
    // the reference evaluates the denominator with ZQP15 as left by the forward sweep, i.e. the
    // post-adjustment humidity (:1712)
    zqp1_ -= zzz * c.rcpd * c.rvtmp2 * SQA_(csc2_rcp(c.rcpd + c.rcpd * c.rvtmp2 * q5adj));
  }
  g.paph_hi = zdp_ - paph_g;
  g.paph_lo = paph_g - zdp_;
  g.pap = pap_;
  g.pqs = pqs_;
  g.plude = plude_;
  g.plu1 = plu1_;
  g.pmfu = pmf_;
  g.pmfd = pmf_;
  g.pi = zi_;     g.gi = dt * zi_;
  g.pl = zl_;     g.gl = dt * zl_;
  g.pq = zqp1_;   g.gq = dt * zqp1_;   g.psupsat = dt * zqp1_;
  g.pt = ztp1_;   g.gt = dt * ztp1_;
}
