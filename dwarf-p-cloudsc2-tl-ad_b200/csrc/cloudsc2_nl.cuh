// cloudsc2_nl.cuh -- one level of the nonlinear CLOUDSC2 column physics for one column, all
// state in registers.  Follows reference src/cloudsc2_nl/cloudsc2.F90:339-725 (LPHYLIN branch,
// LEVAPLS2 = LDRAIN1D = .FALSE.) with the saturation adjustment CUADJTQS (cuadjtqs.F90:212-244,
// inlined in the reference at cloudsc2.F90:630-669).
//
// What is not here, and why:
//  * the precipitation-evaporation block (:556-591): LLO2 is statically false in this dwarf, so
//    ZEVAPR = ZEVAPS = 0, ZCOVPCLR/ZCOVPTOT/ZCORQS/ZQLIM feed nothing and PCOVPTOT stays 0;
//  * the (KLON,KLEV) work arrays: every one of them is level-local and collapses to a scalar.
// Divisions that share a denominator are done as one reciprocal + multiplies; the statement
// order of the reference (pre-melt T for ZFWAT/ZRHO, post-melt T for ice autoconversion and the
// rain/snow split, phase of the adjustment fixed by the pre-adjustment T) is kept.
#pragma once
#include "cloudsc2_common.cuh"

// Inputs of one level of one column.
struct LevIn {
  double paph1;   // PAPHP1(JK+1)
  double pap, pt, pq, pl, pi, plude, plu1 /*PLU(JK+1)*/, pmfu, pmfd, gt, gq, gl, gi, psupsat;
};
// Outputs of one level.
struct LevOut {
  double tent, tenq, tenl, teni, pclc, rfln, sfln;
};
// State carried down the column.
struct Carry {
  double paph0;   // PAPHP1(JK)
  double rfl, sfl;
};

// Two-iteration saturation adjustment (cuadjtqs.F90:118-130 phase select, :212-244).
__device__ __forceinline__ void cuadjtqs_point(const KConst &c, double zqp /*1/p*/, double &t,
                                               double &q) {
  const bool liq = csc2_gt_pos(t, c.rtt);
  const double z3es = liq ? c.r3les : c.r3ies;
  const double z4es = liq ? c.r4les : c.r4ies;
  const double z5alcp = liq ? c.r5alvcp : c.r5alscp;
  const double zaldcp = liq ? c.ralvdcp : c.ralsdcp;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const double r = csc2_rcp(t - z4es);
    const double foeew = c.r2es * csc2_exp(z3es * (t - c.rtt) * r);
    const double q0 = csc2_min_pos(zqp * foeew, CSC2_ZQMAX);      // ZQSAT before the correction
    // ZCOR = 1/u, u = 1-RETV*q0 ; ZQSAT = q0/u ; ZCOND = (q-ZQSAT)/(1+ZQSAT*ZCOR*Z2S)  (:226-235)
    // with ONE reciprocal:  ZCOND = (q*u - q0) * u / (u*u + q0*Z2S)
    const double u = fma(-c.retv, q0, 1.0);
    const double z2s = z5alcp * (r * r);
    const double cond = (fma(q, u, -q0) * u) * csc2_rcp(fma(q0, z2s, u * u));
    t += zaldcp * cond;
    q -= cond;
  }
}

// One level in two phases:
//   nl_local : everything that depends only on the level's own inputs (first guess, dqs/dT factor,
//              cloud cover, convective detrainment, subsidence, condensation rates, liquid
//              autoconversion, the temperature-independent part of the ice autoconversion) -- ~60 %
//              of the FP64 work, with plenty of independent exp / reciprocal chains;
//   nl_tail  : the part chained to the rain/snow flux arriving from the level above (melting ->
//              post-melt T -> ice autoconversion -> precipitation -> saturation adjustment -> fluxes)
//              -- one long dependent chain.
// (Software-pipelining nl_local(JK+1) with nl_tail(JK) in one loop iteration was tried to give
// each thread two instruction streams: 1.01 ms against 0.925 ms for the plain order at 163 840
// columns -- ptxas does not interleave them and the carried LevLocal costs registers -- so the
// kernels run the phases back to back.)  `pqs` is the saturation humidity of the level (SATUR
// output or caller-supplied).  Straight-line code: every data-dependent IF of the reference is a
// select on values that are computed unconditionally with guarded operands.
struct LevLocal {
  double ztp1, zqp1, zl, zi;          // first guess (pre-melt T)
  double zc2dp, zgdp, pap_inv;        // ZCONS2*ZDP, RG/ZDP, 1/PAPP1
  double zcons, zcons_inv;            // ZCONS2*ZDP/ZLFDCP and its inverse (melting)
  double zlsdcp, zlvdcp, zfwat, zldcpw;
  double pclc, zqlwc, zqiwc, zprr;    // cover, liquid after / ice before autoconversion, rain source
  double zcondl, zcondi, plude;
  double ice_fac, picld;              // ZCKCODTI*(1-EXP(-(ZCLDI/ZLCRIT)**2)), PCLC*ZCLDI (:525-529)
  double paph1;
};

// RV = (RVTMP2 != 0), a compile-time switch so that the level is ONE basic block (a run-time test,
// even a warp-uniform one, splits it and stops the scheduler from interleaving across the split;
// RVTMP2 is 0 in this dwarf: it is never loaded, yoethf.F90:30 vs :79-99).
template <bool RV>
__device__ __forceinline__ LevLocal nl_local(const KConst &c, const CritRH &crh, int jk,
                                             const LevIn &x, double pqs, double paph0) {
  LevLocal L;
  const double dt = c.ptsphy;
  // first guess (cloudsc2.F90:253-260)
  const double ztp1 = x.pt + dt * x.gt;
  const double zqp1 = x.pq + dt * x.gq + x.psupsat;
  const double zl = x.pl + dt * x.gl;
  const double zi = x.pi + dt * x.gi;
  // :268-278
  const double zdp = x.paph1 - paph0;
  double zzz = c.rcpd_inv;
  if (RV) zzz = csc2_rcp(c.rcpd + c.rcpd * c.rvtmp2 * zqp1);
  const double zlfdcp = c.rlmlt * zzz, zlsdcp = c.rlstt * zzz, zlvdcp = c.rlvtt * zzz;
  const double pap_inv = csc2_rcp(x.pap);
  const double zdp_inv = csc2_rcp(zdp);

  // dqs/dT correction factor (:349-375)
  const bool cold = csc2_lt_pos(ztp1, c.rtt);
  const double rw = csc2_rcp(ztp1 - c.r4les), ri = csc2_rcp(ztp1 - c.r4ies);
  const double tanh_p1 = csc2_pin(csc2_tanh_p1(0.17 * (ztp1 - c.rlptrc)));   // unconditional
  const double zfwat = cold ? 0.545 * tanh_p1 : 1.0;
  const double zfoeew = c.r2es * csc2_exp((cold ? c.r3ies * ri : c.r3les * rw) * (ztp1 - c.rtt));
  const double zfacw = c.r5les * (rw * rw), zfaci = c.r5ies * (ri * ri);
  const double zfac = fma(zfwat, zfacw - zfaci, zfaci);     // ZFWAT*ZFACW+(1-ZFWAT)*ZFACI
  // ZCOR = 1/(1-RETV*ZESDP) with ZESDP = MIN(ZFOEEW/PAPP1, ZQMAX) shares the reciprocal of the
  // subsidence section: 1/(1-RETV*ZFOEEW/PAPP1) = PAPP1 * ZFAC2, ZFAC2 = 1/(PAPP1-RETV*ZFOEEW) (:451)
  const double zfac2 = csc2_rcp(x.pap - c.retv * zfoeew);
  const double zcor = csc2_gt_pos(zfoeew * pap_inv, CSC2_ZQMAX) ? c.zcor_cap : x.pap * zfac2;
  const double zdqsdtemp = zfac * zcor * pqs;

  // critical humidity, ice supersaturation (:384-408)
  const double zcrh2 = crit_rh(crh, CSC2_CETA(jk), CSC2_SQ1MCETA(jk));
  const double zsupsat = csc2_lt_pos(ztp1, c.rtice) ? (1.8 - 3.e-03 * ztp1) : 1.0;
  const double zqsat = pqs * zsupsat;
  const double zqcrit = zcrh2 * zqsat;

  // uniform total-water distribution (:412-427)
  const double zscalm = CSC2_ZSCALM(jk);
  const double zqt = zqp1 + zl + zi;
  double pclc, zqc;
  {
    const bool clear = zqt <= zqcrit, overcast = zqt >= zqsat;
    const bool partial = !clear && !overcast;
    const double zqpd = zqsat - zqt, zqcd = zqsat - zqcrit;
    // in the partial branch zqpd > 0 and the denominator is > (1-zscalm)*zqcd > 0
    const double den = partial ? (zqcd - zscalm * (zqt - zqcrit)) : 1.0;
    const double root = csc2_sqrt((partial ? zqpd : 0.0) * csc2_rcp(den));
    const double pc = 1.0 - root;
    pclc = partial ? pc : (overcast ? 1.0 : 0.0);
    zqc = partial ? (zscalm * zqpd + (1.0 - zscalm) * zqcd) * (pc * pc)
                  : (overcast ? (1.0 - zscalm) * zqcd : 0.0);
  }

  // convective detrainment (:431-444)
  const double zgdp = c.rg * zdp_inv;
  const double zlude = x.plude * dt * zgdp;
  {
    const bool llo1 = jk < c.klev - 1 && csc2_ge_pos(zlude, c.rlmin) && csc2_ge_pos(x.plu1, CSC2_ZEPS2);
    const double e = csc2_pin(csc2_expn(-zlude * csc2_rcp(llo1 ? x.plu1 : 1.0)));   // unconditional
    pclc = llo1 ? fma(pclc - 1.0, e, 1.0) : pclc;              // PCLC+(1-PCLC)*(1-EXP(..))
    zqc = llo1 ? zqc + zlude : zqc;
  }

  // compensating subsidence (:448-460)
  const double zldcp = fma(zfwat, zlvdcp - zlsdcp, zlsdcp);   // ZFWAT*ZLVDCP+(1-ZFWAT)*ZLSDCP
  {
    const double zfac1 = csc2_rcp(c.rd * ztp1);
    const double zrho = x.pap * zfac1;
    const double zrodqsdp = -zrho * pqs * zfac2;
    const double zfac3 = csc2_rcp(1.0 + zldcp * zdqsdtemp);
    const double dtdzmo = c.rg * (c.rcpd_inv - zldcp * zrodqsdp) * zfac3;
    const double zdqsdz = zdqsdtemp * dtdzmo - c.rg * zrodqsdp;
    const double zfac4 = c.rd * ztp1 * pap_inv;                  // 1/ZRHO
    const double zdqc = dmin_(zdqsdz * (x.pmfu + x.pmfd) * dt * zfac4, zqc);
    zqc = zqc - zdqc;
  }

  // new condensate and condensation rates (:464-469)
  double zqlwc = zqc * zfwat;
  const double zqiwc = zqc - zqlwc;                            // ZQC*(1-ZFWAT)
  L.zcondl = (zqlwc - zl) * c.zqtmst;
  L.zcondi = (zqiwc - zi) * c.zqtmst;

  // melting of incoming snow (:487-498): the level-local factors
  {
    // ZCONS = ZCONS2*ZDP/ZLFDCP and its inverse without a division: 1/ZLFDCP = (1/RLMLT)/ZZZ
    const double zzz_inv = RV ? c.rcpd + c.rcpd * c.rvtmp2 * zqp1 : c.rcpd;
    L.zcons = c.zcons2 * zdp * (c.rlmlt_inv * zzz_inv);
    L.zcons_inv = c.zcons2_inv * zdp_inv * zlfdcp;
  }

  // autoconversion (:504-534): liquid completely, ice up to the factor that needs the post-melt T
  {
    const bool cloudy = csc2_gt_pos(pclc, CSC2_ZEPS2);
    const double pclc_inv = csc2_rcp(cloudy ? pclc : 1.0);
    const double zcldl = zqlwc * pclc_inv;
    const double rl = zcldl * c.rlcrit_inv;
    const double zdl = c.zckcodtl * (1.0 - csc2_expn(-(rl * rl)));
    const double zlnew = pclc * zcldl * csc2_exp(-zdl);
    const double zcldi = zqiwc * pclc_inv;
    const double rr = zcldi * c.rlcrit_inv;
    L.zprr = cloudy ? zqlwc - zlnew : 0.0;
    zqlwc = zqlwc - L.zprr;
    // not cloudy: ice_fac = 0 and picld = ZQIWC make ZINEW = ZQIWC, i.e. ZPRS = 0 exactly
    const double zexp2 = csc2_pin(csc2_expn(-(rr * rr)));                        // unconditional
    L.ice_fac = cloudy ? c.zckcodti * (1.0 - zexp2) : 0.0;
    L.picld = cloudy ? pclc * zcldi : zqiwc;
  }
  L.ztp1 = ztp1; L.zqp1 = zqp1; L.zl = zl; L.zi = zi;
  L.zc2dp = c.zcons2 * zdp; L.zgdp = zgdp; L.pap_inv = pap_inv;
  L.zlsdcp = zlsdcp; L.zlvdcp = zlvdcp; L.zfwat = zfwat; L.zldcpw = zldcp;   // as written at :609-610
  L.pclc = pclc; L.zqlwc = zqlwc; L.zqiwc = zqiwc; L.plude = x.plude;
  L.paph1 = x.paph1;
  return L;
}

__device__ __forceinline__ void nl_tail(const KConst &c, const LevLocal &L, Carry &st, LevOut &y) {
  const double dt = c.ptsphy;
  double ztp1 = L.ztp1, zqp1 = L.zqp1;
  // melting of incoming snow (:487-498); with ZSFL == 0 the statements reduce to the identity
  const double zsnmlt = dmin_(st.sfl, L.zcons * dmax_(0.0, ztp1 - c.zmeltp2));
  double zrfln = st.rfl + zsnmlt;
  double zsfln = st.sfl - zsnmlt;
  ztp1 = ztp1 - zsnmlt * L.zcons_inv;

  // ice autoconversion on the post-melt T (:522-533)
  const double zdi = L.ice_fac * csc2_exp(0.025 * (ztp1 - c.rtt));
  const double zinew = L.picld * csc2_exp(-zdi);
  const double zprs = L.zqiwc - zinew;
  const double zqiwc = L.zqiwc - zprs;

  // new precipitation, rain/snow split on the post-melt T (:538-552)
  const double zdr = L.zc2dp * (L.zprr + zprs);
  const bool frz1 = csc2_lt_pos(ztp1, c.rtt);
  double zrfreeze = frz1 ? L.zc2dp * L.zprr : 0.0;
  zsfln += frz1 ? zdr : 0.0;
  zrfln += frz1 ? 0.0 : zdr;

  // first-guess T and q after the tendencies (:601-618)
  double zcondl = L.zcondl, zcondi = L.zcondi;
  const double plude_gdp = L.plude * L.zgdp;
  {
    const double zdqdt = -(zcondl + zcondi) + plude_gdp;
    const double zdtdt = L.zlvdcp * zcondl + L.zlsdcp * zcondi -
                         (L.plude * L.zldcpw - (L.zlsdcp - L.zlvdcp) * zrfreeze) * L.zgdp;
    ztp1 = ztp1 + dt * zdtdt;
    zqp1 = zqp1 + dt * zdqdt;
  }
  const double zqold = zqp1;

  // saturation adjustment (:622-670)
  cuadjtqs_point(c, L.pap_inv, ztp1, zqp1);

  // excess water to precipitation (:672-692)
  {
    const double zdq = dmax_(0.0, zqold - zqp1);
    const double zdr2 = L.zc2dp * zdq;
    const bool frz2 = csc2_lt_pos(ztp1, c.rtt);
    zrfreeze += frz2 ? L.zfwat * zdr2 : 0.0;
    zcondi += frz2 ? zdq * c.zqtmst : 0.0;
    zcondl += frz2 ? 0.0 : zdq * c.zqtmst;
    zsfln += frz2 ? zdr2 : 0.0;
    zrfln += frz2 ? 0.0 : zdr2;
  }

  // final tendencies and fluxes (:694-716)
  y.tenq = -(zcondl + zcondi) + plude_gdp;
  y.tent = L.zlvdcp * zcondl + L.zlsdcp * zcondi -
           (L.plude * L.zldcpw - (L.zlsdcp - L.zlvdcp) * zrfreeze) * L.zgdp;
  y.tenl = (L.zqlwc - L.zl) * c.zqtmst;
  y.teni = (zqiwc - L.zi) * c.zqtmst;
  y.pclc = L.pclc;
  y.rfln = zrfln;
  y.sfln = zsfln;
  // carry (:720-723)
  st.rfl = zrfln;
  st.sfl = zsfln;
  st.paph0 = L.paph1;
}

// Both phases back to back (AD forward sweep, Taylor-test kernel).
template <bool RV>
__device__ __forceinline__ void nl_level(const KConst &c, const CritRH &crh, int jk,
                                         const LevIn &x, double pqs, Carry &st, LevOut &y) {
  const LevLocal L = nl_local<RV>(c, crh, jk, x, pqs, st.paph0);
  nl_tail(c, L, st, y);
}
