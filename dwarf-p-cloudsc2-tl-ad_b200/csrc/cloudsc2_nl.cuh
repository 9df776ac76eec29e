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
  const bool liq = t > c.rtt;
  const double z3es = liq ? c.r3les : c.r3ies;
  const double z4es = liq ? c.r4les : c.r4ies;
  const double z5alcp = liq ? c.r5alvcp : c.r5alscp;
  const double zaldcp = liq ? c.ralvdcp : c.ralsdcp;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const double r = csc2_rcp(t - z4es);
    const double foeew = c.r2es * csc2_exp(z3es * (t - c.rtt) * r);
    double qsat = dmin_(zqp * foeew, CSC2_ZQMAX);
    const double cor = csc2_rcp(1.0 - c.retv * qsat);
    qsat *= cor;
    const double z2s = z5alcp * (r * r);
    const double den = csc2_rcp(1.0 + qsat * cor * z2s);
    const double cond = (q - qsat) * den;
    t += zaldcp * cond;
    q -= cond;
  }
}

// One level.  `pqs` is the saturation humidity of the level (SATUR output or caller-supplied).
// Straight-line code: every data-dependent IF of the reference is a select on values that are
// computed unconditionally with guarded operands, so that the compiler can overlap the
// independent exp / reciprocal chains of a level (the kernel is FP64-issue bound, not HBM bound).
__device__ __forceinline__ void nl_level(const KConst &c, const CritRH &crh, int jk,
                                         const LevIn &x, double pqs, Carry &st, LevOut &y) {
  const double dt = c.ptsphy;
  // first guess (cloudsc2.F90:253-260)
  double ztp1 = x.pt + dt * x.gt;
  double zqp1 = x.pq + dt * x.gq + x.psupsat;
  const double zl = x.pl + dt * x.gl;
  const double zi = x.pi + dt * x.gi;
  // :268-278
  const double zdp = x.paph1 - st.paph0;
  double zzz = c.rcpd_inv;
  if (c.rvtmp2 != 0.0) zzz = csc2_rcp(c.rcpd + c.rcpd * c.rvtmp2 * zqp1);
  const double zlfdcp = c.rlmlt * zzz, zlsdcp = c.rlstt * zzz, zlvdcp = c.rlvtt * zzz;
  const double pap_inv = csc2_rcp(x.pap);
  const double zdp_inv = csc2_rcp(zdp);

  // dqs/dT correction factor (:349-375)
  const bool cold = ztp1 < c.rtt;
  const double rw = csc2_rcp(ztp1 - c.r4les), ri = csc2_rcp(ztp1 - c.r4ies);
  const double zfwat = cold ? 0.545 * csc2_tanh_p1(0.17 * (ztp1 - c.rlptrc)) : 1.0;
  const double zfoeew = c.r2es * csc2_exp((cold ? c.r3ies * ri : c.r3les * rw) * (ztp1 - c.rtt));
  const double zfacw = c.r5les * (rw * rw), zfaci = c.r5ies * (ri * ri);
  const double zfac = zfwat * zfacw + (1.0 - zfwat) * zfaci;
  // ZCOR = 1/(1-RETV*ZESDP) with ZESDP = MIN(ZFOEEW/PAPP1, ZQMAX) shares the reciprocal of the
  // subsidence section: 1/(1-RETV*ZFOEEW/PAPP1) = PAPP1 * ZFAC2, ZFAC2 = 1/(PAPP1-RETV*ZFOEEW) (:451)
  const double zfac2 = csc2_rcp(x.pap - c.retv * zfoeew);
  const double zcor = (zfoeew * pap_inv > CSC2_ZQMAX) ? c.zcor_cap : x.pap * zfac2;
  const double zdqsdtemp = zfac * zcor * pqs;

  // critical humidity, ice supersaturation (:384-408)
  const double zcrh2 = crit_rh(crh, CSC2_CETA(jk), CSC2_SQ1MCETA(jk));
  const double zsupsat = (ztp1 < c.rtice) ? (1.8 - 3.e-03 * ztp1) : 1.0;
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
    const bool llo1 = jk < c.klev - 1 && zlude >= c.rlmin && x.plu1 >= CSC2_ZEPS2;
    const double e = csc2_expn(-zlude * csc2_rcp(llo1 ? x.plu1 : 1.0));
    pclc = llo1 ? pclc + (1.0 - pclc) * (1.0 - e) : pclc;
    zqc = llo1 ? zqc + zlude : zqc;
  }

  // compensating subsidence (:448-460)
  {
    const double zfac1 = csc2_rcp(c.rd * ztp1);
    const double zrho = x.pap * zfac1;
    const double zrodqsdp = -zrho * pqs * zfac2;
    const double zldcp = zfwat * zlvdcp + (1.0 - zfwat) * zlsdcp;
    const double zfac3 = csc2_rcp(1.0 + zldcp * zdqsdtemp);
    const double dtdzmo = c.rg * (c.rcpd_inv - zldcp * zrodqsdp) * zfac3;
    const double zdqsdz = zdqsdtemp * dtdzmo - c.rg * zrodqsdp;
    const double zfac4 = c.rd * ztp1 * pap_inv;                  // 1/ZRHO
    const double zdqc = dmin_(zdqsdz * (x.pmfu + x.pmfd) * dt * zfac4, zqc);
    zqc = zqc - zdqc;
  }

  // new condensate and condensation rates (:464-469)
  double zqlwc = zqc * zfwat;
  double zqiwc = zqc * (1.0 - zfwat);
  double zcondl = (zqlwc - zl) * c.zqtmst;
  double zcondi = (zqiwc - zi) * c.zqtmst;

  // melting of incoming snow (:487-498); with ZSFL == 0 the statements reduce to the identity
  double zrfln, zsfln;
  {
    // ZCONS = ZCONS2*ZDP/ZLFDCP and its inverse without a division: 1/ZLFDCP = (1/RLMLT)/ZZZ
    const double zzz_inv = (c.rvtmp2 != 0.0) ? c.rcpd + c.rcpd * c.rvtmp2 * zqp1 : c.rcpd;
    const double lf_inv = c.rlmlt_inv * zzz_inv;
    const double zcons = c.zcons2 * zdp * lf_inv;
    const double zcons_inv = c.zcons2_inv * zdp_inv * zlfdcp;
    const double zsnmlt = dmin_(st.sfl, zcons * dmax_(0.0, ztp1 - c.zmeltp2));
    zrfln = st.rfl + zsnmlt;
    zsfln = st.sfl - zsnmlt;
    ztp1 = ztp1 - zsnmlt * zcons_inv;
  }

  // autoconversion liquid / ice (:504-534)
  double zprr, zprs;
  {
    const bool cloudy = pclc > CSC2_ZEPS2;
    const double pclc_inv = csc2_rcp(cloudy ? pclc : 1.0);
    const double zcldl = zqlwc * pclc_inv;
    const double rl = zcldl * c.rlcrit_inv;
    const double zdl = c.zckcodtl * (1.0 - csc2_expn(-(rl * rl)));
    const double zlnew = pclc * zcldl * csc2_exp(-zdl);
    const double zcldi = zqiwc * pclc_inv;
    const double rr = zcldi * c.rlcrit_inv;
    const double zdi = c.zckcodti * csc2_exp(0.025 * (ztp1 - c.rtt)) * (1.0 - csc2_expn(-(rr * rr)));
    const double zinew = pclc * zcldi * csc2_exp(-zdi);
    zprr = cloudy ? zqlwc - zlnew : 0.0;
    zprs = cloudy ? zqiwc - zinew : 0.0;
    zqlwc = zqlwc - zprr;
    zqiwc = zqiwc - zprs;
  }

  // new precipitation, rain/snow split on the post-melt T (:538-552)
  const double zc2dp = c.zcons2 * zdp;
  const double zdr = zc2dp * (zprr + zprs);
  const bool frz1 = ztp1 < c.rtt;
  double zrfreeze = frz1 ? zc2dp * zprr : 0.0;
  zsfln += frz1 ? zdr : 0.0;
  zrfln += frz1 ? 0.0 : zdr;

  // first-guess T and q after the tendencies (:601-618)
  const double zldcpw = zfwat * zlvdcp + (1.0 - zfwat) * zlsdcp;   // as written at :609-610
  {
    const double zdqdt = -(zcondl + zcondi) + x.plude * zgdp;
    const double zdtdt = zlvdcp * zcondl + zlsdcp * zcondi -
                         (x.plude * zldcpw - (zlsdcp - zlvdcp) * zrfreeze) * zgdp;
    ztp1 = ztp1 + dt * zdtdt;
    zqp1 = zqp1 + dt * zdqdt;
  }
  const double zqold = zqp1;

  // saturation adjustment (:622-670)
  cuadjtqs_point(c, pap_inv, ztp1, zqp1);

  // excess water to precipitation (:672-692)
  {
    const double zdq = dmax_(0.0, zqold - zqp1);
    const double zdr2 = zc2dp * zdq;
    const bool frz2 = ztp1 < c.rtt;
    zrfreeze += frz2 ? zfwat * zdr2 : 0.0;
    zcondi += frz2 ? zdq * c.zqtmst : 0.0;
    zcondl += frz2 ? 0.0 : zdq * c.zqtmst;
    zsfln += frz2 ? zdr2 : 0.0;
    zrfln += frz2 ? 0.0 : zdr2;
  }

  // final tendencies and fluxes (:694-716)
  y.tenq = -(zcondl + zcondi) + x.plude * zgdp;
  y.tent = zlvdcp * zcondl + zlsdcp * zcondi -
           (x.plude * zldcpw - (zlsdcp - zlvdcp) * zrfreeze) * zgdp;
  y.tenl = (zqlwc - zl) * c.zqtmst;
  y.teni = (zqiwc - zi) * c.zqtmst;
  y.pclc = pclc;
  y.rfln = zrfln;
  y.sfln = zsfln;
  // carry (:720-723)
  st.rfl = zrfln;
  st.sfl = zsfln;
  st.paph0 = x.paph1;
}
