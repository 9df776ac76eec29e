// cloudsc2_tl.cuh -- one level of the tangent-linear CLOUDSC2TL for one column, trajectory ("5")
// and perturbation advanced in lockstep, all in registers.  Follows reference
// src/cloudsc2_tl/cloudsc2tl.F90:455-1103 and CUADJTQSTL (cuadjtqstl.F90:333-405, KCALL==0).
// Every MIN/MAX/IF switches on the trajectory, the perturbation follows the taken branch.
// The LLO2 evaporation block (:845-943) is statically dead (see cloudsc2_nl.cuh).
#pragma once
#include "cloudsc2_nl.cuh"

struct CarryTL {
  double paph0, rfl, sfl;     // perturbations of the carried state
};

#define SQ_(x) ((x) * (x))

// CUADJTQSTL, KCALL==0: two sweeps, phase fixed by the incoming trajectory T.
__device__ __forceinline__ void cuadjtqstl_point(const KConst &c, double psp5_inv, double psp,
                                                 double &t5, double &q5, double &t, double &q) {
  const bool liq = t5 > c.rtt;
  const double z3es = liq ? c.r3les : c.r3ies;
  const double z4es = liq ? c.r4les : c.r4ies;
  const double z5alcp = liq ? c.r5alvcp : c.r5alscp;
  const double zaldcp = liq ? c.ralvdcp : c.ralsdcp;
  const double zqp = -psp * (psp5_inv * psp5_inv);
  const double zqp5 = psp5_inv;
  const double k3 = z3es * (c.rtt - z4es);
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const double r = csc2_rcp(t5 - z4es);
    const double r2 = r * r;
    const double foeew5 = c.r2es * csc2_exp(z3es * (t5 - c.rtt) * r);
    const double foeew = k3 * t * foeew5 * r2;
    double qsat = zqp5 * foeew + zqp * foeew5;
    double qsat5 = zqp5 * foeew5;
    const bool cap = qsat5 > CSC2_ZQMAX;
    qsat = cap ? 0.0 : qsat;
    qsat5 = cap ? CSC2_ZQMAX : qsat5;
    const double cor5 = csc2_rcp(1.0 - c.retv * qsat5);
    const double cor = (c.retv * qsat) * (cor5 * cor5);
    qsat = qsat5 * cor + qsat * cor5;
    qsat5 = qsat5 * cor5;
    const double z2s5 = z5alcp * r2;
    const double z2s = -2.0 * t * z2s5 * r;
    const double den = csc2_rcp(1.0 + qsat5 * cor5 * z2s5);
    const double cond5 = (q5 - qsat5) * den;
    const double cond = (q - qsat) * den -
                        cond5 * (qsat * cor5 * z2s5 + qsat5 * cor * z2s5 + qsat5 * cor5 * z2s) * den;
    t += zaldcp * cond;   t5 += zaldcp * cond5;
    q -= cond;            q5 -= cond5;
  }
}

// x5/pqs5 : trajectory inputs ; dx/dpqs : perturbations.  y5/dy : outputs.
// Straight-line like nl_level: the reference's IFs (which all test trajectory values) are selects.
// RV = (RVTMP2 != 0), LREG = YRNCL%LREGCL: compile-time so that the level is one basic block.
template <bool RV, bool LREG>
__device__ __forceinline__ void tl_level(const KConst &c, const CritRH &crh, int jk,
                                         const LevIn &x5, double pqs5, const LevIn &dx, double dpqs,
                                         Carry &st5, CarryTL &st, LevOut &y5, LevOut &dy) {
  const double dt = c.ptsphy;
  constexpr bool lreg = LREG;
  // first guess (cloudsc2tl.F90:342-353)
  double ztp1 = dx.pt + dt * dx.gt;
  double ztp15 = x5.pt + dt * x5.gt;
  double zqp1 = dx.pq + dt * dx.gq + dx.psupsat;
  double zqp15 = x5.pq + dt * x5.gq + x5.psupsat;
  const double zl = dx.pl + dt * dx.gl, zl5 = x5.pl + dt * x5.gl;
  const double zi = dx.pi + dt * dx.gi, zi5 = x5.pi + dt * x5.gi;
  // :362-376
  const double zdp = dx.paph1 - st.paph0;
  const double zdp5 = x5.paph1 - st5.paph0;
  double zzz5 = c.rcpd_inv, zzz = 0.0;
  if (RV) {
    zzz5 = csc2_rcp(c.rcpd + c.rcpd * c.rvtmp2 * zqp15);
    zzz = -c.rcpd * c.rvtmp2 * zqp1 * (zzz5 * zzz5);
  }
  const double zlfdcp = c.rlmlt * zzz, zlfdcp5 = c.rlmlt * zzz5;
  const double zlsdcp = c.rlstt * zzz, zlsdcp5 = c.rlstt * zzz5;
  const double zlvdcp = c.rlvtt * zzz, zlvdcp5 = c.rlvtt * zzz5;
  const double pap5_inv = csc2_rcp(x5.pap);

  // dqs/dT correction factor (:463-491)
  const bool cold = ztp15 < c.rtt;
  const double rw = csc2_rcp(ztp15 - c.r4les), ri = csc2_rcp(ztp15 - c.r4ies);
  double tanh_p1, sech2;
  csc2_tanh_p1_sech2(0.17 * (ztp15 - c.rlptrc), tanh_p1, sech2);
  const double zfwat = cold ? (0.545 * 0.17) * ztp1 * sech2 : 0.0;
  const double zfwat5 = cold ? 0.545 * tanh_p1 : 1.0;
  const double rsel = cold ? ri : rw;
  const double z3sel = cold ? c.r3ies : c.r3les;
  const double z4sel = cold ? c.r4ies : c.r4les;
  const double zfoeew5 = c.r2es * csc2_exp(z3sel * (ztp15 - c.rtt) * rsel);
  const double zfoeew = z3sel * (c.rtt - z4sel) * ztp1 * zfoeew5 * (rsel * rsel);
  double zesdp = zfoeew * pap5_inv - dx.pap * zfoeew5 * (pap5_inv * pap5_inv);
  double zesdp5 = zfoeew5 * pap5_inv;
  {
    const bool cap = zesdp5 > CSC2_ZQMAX;
    zesdp = cap ? 0.0 : zesdp;
    zesdp5 = cap ? CSC2_ZQMAX : zesdp5;
  }
  const double zfacw5 = c.r5les * (rw * rw), zfaci5 = c.r5ies * (ri * ri);
  const double zfacw = -2.0 * ztp1 * zfacw5 * rw, zfaci = -2.0 * ztp1 * zfaci5 * ri;
  const double zfac = zfwat5 * zfacw + zfacw5 * zfwat + (1.0 - zfwat5) * zfaci - zfaci5 * zfwat;
  const double zfac5 = zfwat5 * zfacw5 + (1.0 - zfwat5) * zfaci5;
  const double zcor5 = csc2_rcp(1.0 - c.retv * zesdp5);
  const double zcor = c.retv * zesdp * (zcor5 * zcor5);
  const double zdqsdtemp = zfac5 * zcor5 * dpqs + zfac5 * pqs5 * zcor + zcor5 * pqs5 * zfac;
  const double zdqsdtemp5 = zfac5 * zcor5 * pqs5;

  // critical humidity, ice supersaturation (:505-539)
  const double zcrh2 = crit_rh(crh, CSC2_CETA(jk), CSC2_SQ1MCETA(jk));
  const bool vcold = ztp15 < c.rtice;
  const double zsupsat5 = vcold ? 1.8 - 3.e-03 * ztp15 : 1.0;
  const double zsupsat = vcold ? -3.e-03 * ztp1 : 0.0;
  const double zqsat5 = pqs5 * zsupsat5;
  const double zqsat = dpqs * zsupsat5 + pqs5 * zsupsat;
  const double zqcrit5 = zcrh2 * zqsat5, zqcrit = zcrh2 * zqsat;

  // uniform distribution (:543-593)
  const double zscalm = CSC2_ZSCALM(jk);
  const double zqt = zqp1 + zl + zi, zqt5 = zqp15 + zl5 + zi5;
  double pclc, pclc5, zqc, zqc5;
  {
    const bool clear = zqt5 <= zqcrit5, overcast = zqt5 >= zqsat5;
    const bool partial = !clear && !overcast;
    const double zqpd = zqsat - zqt, zqpd5 = zqsat5 - zqt5;
    const double zqcd = zqsat - zqcrit, zqcd5 = zqsat5 - zqcrit5;
    // guarded operands outside the partial branch (results are discarded there)
    const double den5 = partial ? zqcd5 - zscalm * (zqt5 - zqcrit5) : 1.0;
    const double zqpd5g = partial ? zqpd5 : 1.0;
    const double den5_inv = csc2_rcp(den5);
    const double zsqrt5 = csc2_sqrt(zqpd5g * den5_inv);
    const double pc5 = 1.0 - zsqrt5;
    double pc = -(0.5 * csc2_rcp(zsqrt5)) * (zqpd * den5 - zqpd5 * (zqcd - zscalm * (zqt - zqcrit))) *
                (den5_inv * den5_inv);
    if (lreg) {   // :575-580
      const double zrat = zqpd5g * csc2_rcp(partial ? zqcd5 : 1.0);
      const double b = 1.0 - zscalm * (1.0 - zrat);
      const double zyyy = dmin_(0.3, 3.5 * csc2_sqrt(zrat * (b * b * b)) * csc2_rcp(1.0 - zscalm));
      pc = zyyy * pc;
    }
    const double m5 = zscalm * zqpd5 + (1.0 - zscalm) * zqcd5;
    const double qcp = (zscalm * zqpd + (1.0 - zscalm) * zqcd) * (pc5 * pc5) + m5 * 2.0 * pc5 * pc;
    const double qcp5 = m5 * (pc5 * pc5);
    pclc = partial ? pc : 0.0;
    pclc5 = partial ? pc5 : (overcast ? 1.0 : 0.0);
    zqc = partial ? qcp : (overcast ? (1.0 - zscalm) * zqcd : 0.0);
    zqc5 = partial ? qcp5 : (overcast ? (1.0 - zscalm) * zqcd5 : 0.0);
  }

  // convective component (:597-628)
  const double zdp5_inv = csc2_rcp(zdp5);
  const double zgdp5 = c.rg * zdp5_inv;
  const double zgdp = -zgdp5 * zdp * zdp5_inv;
  const double zlude5 = x5.plude * dt * zgdp5;
  const double zlude = dt * zgdp5 * dx.plude + dt * x5.plude * zgdp;
  {
    const bool llo1 = jk < c.klev - 1 && zlude5 >= c.rlmin && x5.plu1 >= CSC2_ZEPS2;
    const double plu_inv = csc2_rcp(llo1 ? x5.plu1 : 1.0);
    const double e = csc2_expn(-zlude5 * plu_inv);
    const double pcn = pclc - pclc * (1.0 - e) + ((1.0 - pclc5) * plu_inv) * e * zlude -
                       ((1.0 - pclc5) * zlude5 * (plu_inv * plu_inv)) * e * dx.plu1;
    const double pcn5 = pclc5 + (1.0 - pclc5) * (1.0 - e);
    pclc = llo1 ? pcn : pclc;
    pclc5 = llo1 ? pcn5 : pclc5;
    zqc = llo1 ? zqc + zlude : zqc;
    zqc5 = llo1 ? zqc5 + zlude5 : zqc5;
  }

  // compensating subsidence (:632-669)
  {
    const double zfac1 = csc2_rcp(c.rd * ztp15);
    const double ztp15_inv = c.rd * zfac1;
    const double zrho = (dx.pap - ztp1 * x5.pap * ztp15_inv) * zfac1;
    const double zrho5 = x5.pap * zfac1;
    const double zfac2 = csc2_rcp(x5.pap - c.retv * zfoeew5);
    const double zrodqsdp = (-zrho * pqs5 - zrho5 * dpqs +
                             zrho5 * pqs5 * (dx.pap - c.retv * zfoeew) * zfac2) * zfac2;
    const double zrodqsdp5 = -zrho5 * pqs5 * zfac2;
    const double zldcp = zfwat * zlvdcp5 + zfwat5 * zlvdcp + (1.0 - zfwat5) * zlsdcp - zfwat * zlsdcp5;
    const double zldcp5 = zfwat5 * zlvdcp5 + (1.0 - zfwat5) * zlsdcp5;
    const double zfac3 = csc2_rcp(1.0 + zldcp5 * zdqsdtemp5);
    const double dtdzmo5 = c.rg * (c.rcpd_inv - zldcp5 * zrodqsdp5) * zfac3;
    const double dtdzmo = -(c.rg * (zldcp * zrodqsdp5 + zldcp5 * zrodqsdp) +
                            dtdzmo5 * (zldcp5 * zdqsdtemp + zldcp * zdqsdtemp5)) * zfac3;
    const double zdqsdz = zdqsdtemp5 * dtdzmo + zdqsdtemp * dtdzmo5 - c.rg * zrodqsdp;
    const double zdqsdz5 = zdqsdtemp5 * dtdzmo5 - c.rg * zrodqsdp5;
    const double zfac4 = c.rd * ztp15 * pap5_inv;   // 1/ZRHO5
    const double mf5 = x5.pmfu + x5.pmfd;
    const double zdqc5t = zdqsdz5 * mf5 * dt * zfac4;
    const bool llo3 = zdqc5t < zqc5;
    double zdqct = (dt * (zdqsdz * mf5 + zdqsdz5 * (dx.pmfu + dx.pmfd)) - zdqc5t * zrho) * zfac4;
    if (lreg) zdqct = zdqct * 0.1;   // :657
    const double zdqc = llo3 ? zdqct : zqc;
    const double zdqc5 = llo3 ? zdqc5t : zqc5;
    zqc = zqc - zdqc;
    zqc5 = zqc5 - zdqc5;
  }

  // condensate and condensation rates (:673-685)
  double zqlwc = zqc * zfwat5 + zqc5 * zfwat, zqlwc5 = zqc5 * zfwat5;
  double zqiwc = zqc * (1.0 - zfwat5) - zqc5 * zfwat, zqiwc5 = zqc5 * (1.0 - zfwat5);
  double zcondl = (zqlwc - zl) * c.zqtmst, zcondl5 = (zqlwc5 - zl5) * c.zqtmst;
  double zcondi = (zqiwc - zi) * c.zqtmst, zcondi5 = (zqiwc5 - zi5) * c.zqtmst;

  // melting of incoming snow (:707-738)
  double zrfln, zrfln5, zsfln, zsfln5;
  {
    const bool melt = st5.sfl != 0.0;
    const double lf5_inv = csc2_rcp(zlfdcp5);
    const double zcons5 = c.zcons2 * zdp5 * lf5_inv;
    const double zcons = c.zcons2 * (zdp * zlfdcp5 - zdp5 * zlfdcp) * (lf5_inv * lf5_inv);
    const bool warm2 = (ztp15 - c.zmeltp2) > 0.0;
    const double zz2s = warm2 ? zcons5 * ztp1 + zcons * (ztp15 - c.zmeltp2) : 0.0;
    const double zz2s5 = warm2 ? zcons5 * (ztp15 - c.zmeltp2) : 0.0;
    const bool all = st5.sfl <= zz2s5;
    const double zsnmlt = melt ? (all ? st.sfl : zz2s) : 0.0;
    const double zsnmlt5 = melt ? (all ? st5.sfl : zz2s5) : 0.0;
    zrfln = st.rfl + zsnmlt;   zrfln5 = st5.rfl + zsnmlt5;
    zsfln = st.sfl - zsnmlt;   zsfln5 = st5.sfl - zsnmlt5;
    const double zcons5_inv = csc2_rcp(zcons5);
    ztp1 = ztp1 - (zsnmlt * zcons5 - zcons * zsnmlt5) * (zcons5_inv * zcons5_inv);
    ztp15 = ztp15 - zsnmlt5 * zcons5_inv;
  }

  // autoconversion (:742-819)
  double zprr, zprr5, zprs, zprs5;
  {
    const bool cloudy = pclc5 > CSC2_ZEPS2;
    const double pclc5_inv = csc2_rcp(cloudy ? pclc5 : 1.0);
    const double rl2 = c.rlcrit_inv * c.rlcrit_inv;
    const double zcldl5 = zqlwc5 * pclc5_inv;
    const double zcldl = zqlwc * pclc5_inv - zcldl5 * pclc * pclc5_inv;
    const double zexp35 = csc2_expn(-SQ_(zcldl5 * c.rlcrit_inv));
    const double zdl5 = c.zckcodtl * (1.0 - zexp35);
    const double zexpdl5 = csc2_exp(-zdl5);
    const double zdl = (2.0 * (lreg ? c.zckcodtla : c.zckcodtl) * rl2) * zexp35 * zcldl5 * zcldl;
    const double zlnew = zcldl5 * zexpdl5 * pclc + pclc5 * zexpdl5 * zcldl -
                         pclc5 * zcldl5 * zexpdl5 * zdl;
    const double zlnew5 = pclc5 * zcldl5 * zexpdl5;
    zprr = cloudy ? zqlwc - zlnew : 0.0;     zprr5 = cloudy ? zqlwc5 - zlnew5 : 0.0;
    zqlwc = zqlwc - zprr;                    zqlwc5 = zqlwc5 - zprr5;

    const double zcldi5 = zqiwc5 * pclc5_inv;
    const double zcldi = zqiwc * pclc5_inv - zcldi5 * pclc * pclc5_inv;
    const double zexp15 = csc2_exp(0.025 * (ztp15 - c.rtt));
    const double zexp25 = csc2_expn(-SQ_(zcldi5 * c.rlcrit_inv));
    const double zdi5 = c.zckcodti * zexp15 * (1.0 - zexp25);
    const double zexpdi5 = csc2_exp(-zdi5);
    const double zdi = (lreg ? c.zckcodtia : c.zckcodti) * zexp15 *
                       (zexp25 * (2.0 * zcldi5 * zcldi * rl2 - 0.025 * ztp1) + 0.025 * ztp1);
    const double zinew = zcldi5 * zexpdi5 * pclc + pclc5 * zexpdi5 * zcldi -
                         pclc5 * zcldi5 * zexpdi5 * zdi;
    const double zinew5 = pclc5 * zcldi5 * zexpdi5;
    zprs = cloudy ? zqiwc - zinew : 0.0;     zprs5 = cloudy ? zqiwc5 - zinew5 : 0.0;
    zqiwc = zqiwc - zprs;                    zqiwc5 = zqiwc5 - zprs5;
  }

  // new precipitation (:823-843)
  const double zdr = c.zcons2 * (zdp5 * (zprr + zprs) + zdp * (zprr5 + zprs5));
  const double zdr5 = c.zcons2 * zdp5 * (zprr5 + zprs5);
  const bool frz1 = ztp15 < c.rtt;
  double zrfreeze5 = frz1 ? c.zcons2 * zdp5 * zprr5 : 0.0;
  double zrfreeze = frz1 ? c.zcons2 * (zdp * zprr5 + zdp5 * zprr) : 0.0;
  zsfln += frz1 ? zdr : 0.0;    zsfln5 += frz1 ? zdr5 : 0.0;
  zrfln += frz1 ? 0.0 : zdr;    zrfln5 += frz1 ? 0.0 : zdr5;

  // incrementation of T and q (:949-989)
  const double zldw5 = zfwat5 * zlvdcp5 + (1.0 - zfwat5) * zlsdcp5;
  const double zldw = zfwat * (zlvdcp5 - zlsdcp5) + (zfwat5 * zlvdcp + (1.0 - zfwat5) * zlsdcp);
  {
    const double zdqdt = -(zcondl + zcondi) + dx.plude * zgdp5 + x5.plude * zgdp;
    const double zdqdt5 = -(zcondl5 + zcondi5) + x5.plude * zgdp5;
    const double br5 = x5.plude * zldw5 - (zlsdcp5 - zlvdcp5) * zrfreeze5;
    const double zdtdt = zlvdcp * zcondl5 + zlsdcp * zcondi5 + zlvdcp5 * zcondl + zlsdcp5 * zcondi -
                         (dx.plude * zldw5 + x5.plude * zldw - (zlsdcp - zlvdcp) * zrfreeze5 -
                          (zlsdcp5 - zlvdcp5) * zrfreeze) * zgdp5 -
                         br5 * zgdp;
    const double zdtdt5 = zlvdcp5 * zcondl5 + zlsdcp5 * zcondi5 - br5 * zgdp5;
    ztp1 += dt * zdtdt;     ztp15 += dt * zdtdt5;
    zqp1 += dt * zdqdt;     zqp15 += dt * zdqdt5;
  }
  const double zqold = zqp1, zqold5 = zqp15;

  // saturation adjustment (:993-997)
  cuadjtqstl_point(c, pap5_inv, dx.pap, ztp15, zqp15, ztp1, zqp1);

  // excess to precipitation (:999-1046)
  {
    const bool exc = (zqold5 - zqp15) >= 0.0;
    const double zdq5 = exc ? zqold5 - zqp15 : 0.0;
    double zdq = exc ? zqold - zqp1 : 0.0;
    if (lreg) zdq = zdq * 0.7;
    const double zdr2 = c.zcons2 * (zdp5 * zdq + zdq5 * zdp);
    const double zdr25 = c.zcons2 * zdp5 * zdq5;
    const bool frz2 = ztp15 < c.rtt;
    zrfreeze5 += frz2 ? zfwat5 * zdr25 : 0.0;
    zrfreeze += frz2 ? zfwat * zdr25 + zfwat5 * zdr2 : 0.0;
    zcondi += frz2 ? zdq * c.zqtmst : 0.0;     zcondi5 += frz2 ? zdq5 * c.zqtmst : 0.0;
    zcondl += frz2 ? 0.0 : zdq * c.zqtmst;     zcondl5 += frz2 ? 0.0 : zdq5 * c.zqtmst;
    zsfln += frz2 ? zdr2 : 0.0;                zsfln5 += frz2 ? zdr25 : 0.0;
    zrfln += frz2 ? 0.0 : zdr2;                zrfln5 += frz2 ? 0.0 : zdr25;
  }

  // final tendencies (:1048-1096)
  {
    const double br5 = x5.plude * zldw5 - (zlsdcp5 - zlvdcp5) * zrfreeze5;
    dy.tenq = -(zcondl + zcondi) + dx.plude * zgdp5 + x5.plude * zgdp;
    y5.tenq = -(zcondl5 + zcondi5) + x5.plude * zgdp5;
    dy.tent = zlvdcp * zcondl5 + zlsdcp * zcondi5 + zlvdcp5 * zcondl + zlsdcp5 * zcondi -
              (dx.plude * zldw5 + x5.plude * zldw - (zlsdcp - zlvdcp) * zrfreeze5 -
               (zlsdcp5 - zlvdcp5) * zrfreeze) * zgdp5 -
              br5 * zgdp;
    y5.tent = zlvdcp5 * zcondl5 + zlsdcp5 * zcondi5 - br5 * zgdp5;
  }
  dy.tenl = (zqlwc - zl) * c.zqtmst;     y5.tenl = (zqlwc5 - zl5) * c.zqtmst;
  dy.teni = (zqiwc - zi) * c.zqtmst;     y5.teni = (zqiwc5 - zi5) * c.zqtmst;
  dy.pclc = pclc;                        y5.pclc = pclc5;
  dy.rfln = zrfln;                       y5.rfln = zrfln5;
  dy.sfln = zsfln;                       y5.sfln = zsfln5;
  st.rfl = zrfln;   st.sfl = zsfln;   st.paph0 = dx.paph1;
  st5.rfl = zrfln5; st5.sfl = zsfln5; st5.paph0 = x5.paph1;
}
