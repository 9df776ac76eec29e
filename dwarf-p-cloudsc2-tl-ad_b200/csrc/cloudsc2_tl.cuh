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
    const double r = 1.0 / (t5 - z4es);
    const double r2 = r * r;
    const double foeew5 = c.r2es * exp(z3es * (t5 - c.rtt) * r);
    const double foeew = k3 * t * foeew5 * r2;
    double qsat = zqp5 * foeew + zqp * foeew5;
    double qsat5 = zqp5 * foeew5;
    if (qsat5 > CSC2_ZQMAX) { qsat = 0.0; qsat5 = CSC2_ZQMAX; }
    const double cor5 = 1.0 / (1.0 - c.retv * qsat5);
    const double cor = (c.retv * qsat) * (cor5 * cor5);
    qsat = qsat5 * cor + qsat * cor5;
    qsat5 = qsat5 * cor5;
    const double z2s5 = z5alcp * r2;
    const double z2s = -2.0 * t * z2s5 * r;
    const double den = 1.0 / (1.0 + qsat5 * cor5 * z2s5);
    const double cond5 = (q5 - qsat5) * den;
    const double cond = (q - qsat) * den -
                        cond5 * (qsat * cor5 * z2s5 + qsat5 * cor * z2s5 + qsat5 * cor5 * z2s) * den;
    t += zaldcp * cond;   t5 += zaldcp * cond5;
    q -= cond;            q5 -= cond5;
  }
}

// x5/pqs5 : trajectory inputs ; dx/dpqs : perturbations.  y5/dy : outputs.
__device__ __forceinline__ void tl_level(const KConst &c, const CritRH &crh, int jk,
                                         const LevIn &x5, double pqs5, const LevIn &dx, double dpqs,
                                         Carry &st5, CarryTL &st, LevOut &y5, LevOut &dy) {
  const double dt = c.ptsphy;
  const bool lreg = c.lregcl != 0;
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
  if (c.rvtmp2 != 0.0) {
    zzz5 = 1.0 / (c.rcpd + c.rcpd * c.rvtmp2 * zqp15);
    zzz = -c.rcpd * c.rvtmp2 * zqp1 * (zzz5 * zzz5);
  }
  const double zlfdcp = c.rlmlt * zzz, zlfdcp5 = c.rlmlt * zzz5;
  const double zlsdcp = c.rlstt * zzz, zlsdcp5 = c.rlstt * zzz5;
  const double zlvdcp = c.rlvtt * zzz, zlvdcp5 = c.rlvtt * zzz5;
  const double pap5_inv = 1.0 / x5.pap;

  // dqs/dT correction factor (:463-491)
  const double rw = 1.0 / (ztp15 - c.r4les), ri = 1.0 / (ztp15 - c.r4ies);
  double zfwat, zfwat5, zfoeew5, zfoeew;
  if (ztp15 < c.rtt) {
    const double a = 0.17 * (ztp15 - c.rlptrc);
    const double ch = cosh(a);
    zfwat = 0.545 * 0.17 * ztp1 / (ch * ch);
    zfwat5 = 0.545 * (tanh(a) + 1.0);
    zfoeew5 = c.r2es * exp(c.r3ies * (ztp15 - c.rtt) * ri);
    zfoeew = c.r3ies * (c.rtt - c.r4ies) * ztp1 * zfoeew5 * (ri * ri);
  } else {
    zfwat = 0.0;
    zfwat5 = 1.0;
    zfoeew5 = c.r2es * exp(c.r3les * (ztp15 - c.rtt) * rw);
    zfoeew = c.r3les * (c.rtt - c.r4les) * ztp1 * zfoeew5 * (rw * rw);
  }
  double zesdp = zfoeew * pap5_inv - dx.pap * zfoeew5 * (pap5_inv * pap5_inv);
  double zesdp5 = zfoeew5 * pap5_inv;
  if (zesdp5 > CSC2_ZQMAX) { zesdp = 0.0; zesdp5 = CSC2_ZQMAX; }
  const double zfacw5 = c.r5les * (rw * rw), zfaci5 = c.r5ies * (ri * ri);
  const double zfacw = -2.0 * ztp1 * zfacw5 * rw, zfaci = -2.0 * ztp1 * zfaci5 * ri;
  const double zfac = zfwat5 * zfacw + zfacw5 * zfwat + (1.0 - zfwat5) * zfaci - zfaci5 * zfwat;
  const double zfac5 = zfwat5 * zfacw5 + (1.0 - zfwat5) * zfaci5;
  const double zcor5 = 1.0 / (1.0 - c.retv * zesdp5);
  const double zcor = c.retv * zesdp * (zcor5 * zcor5);
  const double zdqsdtemp = zfac5 * zcor5 * dpqs + zfac5 * pqs5 * zcor + zcor5 * pqs5 * zfac;
  const double zdqsdtemp5 = zfac5 * zcor5 * pqs5;

  // critical humidity, ice supersaturation (:505-539)
  const double zcrh2 = crit_rh(crh, c.ceta[jk]);
  double zsupsat5 = 1.0, zsupsat = 0.0;
  if (ztp15 < c.rtice) { zsupsat5 = 1.8 - 3.e-03 * ztp15; zsupsat = -3.e-03 * ztp1; }
  const double zqsat5 = pqs5 * zsupsat5;
  const double zqsat = dpqs * zsupsat5 + pqs5 * zsupsat;
  const double zqcrit5 = zcrh2 * zqsat5, zqcrit = zcrh2 * zqsat;

  // uniform distribution (:543-593)
  const double zscalm = c.zscalm[jk];
  const double zqt = zqp1 + zl + zi, zqt5 = zqp15 + zl5 + zi5;
  double pclc, pclc5, zqc, zqc5;
  if (zqt5 <= zqcrit5) {
    pclc = 0.0; pclc5 = 0.0; zqc = 0.0; zqc5 = 0.0;
  } else if (zqt5 >= zqsat5) {
    pclc = 0.0; pclc5 = 1.0;
    zqc = (1.0 - zscalm) * (zqsat - zqcrit);
    zqc5 = (1.0 - zscalm) * (zqsat5 - zqcrit5);
  } else {
    const double zqpd = zqsat - zqt, zqpd5 = zqsat5 - zqt5;
    const double zqcd = zqsat - zqcrit, zqcd5 = zqsat5 - zqcrit5;
    const double den5 = zqcd5 - zscalm * (zqt5 - zqcrit5);
    const double den5_inv = 1.0 / den5;
    const double zsqrt5 = sqrt(zqpd5 * den5_inv);
    pclc5 = 1.0 - zsqrt5;
    pclc = -(0.5 / zsqrt5) * (zqpd * den5 - zqpd5 * (zqcd - zscalm * (zqt - zqcrit))) *
           (den5_inv * den5_inv);
    if (lreg) {   // :575-580
      const double zrat = zqpd5 / zqcd5;
      const double b = 1.0 - zscalm * (1.0 - zrat);
      const double zyyy = dmin_(0.3, 3.5 * sqrt(zrat * (b * b * b)) / (1.0 - zscalm));
      pclc = zyyy * pclc;
    }
    const double m5 = zscalm * zqpd5 + (1.0 - zscalm) * zqcd5;
    zqc = (zscalm * zqpd + (1.0 - zscalm) * zqcd) * (pclc5 * pclc5) + m5 * 2.0 * pclc5 * pclc;
    zqc5 = m5 * (pclc5 * pclc5);
  }

  // convective component (:597-628)
  const double zdp5_inv = 1.0 / zdp5;
  const double zgdp5 = c.rg * zdp5_inv;
  const double zgdp = -zgdp5 * zdp * zdp5_inv;
  const double zlude5 = x5.plude * dt * zgdp5;
  const double zlude = dt * zgdp5 * dx.plude + dt * x5.plude * zgdp;
  if (jk < c.klev - 1 && zlude5 >= c.rlmin && x5.plu1 >= CSC2_ZEPS2) {
    const double plu_inv = 1.0 / x5.plu1;
    const double e = exp(-zlude5 * plu_inv);
    pclc = pclc - pclc * (1.0 - e) + ((1.0 - pclc5) * plu_inv) * e * zlude -
           ((1.0 - pclc5) * zlude5 * (plu_inv * plu_inv)) * e * dx.plu1;
    pclc5 = pclc5 + (1.0 - pclc5) * (1.0 - e);
    zqc = zqc + zlude;
    zqc5 = zqc5 + zlude5;
  }

  // compensating subsidence (:632-669)
  {
    const double zfac1 = 1.0 / (c.rd * ztp15);
    const double ztp15_inv = c.rd * zfac1;
    const double zrho = (dx.pap - ztp1 * x5.pap * ztp15_inv) * zfac1;
    const double zrho5 = x5.pap * zfac1;
    const double zfac2 = 1.0 / (x5.pap - c.retv * zfoeew5);
    const double zrodqsdp = (-zrho * pqs5 - zrho5 * dpqs +
                             zrho5 * pqs5 * (dx.pap - c.retv * zfoeew) * zfac2) * zfac2;
    const double zrodqsdp5 = -zrho5 * pqs5 * zfac2;
    const double zldcp = zfwat * zlvdcp5 + zfwat5 * zlvdcp + (1.0 - zfwat5) * zlsdcp - zfwat * zlsdcp5;
    const double zldcp5 = zfwat5 * zlvdcp5 + (1.0 - zfwat5) * zlsdcp5;
    const double zfac3 = 1.0 / (1.0 + zldcp5 * zdqsdtemp5);
    const double dtdzmo5 = c.rg * (c.rcpd_inv - zldcp5 * zrodqsdp5) * zfac3;
    const double dtdzmo = -(c.rg * (zldcp * zrodqsdp5 + zldcp5 * zrodqsdp) +
                            dtdzmo5 * (zldcp5 * zdqsdtemp + zldcp * zdqsdtemp5)) * zfac3;
    const double zdqsdz = zdqsdtemp5 * dtdzmo + zdqsdtemp * dtdzmo5 - c.rg * zrodqsdp;
    const double zdqsdz5 = zdqsdtemp5 * dtdzmo5 - c.rg * zrodqsdp5;
    const double zfac4 = c.rd * ztp15 * pap5_inv;   // 1/ZRHO5
    const double mf5 = x5.pmfu + x5.pmfd;
    const double zdqc5t = zdqsdz5 * mf5 * dt * zfac4;
    double zdqc, zdqc5;
    if (zdqc5t < zqc5) {   // LLO3
      zdqc5 = zdqc5t;
      zdqc = (dt * (zdqsdz * mf5 + zdqsdz5 * (dx.pmfu + dx.pmfd)) - zdqc5 * zrho) * zfac4;
      if (lreg) zdqc = zdqc * 0.1;   // :657
    } else {
      zdqc5 = zqc5;
      zdqc = zqc;
    }
    zqc = zqc - zdqc;
    zqc5 = zqc5 - zdqc5;
  }

  // condensate and condensation rates (:673-685)
  double zqlwc = zqc * zfwat5 + zqc5 * zfwat, zqlwc5 = zqc5 * zfwat5;
  double zqiwc = zqc * (1.0 - zfwat5) - zqc5 * zfwat, zqiwc5 = zqc5 * (1.0 - zfwat5);
  double zcondl = (zqlwc - zl) * c.zqtmst, zcondl5 = (zqlwc5 - zl5) * c.zqtmst;
  double zcondi = (zqiwc - zi) * c.zqtmst, zcondi5 = (zqiwc5 - zi5) * c.zqtmst;

  // melting of incoming snow (:707-738)
  double zrfln = st.rfl, zrfln5 = st5.rfl, zsfln = st.sfl, zsfln5 = st5.sfl;
  if (st5.sfl != 0.0) {
    const double lf5_inv = 1.0 / zlfdcp5;
    const double zcons5 = c.zcons2 * zdp5 * lf5_inv;
    const double zcons = c.zcons2 * (zdp * zlfdcp5 - zdp5 * zlfdcp) * (lf5_inv * lf5_inv);
    double zz2s = 0.0, zz2s5 = 0.0;
    if ((ztp15 - c.zmeltp2) > 0.0) {
      zz2s = zcons5 * ztp1 + zcons * (ztp15 - c.zmeltp2);
      zz2s5 = zcons5 * (ztp15 - c.zmeltp2);
    }
    double zsnmlt, zsnmlt5;
    if (st5.sfl <= zz2s5) { zsnmlt = st.sfl; zsnmlt5 = st5.sfl; }
    else { zsnmlt = zz2s; zsnmlt5 = zz2s5; }
    zrfln = st.rfl + zsnmlt;   zrfln5 = st5.rfl + zsnmlt5;
    zsfln = st.sfl - zsnmlt;   zsfln5 = st5.sfl - zsnmlt5;
    const double zcons5_inv = 1.0 / zcons5;
    ztp1 = ztp1 - (zsnmlt * zcons5 - zcons * zsnmlt5) * (zcons5_inv * zcons5_inv);
    ztp15 = ztp15 - zsnmlt5 * zcons5_inv;
  }

  // autoconversion (:742-819)
  double zprr = 0.0, zprr5 = 0.0, zprs = 0.0, zprs5 = 0.0;
  if (pclc5 > CSC2_ZEPS2) {
    const double pclc5_inv = 1.0 / pclc5;
    const double rl2 = c.rlcrit_inv * c.rlcrit_inv;
    {
      const double zcldl5 = zqlwc5 * pclc5_inv;
      const double zcldl = zqlwc * pclc5_inv - zcldl5 * pclc * pclc5_inv;
      const double zexp35 = exp(-SQ_(zcldl5 * c.rlcrit_inv));
      const double zd5 = c.zckcodtl * (1.0 - zexp35);
      const double zexpdl5 = exp(-zd5);
      const double zd = (2.0 * (lreg ? c.zckcodtla : c.zckcodtl) * rl2) * zexp35 * zcldl5 * zcldl;
      const double zlnew = zcldl5 * zexpdl5 * pclc + pclc5 * zexpdl5 * zcldl -
                           pclc5 * zcldl5 * zexpdl5 * zd;
      const double zlnew5 = pclc5 * zcldl5 * zexpdl5;
      zprr = zqlwc - zlnew;     zprr5 = zqlwc5 - zlnew5;
      zqlwc = zqlwc - zprr;     zqlwc5 = zqlwc5 - zprr5;
    }
    {
      const double zcldi5 = zqiwc5 * pclc5_inv;
      const double zcldi = zqiwc * pclc5_inv - zcldi5 * pclc * pclc5_inv;
      const double zexp15 = exp(0.025 * (ztp15 - c.rtt));
      const double zexp25 = exp(-SQ_(zcldi5 * c.rlcrit_inv));
      const double zd5 = c.zckcodti * zexp15 * (1.0 - zexp25);
      const double zexpdi5 = exp(-zd5);
      const double zd = (lreg ? c.zckcodtia : c.zckcodti) * zexp15 *
                        (zexp25 * (2.0 * zcldi5 * zcldi * rl2 - 0.025 * ztp1) + 0.025 * ztp1);
      const double zinew = zcldi5 * zexpdi5 * pclc + pclc5 * zexpdi5 * zcldi -
                           pclc5 * zcldi5 * zexpdi5 * zd;
      const double zinew5 = pclc5 * zcldi5 * zexpdi5;
      zprs = zqiwc - zinew;     zprs5 = zqiwc5 - zinew5;
      zqiwc = zqiwc - zprs;     zqiwc5 = zqiwc5 - zprs5;
    }
  }

  // new precipitation (:823-843)
  const double zdr = c.zcons2 * (zdp5 * (zprr + zprs) + zdp * (zprr5 + zprs5));
  const double zdr5 = c.zcons2 * zdp5 * (zprr5 + zprs5);
  double zrfreeze = 0.0, zrfreeze5 = 0.0;
  if (ztp15 < c.rtt) {
    zrfreeze5 = c.zcons2 * zdp5 * zprr5;
    zrfreeze = c.zcons2 * (zdp * zprr5 + zdp5 * zprr);
    zsfln += zdr;   zsfln5 += zdr5;
  } else {
    zrfln += zdr;   zrfln5 += zdr5;
  }

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
    double zdq = 0.0, zdq5 = 0.0;
    if ((zqold5 - zqp15) >= 0.0) {
      zdq5 = zqold5 - zqp15;
      zdq = zqold - zqp1;
      if (lreg) zdq = zdq * 0.7;
    }
    const double zdr2 = c.zcons2 * (zdp5 * zdq + zdq5 * zdp);
    const double zdr25 = c.zcons2 * zdp5 * zdq5;
    if (ztp15 < c.rtt) {
      zrfreeze5 += zfwat5 * zdr25;
      zrfreeze += zfwat * zdr25 + zfwat5 * zdr2;
      zcondi += zdq * c.zqtmst;     zcondi5 += zdq5 * c.zqtmst;
      zsfln += zdr2;                zsfln5 += zdr25;
    } else {
      zcondl += zdq * c.zqtmst;     zcondl5 += zdq5 * c.zqtmst;
      zrfln += zdr2;                zrfln5 += zdr25;
    }
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
