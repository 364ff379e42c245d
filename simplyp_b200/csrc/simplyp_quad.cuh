// simplyp_quad.cuh — one ensemble member integrated by a QUAD of 4 adjacent lanes.
//
// Why: with one thread per member a 10^4-member ensemble is 313 warps on 592 SM sub-partitions and every
// thread is one long dependent fp64 chain (measured 3.3 cycles per issued instruction, 8.1-cycle DFMA
// latency): the run is latency-bound and most of the machine idles.  ode_f (model.py:58-187) however has
// four-fold structure: two identical soil boxes, a groundwater box and a reach whose three in-stream masses
// obey the same linear equation.  Here the 11 integrated components are dealt to 4 lanes
//
//      lane 0: VsA , Msus , Msus_out        own exp: exp(-mu VsA)     own gate: soil A   -> QsA
//      lane 1: VsS , TDPr , TDPr_out        own exp: exp(-mu VsS)     own gate: soil S   -> QsS
//      lane 2: Vg  , PPr  , PPr_out         own exp: Qr^k_M           own gate: groundwater -> Qg
//      lane 3: Qr  ,  -   , Qr_av           own exp: Qr^b_Q           (no gate)
//
// and every lane executes the SAME instruction stream on lane-specific coefficients (LaneCoef), so there is
// no divergence inside a quad: one broadcast of Qr, one log, ONE exp and ONE gate per lane per RHS evaluation
// instead of 4 exps + 3 gates per thread, five more broadcasts (QsA, QsS, Qg, Qr^b, Qr^k) through warp
// shuffles, then two short linear forms.  Runge-Kutta combinations, error norm (butterfly-reduced over the
// quad) and accumulators are per lane.  Registers per lane drop to about a third, the per-step dependent
// chain to about a quarter.  The once-a-day algebra (begin_day / end_day of simplyp_core.cuh, statistics,
// output row, routing) is unchanged and runs on lane 0 of the quad, which gathers/scatters the state through
// shared memory.
#pragma once

#include "simplyp_core.cuh"

namespace simplyp {

struct LaneCoef {
  double eY, eL;                    // own exp argument: eY*yA + eL*ln(Qr)
  double gx1, gx0, gu, g0, g1;      // own gate: x = yA*gx1 + gx0 ; G = g0 + gate(x*gu)*x*g1
  double a0, aE, aSA, aSS, aG, aR;  // slot A: dA = (a0 + aE*e + aSA*QsA + aSS*QsS + aG*Qg + aR*Qr) * (mulqb ? Qr^b : 1)
  double b0, bK, bSA, bSS, bG;      // slot B: dB = b0 + bK*Qr^k + bSA*QsA + bSS*QsS + bG*Qg - yB*r
  double cR;                        // r = cR*Qr^b  (= Qr/Vr)
  double accQ;                      // accumulator: dacc = yB*r + accQ*Qr
  int mulqb;
};

// Lane-specific coefficients of one day from the per-member constants (same algebra as rhs()).
SP_HD void build_lane_coef(const Hot& h, int ql, LaneCoef& c) {
  c.eY = c.eL = 0.0;
  c.gx1 = c.gx0 = c.gu = c.g0 = c.g1 = 0.0;
  c.a0 = c.aE = c.aSA = c.aSS = c.aG = c.aR = 0.0;
  c.b0 = c.bK = c.bSA = c.bSS = c.bG = 0.0;
  c.cR = h.cR;
  c.accQ = 0.0;
  c.mulqb = 0;
  if (ql == 0 || ql == 1) {                 // soil boxes (:105-110)
    c.eY = -h.mu;
    c.gx1 = 1.0; c.gx0 = -h.fc; c.gu = h.inv_fcd; c.g1 = (ql == 0) ? h.inv_TsA : h.inv_TsS;
    c.a0 = h.Pin - h.aE; c.aE = h.aE;
    if (ql == 0) { c.aSA = -1.0; c.b0 = h.MsusUS; c.bK = h.cM; }                       // sediment (:138-147)
    else         { c.aSS = -1.0; c.b0 = h.t0; c.bSA = h.tA; c.bSS = h.tS; c.bG = h.tG; }  // TDP (:154-168)
  } else if (ql == 2) {                     // groundwater (:121-124) + PP (:171-180)
    c.eL = h.kM;
    c.gx1 = h.inv_Tg; c.gx0 = -h.Qg_min; c.gu = h.inv_Qgd; c.g0 = h.Qg_min; c.g1 = 1.0;
    c.aSA = h.beta * h.fA; c.aSS = h.beta * h.fS; c.aG = -1.0;
    c.b0 = h.PPUS; c.bK = h.cP;
  } else {                                  // reach flow (:127-132)
    c.eL = h.bQ;
    const double omb = 1.0 - h.beta;
    c.a0 = h.kQ * h.qin0; c.aSA = h.kQ * omb * h.fA; c.aSS = h.kQ * omb * h.fS; c.aG = h.kQ; c.aR = -h.kQ;
    c.mulqb = 1;
    c.accQ = 1.0;
  }
}

// First half of a RHS evaluation on one lane: the lane's own exponential and gated flow.
SP_HD void lane_phase1(const LaneCoef& c, double yA, double lq, double& e, double& G) {
  e = sp_exp_core(fma(c.eY, yA, c.eL * lq));
  const double x = fma(yA, c.gx1, c.gx0);
  G = fma(gate(x * c.gu) * x, c.g1, c.g0);
}

// Second half, after the quad has exchanged QsA, QsS, Qg, Qr^b (qb) and Qr^k (qk).
SP_HD void lane_phase2(const LaneCoef& c, double yB, double e, double QsA, double QsS, double Qg, double Qr,
                       double qb, double qk, double& dA, double& dB, double& dacc) {
  const double L = fma(c.aR, Qr, fma(c.aG, Qg, fma(c.aSS, QsS, fma(c.aSA, QsA, fma(c.aE, e, c.a0)))));
  dA = c.mulqb ? L * qb : L;
  const double out = yB * (c.cR * qb);
  dB = fma(c.bG, Qg, fma(c.bSS, QsS, fma(c.bSA, QsA, fma(c.bK, qk, c.b0)))) - out;
  dacc = fma(c.accQ, Qr, out);
}

// Which of the 7 live states / 4 accumulators a lane owns (index into y[NL] / acc[NA]; -1 = none).
SP_HD int quad_slotA(int ql) { return ql == 0 ? iVsA : (ql == 1 ? iVsS : (ql == 2 ? iVg : iQr)); }
SP_HD int quad_slotB(int ql) { return ql == 0 ? iMsus : (ql == 1 ? iTDPr : (ql == 2 ? iPPr : -1)); }
SP_HD int quad_acc(int ql) { return ql == 0 ? 1 : (ql == 1 ? 2 : (ql == 2 ? 3 : 0)); }

// Per-lane Runge-Kutta state of a quad member.
struct LaneRK {
  double yA, yB, acc;       // the lane's slots
  double k1A, k1B, a1;      // derivatives at the current point (FSAL / reused after a rejection)
};

#if defined(__CUDACC__)
// One embedded RK5(4) step attempt of a quad (device).  All 4 lanes of the quad call this together
// (`qmask` = their bits in the warp, `q0` = lane index of the quad's lane 0).  Returns the scaled RMS error
// of the member (identical on the 4 lanes).
struct QuadEval {
  unsigned qmask;
  int q0;
  __device__ __forceinline__ void operator()(const LaneCoef& c, double yA, double yB, double& dA, double& dB,
                                             double& dacc) const {
    const double Qr = __shfl_sync(qmask, yA, q0 + 3);
    const double lq = sp_log(Qr);
    double e, G;
    lane_phase1(c, yA, lq, e, G);
    const double QsA = __shfl_sync(qmask, G, q0 + 0);
    const double QsS = __shfl_sync(qmask, G, q0 + 1);
    const double Qg = __shfl_sync(qmask, G, q0 + 2);
    const double qk = __shfl_sync(qmask, e, q0 + 2);
    const double qb = __shfl_sync(qmask, e, q0 + 3);
    lane_phase2(c, yB, e, QsA, QsS, Qg, Qr, qb, qk, dA, dB, dacc);
  }
};

__device__ __forceinline__ double quad_attempt(const LaneCoef& c, const QuadEval& f, const LaneRK& s, double hh,
                                               double rtol, double atol, double& ynA, double& ynB, double& accn,
                                               double& k7A, double& k7B, double& a7) {
  using namespace dp;
  double kA2, kA3, kA4, kA5, kA6, kB2, kB3, kB4, kB5, kB6, da;
  double sb = b1 * s.a1, se = e1 * s.a1;
  f(c, fma(hh, a21 * s.k1A, s.yA), fma(hh, a21 * s.k1B, s.yB), kA2, kB2, da);
  sb = fma(b2, da, sb); se = fma(e2, da, se);
  f(c, fma(hh, fma(a32, kA2, a31 * s.k1A), s.yA), fma(hh, fma(a32, kB2, a31 * s.k1B), s.yB), kA3, kB3, da);
  sb = fma(b3, da, sb); se = fma(e3, da, se);
  f(c, fma(hh, fma(a43, kA3, fma(a42, kA2, a41 * s.k1A)), s.yA),
       fma(hh, fma(a43, kB3, fma(a42, kB2, a41 * s.k1B)), s.yB), kA4, kB4, da);
  sb = fma(b4, da, sb); se = fma(e4, da, se);
  f(c, fma(hh, fma(a54, kA4, fma(a53, kA3, fma(a52, kA2, a51 * s.k1A))), s.yA),
       fma(hh, fma(a54, kB4, fma(a53, kB3, fma(a52, kB2, a51 * s.k1B))), s.yB), kA5, kB5, da);
  sb = fma(b5, da, sb); se = fma(e5, da, se);
  f(c, fma(hh, fma(a65, kA5, fma(a64, kA4, fma(a63, kA3, fma(a62, kA2, a61 * s.k1A)))), s.yA),
       fma(hh, fma(a65, kB5, fma(a64, kB4, fma(a63, kB3, fma(a62, kB2, a61 * s.k1B)))), s.yB), kA6, kB6, da);
  sb = fma(b6, da, sb); se = fma(e6, da, se);
  ynA = fma(hh, fma(b6, kA6, fma(b5, kA5, fma(b4, kA4, fma(b3, kA3, fma(b2, kA2, b1 * s.k1A))))), s.yA);
  ynB = fma(hh, fma(b6, kB6, fma(b5, kB5, fma(b4, kB4, fma(b3, kB3, fma(b2, kB2, b1 * s.k1B))))), s.yB);
  f(c, ynA, ynB, k7A, k7B, a7);
  accn = fma(hh, sb, s.acc);
  const double eA = hh * fma(e7, k7A, fma(e6, kA6, fma(e5, kA5, fma(e4, kA4, fma(e3, kA3, fma(e2, kA2, e1 * s.k1A))))));
  const double eB = hh * fma(e7, k7B, fma(e6, kB6, fma(e5, kB5, fma(e4, kB4, fma(e3, kB3, fma(e2, kB2, e1 * s.k1B))))));
  const double ec = hh * fma(e7, a7, se);
  const double qA = eA * sp_rcp_fast(fma(rtol, sp_max(fabs(s.yA), fabs(ynA)), atol));
  const double qB = eB * sp_rcp_fast(fma(rtol, sp_max(fabs(s.yB), fabs(ynB)), atol));
  const double qc = ec * sp_rcp_fast(fma(rtol, sp_max(fabs(s.acc), fabs(accn)), atol));
  double sum = fma(qA, qA, fma(qB, qB, qc * qc));
  sum += __shfl_xor_sync(f.qmask, sum, 1);
  sum += __shfl_xor_sync(f.qmask, sum, 2);
  const double en = sqrt(sum * (1.0 / (NL + NA)));
  return (en == en) ? en : INFINITY;
}
#endif  // __CUDACC__

// Host-side lock-step evaluation of the same per-lane functions (used by the test harness to check the
// lane coefficient mapping and the quad step against rhs()/dp5_attempt()).
struct QuadHost {
  LaneCoef c[4];
  void eval(const double (&yA)[4], const double (&yB)[4], double (&dA)[4], double (&dB)[4], double (&dacc)[4]) const {
    const double Qr = yA[3];
    const double lq = sp_log(Qr);
    double e[4], G[4];
    for (int l = 0; l < 4; ++l) lane_phase1(c[l], yA[l], lq, e[l], G[l]);
    for (int l = 0; l < 4; ++l) lane_phase2(c[l], yB[l], e[l], G[0], G[1], G[2], Qr, e[3], e[2], dA[l], dB[l], dacc[l]);
  }
};

}  // namespace simplyp
