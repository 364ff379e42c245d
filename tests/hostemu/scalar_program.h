// scalar_program.h — TEST HARNESS ONLY (tests/hostemu, scripts/): the one-thread-per-item statement of the daily
// step, i.e. ode_f (model.py:58-187) written out on scalars, a Tsitouras 5(4) attempt on 7 states + 4 quadratures and
// the flattened day/step loop.  It was the round-1 CUDA kernel; the product now ships only the quad program
// (simplyp_b200/csrc/simplyp_quad.cuh), and this independent formulation stays as a cross-check of it on the host
// (tests/test_hostemu_parity.py).  Nothing in simplyp_b200/ includes this file.
#pragma once

#include "../../simplyp_b200/csrc/simplyp_core.cuh"

namespace simplyp {

// ------------------------------------------------------------------------------------------
// ode_f: derivatives of the 8 live states; the 4 accumulator derivatives come out separately.
SP_HD void rhs(const Hot& c, const double (&y)[NL], double (&dy)[NL], double (&da)[NA]) {
  const double VsA = y[iVsA], VsS = y[iVsS], Vg = y[iVg], Qr = y[iQr];
  // soil boxes (:105-110)
  const double xA = VsA - c.fc, xS = VsS - c.fc;
  const double QsA = xA * gate(xA * c.inv_fcd) * c.inv_TsA;
  const double QsS = xS * gate(xS * c.inv_fcd) * c.inv_TsS;
  dy[iVsA] = c.Pin - c.aE * (1.0 - sp_exp_core(-c.mu * VsA)) - QsA;
  dy[iVsS] = c.Pin - c.aE * (1.0 - sp_exp_core(-c.mu * VsS)) - QsS;
  // groundwater (:121-124)
  const double xg = Vg * c.inv_Tg - c.Qg_min;
  const double Qg = c.Qg_min + gate(xg * c.inv_Qgd) * xg;
  const double soil = c.fA * QsA + c.fS * QsS;
  dy[iVg] = c.beta * soil - Qg;
  // reach (:127-132); Qr^b_Q and Qr^k_M share one logarithm
  const double net = c.qin0 + (1.0 - c.beta) * soil + Qg - Qr;
  const double lq = sp_log(Qr);
  const double qb = sp_exp_core(c.bQ * lq);
  const double qk = sp_exp_core(c.kM * lq);
  dy[iQr] = net * c.kQ * qb;
  da[0] = Qr;
  // outflow rate of the reach, 1/day: Qr/Vr with Vr on its invariant curve
  const double r = c.cR * qb;
  // sediment (:138-147)
  const double oM = y[iMsus] * r;
  dy[iMsus] = c.cM * qk + c.MsusUS - oM;
  da[1] = oM;
  // TDP (:154-168)
  const double oT = y[iTDPr] * r;
  dy[iTDPr] = c.tA * QsA + c.tS * QsS + c.tG * Qg + c.t0 - oT;
  da[2] = oT;
  // PP (:171-180)
  const double oP = y[iPPr] * r;
  dy[iPPr] = c.cP * qk + c.PPUS - oP;
  da[3] = oP;
}


// ------------------------------------------------------------------------------------------
// Embedded explicit Runge-Kutta 5(4) with FSAL (Tsitouras' pair; -DSP_DOPRI5 selects Dormand-Prince).
// The accumulators are pure quadratures (the RHS does not depend
// on them), so their stage derivatives are folded into two running sums (5th-order weights and
// error weights) instead of being stored per stage.
struct RK {
  double k1[NL];   // derivative at the current (t, y): reused after a rejection, FSAL after acceptance
  double a1[NA];   // accumulator derivatives at the current point
};


// One step attempt of size hh from (y, acc).  On return ynew/accnew hold the 5th-order solution,
// k7/a7 the derivative there, and the return value is the scaled RMS error (<= 1 accepts);
// a non-finite error is returned as +inf.
// Stage derivatives k2..k5 of the live states, parked between stages.  They stay in registers: a
// shared-memory variant ([stage][state][thread] columns, 168 registers, 3 blocks/SM) was measured 8 % slower
// at 1.6e5 members and 75 % slower at 1e4 members (spills + LDS latency on the critical path).
struct RegStages {
  double k[4][NL];
  SP_HD double ld(int j, int i) const { return k[j][i]; }
  SP_HD void st(int j, int i, double v) { k[j][i] = v; }
};

template <class KS>
SP_HD double dp5_attempt(const Hot& c, const double (&y)[NL], const double (&acc)[NA], const RK& rk, double hh,
                         double rtol, double atol, double (&ynew)[NL], double (&accnew)[NA], double (&k7)[NL],
                         double (&a7)[NA], KS& ks) {
  using namespace dp;
  double kk[NL], yt[NL], da[NA];
  double sb[NA], se[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) { sb[i] = b1 * rk.a1[i]; se[i] = e1 * rk.a1[i]; }

  // stage 2
#pragma unroll
  for (int i = 0; i < NL; ++i) yt[i] = y[i] + hh * (a21 * rk.k1[i]);
  rhs(c, yt, kk, da);
#pragma unroll
  for (int i = 0; i < NL; ++i) ks.st(0, i, kk[i]);
#pragma unroll
  for (int i = 0; i < NA; ++i) { sb[i] += b2 * da[i]; se[i] += e2 * da[i]; }
  // stage 3
#pragma unroll
  for (int i = 0; i < NL; ++i) yt[i] = y[i] + hh * (a31 * rk.k1[i] + a32 * kk[i]);
  rhs(c, yt, kk, da);
#pragma unroll
  for (int i = 0; i < NL; ++i) ks.st(1, i, kk[i]);
#pragma unroll
  for (int i = 0; i < NA; ++i) { sb[i] += b3 * da[i]; se[i] += e3 * da[i]; }
  // stage 4
#pragma unroll
  for (int i = 0; i < NL; ++i) yt[i] = y[i] + hh * (a41 * rk.k1[i] + a42 * ks.ld(0, i) + a43 * kk[i]);
  rhs(c, yt, kk, da);
#pragma unroll
  for (int i = 0; i < NL; ++i) ks.st(2, i, kk[i]);
#pragma unroll
  for (int i = 0; i < NA; ++i) { sb[i] += b4 * da[i]; se[i] += e4 * da[i]; }
  // stage 5
#pragma unroll
  for (int i = 0; i < NL; ++i)
    yt[i] = y[i] + hh * (a51 * rk.k1[i] + a52 * ks.ld(0, i) + a53 * ks.ld(1, i) + a54 * kk[i]);
  rhs(c, yt, kk, da);
#pragma unroll
  for (int i = 0; i < NL; ++i) ks.st(3, i, kk[i]);
#pragma unroll
  for (int i = 0; i < NA; ++i) { sb[i] += b5 * da[i]; se[i] += e5 * da[i]; }
  // stage 6
#pragma unroll
  for (int i = 0; i < NL; ++i)
    yt[i] = y[i] + hh * (a61 * rk.k1[i] + a62 * ks.ld(0, i) + a63 * ks.ld(1, i) + a64 * ks.ld(2, i) + a65 * kk[i]);
  rhs(c, yt, kk, da);
#pragma unroll
  for (int i = 0; i < NA; ++i) { sb[i] += b6 * da[i]; se[i] += e6 * da[i]; }
  // 5th-order solution and the part of the error estimate that does not need k7 (kk holds k6)
  double ee[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    const double k2 = ks.ld(0, i), k3 = ks.ld(1, i), k4 = ks.ld(2, i), k5 = ks.ld(3, i);
    ynew[i] = y[i] + hh * (b1 * rk.k1[i] + b2 * k2 + b3 * k3 + b4 * k4 + b5 * k5 + b6 * kk[i]);
    ee[i] = e1 * rk.k1[i] + e2 * k2 + e3 * k3 + e4 * k4 + e5 * k5 + e6 * kk[i];
  }
  rhs(c, ynew, k7, a7);

  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    const double err = hh * (ee[i] + e7 * k7[i]);
    const double sc = atol + rtol * sp_max(fabs(y[i]), fabs(ynew[i]));
    const double q = err * sp_rcp_fast(sc);
    s += q * q;
  }
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    accnew[i] = acc[i] + hh * sb[i];
    const double err = hh * (se[i] + e7 * a7[i]);
    const double sc = atol + rtol * sp_max(fabs(acc[i]), fabs(accnew[i]));
    const double q = err * sp_rcp_fast(sc);
    s += q * q;
  }
  const double en = sqrt(s * (1.0 / (NL + NA)));
  return (en == en) ? en : INFINITY;   // NaN -> reject
}

// Step-size factor of the elementary controller for a 5(4) pair: 0.9*err^(-1/5) in [0.2, 5].
SP_HD double step_factor(double en) {
  if (!(en > 1e-30)) return 5.0;
  if (!(en < 1e30)) return 0.2;
#if defined(__CUDA_ARCH__)
  const double f = 0.9 * (double)exp2f(-0.2f * __log2f((float)en));
#else
  const double f = 0.9 * exp(-0.2 * log(en));
#endif
  return sp_min(5.0, sp_max(0.2, f));
}


// IO policy concept:
//   bool ready(int day);                         // non-blocking: are the inputs of `day` available yet?
//   void forcing(int day, double& P, double& E, double& doy);
//   void upstream(int day, double (&us)[4]);     // area-scaled Qr, Msus, TDP, PP of the parents, summed
//   void emit(int day, const double (&y)[NL], double Vr, const double (&acc)[NA], const double (&non)[13],
//             const Cold& c);                    // y holds the raw end-of-day ODE states
//   void publish(int day);                       // make `day`'s outputs visible to downstream reaches
//
// Reach routing (model.py:508-544) rides on ready()/publish(): a lane whose upstream reaches have not
// finished the day it wants to start simply polls once per loop iteration while the other lanes of its
// warp keep stepping, so upstream and downstream reaches advance as a day-skewed wavefront inside one
// launch and nobody ever blocks a warp-mate.
template <class IO, class KS>
SP_HD void run_member_sc(const double* mp, const double* sp, double A_qr0, int nc_last,
                         const ThreadOptions& opt, int n_days, Cold& c, IO& io, KS& ks, ThreadCounters& cnt) {
  Hot h;
  Flags fl;
  RK rk;
  DayAux aux;
  double y[NL], acc[NA];
  double Kf;
  setup_thread(mp, sp, A_qr0, nc_last, opt.strict_quirks, opt.run_mode_cal, h, c, fl, y, Kf);
  cnt.steps = cnt.rejected = cnt.rhs_evals = 0;
  cnt.status = 0;
  if (n_days <= 0) return;

  const double T = opt.step_len;
  int day = 0;
  double t = 0.0;
  double hstep = 0.05 * T;   // first guess; the controller takes over after the first attempt
  int day_steps = 0;
  bool grow_ok = true;
  int begin = 1;             // the current day has not been started yet
  int alive = 1;             // structured exit: no break/continue, so the warp re-converges every iteration

  while (alive) {
    // `begin` is laundered through an empty asm so that the compiler cannot thread the jump from the
    // day-start block straight into the step body: that would create two copies of the step path that
    // never re-converge (measured: +42 % loop iterations per warp).
    asm volatile("" : "+r"(begin));
    if (!begin) {
      const double rem = T - t;
      const bool last = hstep * 1.0000001 >= rem;
      const double hh = last ? rem : hstep;

      double ynew[NL], accnew[NA], k7[NL], a7[NA];
      const double en = dp5_attempt(h, y, acc, rk, hh, opt.rtol, opt.atol, ynew, accnew, k7, a7, ks);
      cnt.steps += 1;
      cnt.rhs_evals += 6;
      day_steps += 1;

      bool accept = en <= 1.0;
      if (!accept && (day_steps >= opt.max_steps_per_day || hh < 1e-12 * T)) {
        accept = true;           // give up on error control for this step: guarantees forward progress
        cnt.status |= 1;
      }
      double fac = step_factor(en);
      if (accept) {
        t += hh;
#pragma unroll
        for (int i = 0; i < NL; ++i) { y[i] = ynew[i]; rk.k1[i] = k7[i]; }
#pragma unroll
        for (int i = 0; i < NA; ++i) { acc[i] = accnew[i]; rk.a1[i] = a7[i]; }
        if (!grow_ok) fac = sp_min(fac, 1.0);                // no growth right after a rejection
        grow_ok = true;
        const double hnew = hh * fac;
        hstep = (last && hnew < hstep) ? hstep : hnew;       // a clamped final step must not shrink h
      } else {
        cnt.rejected += 1;
        hstep = hh * sp_min(fac, 1.0);
        grow_ok = false;
      }

      if (accept && last) {
        // ---- end of a day: post-ODE algebra (:643-724), output -------------------------------
        double non[13];
        double yraw[NL];
#pragma unroll
        for (int i = 0; i < NL; ++i) yraw[i] = y[i];
        end_day(h, c, fl, opt.dynamic_epc0, aux, y, non);
        bool finite = true;
#pragma unroll
        for (int i = 0; i < NL; ++i) finite = finite && (yraw[i] - yraw[i] == 0.0);
        if (!finite) cnt.status |= 2;
        io.emit(day, yraw, io.wants_vr() ? reach_volume(h, yraw[iQr]) : 0.0, acc, non, c);
        io.publish(day);
        ++day;
        begin = 1;
        alive = day < n_days;
      }
    }
    if (begin && alive) {
      // ---- start of a day: pre-ODE algebra (:497-618).  A lane whose upstream reaches have not yet
      // published this day stays in this state and polls again next iteration.
      if (io.ready(day)) {
        double P, E, doy, us[4];
        io.forcing(day, P, E, doy);
        io.upstream(day, us);
        begin_day(mp, sp, c, fl, opt.dynamic_epc0, opt.dynamic_erod, P, E, doy, us, h, aux);
#pragma unroll
        for (int i = 0; i < NA; ++i) acc[i] = 0.0;
        rhs(h, y, rk.k1, rk.a1);
        cnt.rhs_evals += 1;
        t = 0.0;
        day_steps = 0;
        // the forcing jumps at midnight: restart from a fifth of yesterday's last step size
        // (measured: 1.8 -> 0.6 rejected attempts per day, -5 % attempts, same accuracy)
        hstep = sp_min(hstep * 0.2, T);
        begin = 0;
      }
    }
  }
}


}  // namespace simplyp
