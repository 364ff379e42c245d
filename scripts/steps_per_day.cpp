// Analysis tool (not product, not test): per-member per-day step-attempt counts of the scalar program,
// used to compare warp scheduling policies (flattened vs day-lock-step, sorted vs unsorted members).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include "../tests/hostemu/scalar_program.h"
using namespace simplyp;
struct IO {
  const double* f; const ThreadCounters* cnt; uint16_t* steps; long long last;
  void forcing(int day, double& P, double& E, double& doy) const { P = f[4*day]; E = f[4*day+1]; doy = f[4*day+2]; }
  void upstream(int, double (&us)[4]) const { us[0]=us[1]=us[2]=us[3]=0; }
  bool wants_vr() const { return false; }
  bool ready(int) const { return true; }
  void publish(int) const {}
  void emit(int day, const double (&)[NL], double, const double (&)[NA], const double (&)[13], const Cold&) {
    steps[day] = (uint16_t)(cnt->steps - last); last = cnt->steps;
  }
};
extern "C" int steps_per_day(int M, int D, const double* forcing, const double* mp, const double* scp,
                             double rtol, double atol, uint16_t* steps) {
  ThreadOptions t; t.rtol = rtol; t.atol = atol; t.step_len = 1.0; t.max_steps_per_day = 5000;
  t.snow_on_device = 0; t.dynamic_epc0 = 1; t.dynamic_erod = 1; t.run_mode_cal = 1; t.strict_quirks = 1;
#pragma omp parallel for schedule(dynamic, 16)
  for (int m = 0; m < M; ++m) {
    ThreadCounters cnt; Cold c; RegStages ks;
    IO io{forcing, &cnt, steps + (size_t)m * D, 0};
    run_member_sc(mp + (size_t)m * SIMPLYP_NP_MEMBER, scp, scp[SIMPLYP_SC_A_CATCH], 0, t, D, c, io, ks, cnt);
  }
  return 0;
}
