// TEST HARNESS ONLY — compiles simplyp_core.cuh / simplyp_thread.cuh for the host so that the
// CPU-only test tier (`pytest -m "not gpu"`) can exercise the same per-thread arithmetic and
// control flow the CUDA kernels run, against the oracle, in a container without a GPU.
// Nothing in simplyp_b200/ loads this library; the product has no CPU execution path.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../simplyp_b200/csrc/simplyp_thread.cuh"

using namespace simplyp;

namespace {
struct HostIO {
  const double* fdata; const double* scp; const int32_t* po; const int32_t* pid;
  double* out; int S, D, m, s;

  void forcing(int day, double& P, double& E, double& doy) const {
    P = fdata[4 * day]; E = fdata[4 * day + 1]; doy = fdata[4 * day + 2];
  }
  void upstream(int day, double (&us)[4]) const {
    us[0] = us[1] = us[2] = us[3] = 0.0;
    const double A_this = scp[(size_t)s * SIMPLYP_NP_SC + SIMPLYP_SC_A_CATCH];
    for (int e = po[s]; e < po[s + 1]; ++e) {
      const int p = pid[e];
      const double* row = out + (((size_t)m * S + p) * D + day) * SIMPLYP_NOUT;
      const double A_up = scp[(size_t)p * SIMPLYP_NP_SC + SIMPLYP_SC_A_CATCH];
      us[0] += row[SIMPLYP_O_QR] * (A_up / A_this);
      us[1] += row[SIMPLYP_O_MSUS_FLUX];
      us[2] += row[SIMPLYP_O_TDP_FLUX];
      us[3] += row[SIMPLYP_O_PP_FLUX];
    }
  }
  bool wants_vr() const { return true; }
  bool ready(int) const { return true; }
  void publish(int) const {}
  void emit(int day, const double (&y)[NL], double Vr, const double (&acc)[NA], const double (&non)[13],
            const Cold&) const {
    double* row = out + (((size_t)m * S + s) * D + day) * SIMPLYP_NOUT;
    row[0] = y[iVsA]; row[1] = y[iVsS]; row[2] = y[iVg]; row[3] = Vr; row[4] = y[iQr]; row[5] = acc[0];
    row[6] = y[iMsus]; row[7] = acc[1]; row[8] = y[iTDPr]; row[9] = acc[2]; row[10] = y[iPPr]; row[11] = acc[3];
    for (int i = 0; i < 13; ++i) row[12 + i] = non[i];
  }
};
}  // namespace

extern "C" int hostemu_run(const SimplypDims* dims, const SimplypOptions* opt, const double* forcing,
                           const double* member_params, const double* sc_params, const int32_t* po,
                           const int32_t* pid, double* out, int64_t* diag) {
  const int M = dims->n_members, S = dims->n_sc, D = dims->n_days, Msc = dims->n_sc_param_sets;
  ThreadOptions t;
  t.rtol = opt->rtol; t.atol = opt->atol; t.step_len = opt->step_len;
  t.max_steps_per_day = opt->max_steps_per_day > 0 ? opt->max_steps_per_day : 5000;
  t.dynamic_epc0 = opt->dynamic_epc0; t.dynamic_erod = opt->dynamic_erodibility;
  t.run_mode_cal = opt->run_mode_cal; t.strict_quirks = opt->strict_quirks;
  for (int m = 0; m < M; ++m) {
    const double* mp = member_params + (size_t)m * SIMPLYP_NP_MEMBER;
    const double* scp = sc_params + (size_t)(Msc > 1 ? m : 0) * S * SIMPLYP_NP_SC;
    const double* spl = scp + (size_t)(S - 1) * SIMPLYP_NP_SC;
    const double fNCA_last = spl[SIMPLYP_SC_F_AR] * spl[SIMPLYP_SC_F_NC_AR] + spl[SIMPLYP_SC_F_NC_IG] * spl[SIMPLYP_SC_F_IG];
    const int nc_last = fNCA_last > 0.0 ? 1 : (spl[SIMPLYP_SC_F_NC_S] > 0.0 ? 2 : 0);
    const double A_qr0 = scp[(size_t)opt->sc_qr0 * SIMPLYP_NP_SC + SIMPLYP_SC_A_CATCH];
    for (int s = 0; s < S; ++s) {   // run order is upstream-first
      HostIO io{forcing, scp, po, pid, out, S, D, m, s};
      Cold c;
      ThreadCounters cnt;
      RegStages ks;
      run_member_sc(mp, scp + (size_t)s * SIMPLYP_NP_SC, A_qr0, nc_last, t, D, c, io, ks, cnt);
      if (diag) {
        int64_t* dg = diag + ((size_t)m * S + s) * SIMPLYP_NDIAG;
        dg[0] = cnt.steps; dg[1] = cnt.rejected; dg[2] = cnt.rhs_evals; dg[3] = cnt.status;
      }
    }
  }
  return 0;
}
