"""Analysis tool: BASELINE configs 3 / 5 at full topology (730 / 120 days, fixtures tests/golden/ref_config{3,5}.npz) on a
host build of the quad program compiled with the given flags: parity against the oracle and attempts per day.

  python scripts/stiff_tol_exp.py 3 "" "-DSP_ROS_RATE0=150 -DSP_ROS_GMAX=30"
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from errnorm_validate import runner_for, GOLDEN  # noqa: E402
from simplyp_b200 import model as spm, packing as pk, synthetic  # noqa: E402
from tests import parity  # noqa: E402
from tests.util import max_rel  # noqa: E402


def main():
    cfg = int(sys.argv[1])
    z = np.load(os.path.join(GOLDEN, "ref_config%d.npz" % cfg))
    w = synthetic.scale_config(cfg, n_days=int(z["n_days"]))
    opt = spm.make_options(w["p_SU"], w["p"], w["dyn"], w["topo"])
    sel = [int(s) for s in z["reaches"]]
    met, p, p_SC, topo = w["met"], w["p"], w["p_SC"], w["topo"]
    nc_types = pk.validate_land_use(p_SC.copy(), p["SC_list"])
    for flags in sys.argv[2:] or [""]:
        run = runner_for(flags)
        out, dg = run(w["forcing"], w["member"][:1], w["sc"], topo.parent_offsets, topo.parent_ids, opt)
        worst, fails = 0.0, []
        for k, s in enumerate(sel):
            SC = topo.sc_ids[s]
            A = float(p_SC.loc["A_catch", SC])
            tc, r = spm.raw_to_frames(out[0, s], met.index, A, p["Msoil_m2"], p["f_TDP"], nc_types[SC], met["D_snow_end"])
            tco, ro = spm.raw_to_frames(z["raw"][k], met.index, A, p["Msoil_m2"], p["f_TDP"], nc_types[SC], met["D_snow_end"])
            try:
                parity.assert_frames_close(tc, r, tco, ro, "reach %d" % s)
            except AssertionError as e:
                fails.append(str(e))
            worst = max(worst, max(max_rel(r[c].to_numpy(), ro[c].to_numpy()) for c in ("Q_cumecs", "SS_mgl", "TDP_mgl", "PP_mgl")))
        qr = max_rel(out[0, :, -1, 5], z["qr_last_day"])
        spd = dg[0, :, 0] / float(z["n_days"])
        print("[%s] config %d: worst flow/conc %.2e, last-day Qr of every reach %.2e, attempts/day mean %.2f max %.1f (sampled max %.1f), "
              "rejected %.1f%%, status %d, fails %s" % (flags, cfg, worst, qr, spd.mean(), spd.max(), spd[sel].max(),
                                                        100.0 * dg[0, :, 1].sum() / dg[0, :, 0].sum(), int(dg[..., 3].max()), fails[:3]), flush=True)


if __name__ == "__main__":
    main()
