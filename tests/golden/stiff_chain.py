"""Three-reach chain whose outlet reach is tiny: it carries the runoff of 1,500 km2 over `area` km2 of its own, so its
flow per unit own area — and with it the reach rate constant of ode_f (model.py:127-130) — is huge.  This is the
situation of a main-stem reach in a large network, where the reference's LSODA switches to BDF."""
import numpy as np
import pandas as pd


def stiff_chain_inputs(p, p_SC_col, area=0.05):
    cols = {}
    for i, (A, L) in enumerate(((600.0, 12000.0), (900.0, 9000.0), (float(area), 3000.0)), start=1):
        c = p_SC_col.copy()
        c["A_catch"], c["L_reach"] = A, L
        cols[i] = c
    p_SC = pd.DataFrame(cols)
    p_struc = pd.DataFrame({"Upstream_SCs": pd.Series([np.nan, np.nan, "1, 2"], index=[1, 2, 3], dtype=object),
                            "In_final_flux?": [np.nan, np.nan, 1.0]}, index=pd.Index([1, 2, 3], name="Reach"))
    p = p.copy(deep=True)
    p["SC_list"] = np.arange(1, 4)
    p["SC_Qr0"] = 1
    return p, p_SC, p_struc
