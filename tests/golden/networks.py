"""Network set-ups shared by the fixture generator and the tests."""
import numpy as np
import pandas as pd


def network5_inputs(p, p_LU, p_SC, p_struc):
    """A 5-reach branching network built from the Tarland sheet: 1,2 -> 3 ; 3,4 -> 5, mixed NC land."""
    p = p.copy(deep=True)
    p["SC_list"] = np.arange(1, 6)
    p_LU = p_LU.copy(deep=True)
    base = p_SC[1]
    cols = {}
    spec = {
        1: dict(A_catch=51.7, f_Ar=0.2, f_IG=0.3, f_S=0.5, f_NC_Ar=0.25, f_NC_IG=0.0, f_NC_S=0.0, L_reach=10000.0),
        2: dict(A_catch=23.0, f_Ar=0.1, f_IG=0.15, f_S=0.75, f_NC_Ar=0.0, f_NC_IG=0.0, f_NC_S=0.2, L_reach=6000.0,
                TDPeff=np.nan),
        3: dict(A_catch=34.5, f_Ar=0.45, f_IG=0.3, f_S=0.25, f_NC_Ar=0.0, f_NC_IG=0.0, f_NC_S=0.0, L_reach=8000.0,
                S_Ar=6.0, f_spr=0.3),
        4: dict(A_catch=12.0, f_Ar=0.0, f_IG=0.25, f_S=0.75, f_NC_Ar=0.0, f_NC_IG=0.5, f_NC_S=0.0, L_reach=4000.0,
                TDPeff=0.02),
        5: dict(A_catch=60.25, f_Ar=0.3, f_IG=0.2, f_S=0.5, f_NC_Ar=0.0, f_NC_IG=0.0, f_NC_S=0.1, L_reach=15000.0,
                S_reach=0.5),
    }
    for sc, over in spec.items():
        col = base.copy()
        for k, v in over.items():
            col[k] = v
        cols[sc] = col
    p_SC = pd.DataFrame(cols)
    p_struc = pd.DataFrame({"Upstream_SCs": [np.nan, np.nan, "1, 2", np.nan, "3, 4"],
                            "In_final_flux?": [np.nan, np.nan, np.nan, np.nan, 1.0]},
                           index=pd.Index([1, 2, 3, 4, 5], name="Reach"))
    p_struc["Upstream_SCs"] = p_struc["Upstream_SCs"].astype(object)
    return p, p_LU, p_SC, p_struc
