"""Drop-in ``run_simply_p`` on top of the CUDA integrator.

Same signature, positional order, return tuple, side effects and exceptions as the reference's
``run_simply_p`` (``simplyP/model.py:193-827``); the sub-catchment x day loop (``:365-724``) runs in
``libsimplyp_b200.so``.  What stays on the host is exactly what the reference does outside the loops:
validation and derived rows (``:311-335``), DataFrame assembly and unit conversions (``:736-800``),
the optional CSV dump (``:815-825``).
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd

from . import _cabi
from . import helper_functions as hf
from . import packing as pk

#: default tolerances of the embedded RK integrator.  The reference hard-codes LSODA rtol=0.01
#: (``model.py:640``), which leaves its own output 2e-3..6e-3 away from the converged solution
#:; these defaults keep every daily flow/concentration within 1e-5 (relative)
#: of the converged solution instead.
DEFAULT_RTOL = 1e-7
DEFAULT_ATOL = 1e-10


def make_options(p_SU, p, dynamic_options, topology, step_len=1.0, rtol=None, atol=None,
                 strict_reference_quirks=True):
    """SimplypOptions for a run (tolerances, dynamic options, run mode, the reach that sets the initial flow)."""
    sc_ids = topology.sc_ids
    qr0 = int(p["SC_Qr0"])
    if qr0 not in sc_ids:
        raise KeyError(qr0)
    return _cabi.default_options(
        rtol=DEFAULT_RTOL if rtol is None else float(rtol),
        atol=DEFAULT_ATOL if atol is None else float(atol),
        step_len=float(step_len),
        dynamic_epc0=1 if dynamic_options["Dynamic_EPC0"] == "y" else 0,
        dynamic_erodibility=1 if dynamic_options["Dynamic_erodibility"] == "y" else 0,
        run_mode_cal=1 if p_SU["run_mode"] == "cal" else 0,
        sc_qr0=sc_ids.index(qr0),
        strict_quirks=1 if strict_reference_quirks else 0,
    )


def _prepare_inputs(p_struc, p_LU, p_SC, p):
    """Reference ``model.py:311-361`` (validation, derived rows, in-place mutation of the caller's frames)."""
    p_LU.loc["EPC0_0", :] = 4 * [np.nan]
    p_LU.loc["Plab0", :] = 4 * [np.nan]
    p_LU.loc["TDPs0", :] = 4 * [np.nan]
    p_SC.loc["f_A"] = p_SC.loc["f_IG"] + p_SC.loc["f_Ar"]
    p_SC.loc["f_NC_A"] = (p_SC.loc["f_Ar"] * p_SC.loc["f_NC_Ar"]) + (p_SC.loc["f_NC_IG"] * p_SC.loc["f_IG"])
    nc_types = pk.validate_land_use(p_SC, p["SC_list"])
    if "NC_type" not in p_SC.index:
        # the row holds strings: make the frame object-typed first so pandas does not refuse the assignment
        for col in p_SC.columns:
            p_SC[col] = p_SC[col].astype(object)
    for SC, nc in nc_types.items():
        p_SC.loc["NC_type", SC] = nc
    pk.check_erosion_windows(p)
    topology = pk.build_topology(p_struc, p["SC_list"])
    return topology, nc_types


def _finish_mutations(p_LU, p_SC, p, sc_ids):
    """Values the reference leaves behind in the caller's frames: ``EPC0_0``/``Plab0``/``TDPs0`` of the LAST
    sub-catchment (``model.py:409-422``) and ``TDPeff`` NaN -> 0 (``:462-463``)."""
    last = sc_ids[-1]
    A = float(p_SC.loc["A_catch", last])
    Msoil = p["Msoil_m2"] * 10 ** 6 * A
    for LU in ["A", "S"]:
        p_LU.loc["EPC0_0", LU] = hf.UC_Cinv(p_LU[LU]["EPC0_init_mgl"], A)
        p_LU.loc["Plab0", LU] = 10 ** -6 * (p_LU[LU]["SoilPconc"] - p_LU["S"]["SoilPconc"]) * Msoil
        p_LU.loc["TDPs0", LU] = p_LU[LU]["EPC0_0"] * p["fc"] if LU == "A" else 0
    for SC in sc_ids:
        v = p_SC.loc["TDPeff", SC]
        if isinstance(v, float) and np.isnan(v):
            p_SC.loc["TDPeff", SC] = 0.


def raw_to_frames(raw_sc, index, A_catch, Msoil_m2, f_TDP, nc_type, D_snow_end=None):
    """One sub-catchment's raw [D][25] block -> (df_TC, df_R) exactly as reference ``model.py:736-800``."""
    df_ODE = pd.DataFrame(raw_sc[:, :12], columns=pk.ODE_COLS, index=index)
    df_nonODE = pd.DataFrame(raw_sc[:, 12:], columns=pk.NONODE_COLS, index=index)
    df_TC = pd.concat([df_ODE[["VsA", "VsS", "Vg"]], df_nonODE], axis=1)
    df_TC["TDPs_A_mgl"] = hf.UC_C(df_TC["conc_TDPs_A_kgmm"], A_catch)
    df_TC["EPC0_A_mgl"] = hf.UC_C(df_TC["EPC0_A_kgmm"], A_catch)
    df_TC["Plabile_A_mgkg"] = (10 ** 6 * df_TC["P_labile_A_kg"] / (Msoil_m2 * 10 ** 6 * A_catch))
    if nc_type != "None":
        src = "A" if nc_type == "A" else "S"
        df_TC["VsNC"] = df_TC["Vs" + src]
        df_TC["QsNC"] = df_TC["Qs" + src]
        df_TC["TDPs_NC_mgl"] = hf.UC_C(df_TC["conc_TDPs_NC_kgmm"], A_catch)
        df_TC["Plabile_NC_mgkg"] = (10 ** 6 * df_TC["P_labile_NC_kg"] / (Msoil_m2 * 10 ** 6 * A_catch))
    if D_snow_end is not None:
        df_TC["D_snow"] = D_snow_end
    df_R = df_ODE.drop(["VsA", "VsS", "Vg"], axis=1)
    df_R["Q_cumecs"] = df_R["Qr"] * A_catch * 1000 / 86400
    df_R["SS_mgl"] = hf.UC_C(df_R["Msus_kg/day"] / df_R["Qr"], A_catch)
    df_R["TDP_mgl"] = hf.UC_C(df_R["TDP_kg/day"] / df_R["Qr"], A_catch)
    df_R["PP_mgl"] = hf.UC_C(df_R["PP_kg/day"] / df_R["Qr"], A_catch)
    df_R = derived_P_species(df_R, f_TDP)
    return df_TC.sort_index(axis=1), df_R.sort_index(axis=1)


def run_simply_p(met_df, p_struc, p_SU, p_LU, p_SC, p, dynamic_options, step_len=1., *, rtol=None, atol=None,
                 device=0, strict_reference_quirks=True, verbose=True):
    """Run SimplyP for one parameter set on the GPU and return the reference's 4-tuple
    ``(df_TC_dict, df_R_dict, Kf, output_dict)``.

    ``output_dict`` holds the integrator diagnostics that replace LSODA's info dict: ``nst`` (step
    attempts), ``nfe`` (RHS evaluations), ``nrej`` and ``message`` for the last sub-catchment, plus
    ``per_sc`` with the same counters for every sub-catchment.
    """
    topology, nc_types = _prepare_inputs(p_struc, p_LU, p_SC, p)
    sc_ids = topology.sc_ids
    opt = make_options(p_SU, p, dynamic_options, topology, step_len, rtol, atol, strict_reference_quirks)

    forcing = pk.forcing_matrix(met_df)
    member = pk.member_vector(p, p_LU)[None, :]
    scp = pk.sc_matrix(p_SC, sc_ids)[None, :, :]
    if verbose:
        for SC in sc_ids:
            print("Starting model run for sub-catchment: %s" % SC)
            ups = topology.upstream[SC]
            if ups:
                print("Reaches directly upstream of this reach: %s" % ups)
            else:
                print("No reaches directly upstream of this reach")
    out, diag = _cabi.run_host(forcing, member, scp, topology.parent_offsets, topology.parent_ids, opt, device=device)
    if verbose:
        print("Finished!\n")
    status = int(np.bitwise_or.reduce(diag[0, :, 3]))
    if status & 2:
        import warnings
        warnings.warn("simplyp_b200: non-finite state encountered during integration", RuntimeWarning)
    if status & 1:
        import warnings
        warnings.warn("simplyp_b200: max_steps_per_day reached; error control was relaxed on some days",
                      RuntimeWarning)

    _finish_mutations(p_LU, p_SC, p, sc_ids)

    df_TC_dict, df_R_dict = {}, {}
    snow = met_df["D_snow_end"] if p_SU["inc_snowmelt"] == "y" else None
    Kf = None
    for i, SC in enumerate(sc_ids):
        A = float(p_SC.loc["A_catch", SC])
        df_TC, df_R = raw_to_frames(out[0, i], met_df.index, A, p["Msoil_m2"], p["f_TDP"], nc_types[SC], snow)
        df_TC_dict[SC] = df_TC
        df_R_dict[SC] = df_R
        if p_SU["run_mode"] == "cal":      # model.py:449-453, per SC; the last one is returned (:827)
            Kf = 10 ** -6 * (p_LU["A"]["SoilPconc"] - p_LU["S"]["SoilPconc"]) / hf.UC_Cinv(p_LU["A"]["EPC0_init_mgl"], A)
        else:
            Kf = p["Kf"]

    if verbose:
        if p_SU["run_mode"] == "cal":
            print("Running in calibration mode; the soil P sorption coefficient has been estimated as %s mm/kg\n" % Kf)
        else:
            print("Running in validation or scenario mode, so the soil P sorption coefficient has been read "
                  "from the parameter file")

    if p_SU["save_output_csvs"] == "y":    # model.py:815-825
        for SC in df_R_dict.keys():
            df_TC_dict[SC].to_csv(os.path.join(p_SU["output_fpath"], "Results_TC_SC%s.csv" % SC))
            df_R_toSave = df_R_dict[SC].drop(["Msus_EndOfDay", "PPr_EndOfDay", "Qr", "Qr_EndOfDay",
                                              "TDPr_EndOfDay", "Vr"], axis=1)
            df_R_toSave.to_csv(os.path.join(p_SU["output_fpath"], "Instream_results_Reach%s.csv" % SC))
        if verbose:
            print("Results saved to csv\n")

    last = len(sc_ids) - 1
    output_dict = {
        "nst": int(diag[0, last, 0]), "nrej": int(diag[0, last, 1]), "nfe": int(diag[0, last, 2]),
        "status": int(diag[0, last, 3]),
        "message": "Integration successful." if status == 0 else "Integration finished with status bits %d" % status,
        "rtol": opt.rtol, "atol": opt.atol, "method": "Tsitouras 5(4) embedded Runge-Kutta, quad kernel, sm_100a",
        "per_sc": {SC: {"nst": int(diag[0, i, 0]), "nrej": int(diag[0, i, 1]), "nfe": int(diag[0, i, 2]),
                        "status": int(diag[0, i, 3])} for i, SC in enumerate(sc_ids)},
    }
    return (df_TC_dict, df_R_dict, Kf, output_dict)


def derived_P_species(df_R, f_TDP):
    """TP = TDP + PP and SRP = f_TDP * TDP for fluxes and concentrations (reference ``model.py:831-847``)."""
    df_R["TP_mgl"] = df_R["TDP_mgl"] + df_R["PP_mgl"]
    df_R["TP_kg/day"] = df_R["TDP_kg/day"] + df_R["PP_kg/day"]
    df_R["SRP_mgl"] = df_R["TDP_mgl"] * f_TDP
    df_R["SRP_kg/day"] = df_R["TDP_kg/day"] * f_TDP
    return df_R


def sum_to_waterbody(p_struc, n_SC, df_R_dict, f_TDP):
    """Sum the reaches flagged ``In_final_flux? == 1`` into one series (reference ``model.py:851-900``)."""
    vars_to_sum = ["Q_cumecs", "Msus_kg/day", "TDP_kg/day", "PP_kg/day"]
    reaches = p_struc["In_final_flux?"][p_struc["In_final_flux?"] == 1].index.values
    if len(reaches) > n_SC:
        raise ValueError("Mismatch between the number of subcatchments in the 'Setup' parameter sheet \n"
                         "(parameter 'n_SC') and in the 'Reach_structure' parameter sheet")
    print("Sub-catchments flowing directly into receiving waterbody: %s" % reaches)
    if len(reaches) > 1:
        index = df_R_dict[reaches[0]].index
        df_summed = pd.DataFrame({var: np.sum([df_R_dict[r][var].to_numpy() for r in reaches], axis=0)
                                  for var in vars_to_sum}, index=index, columns=vars_to_sum)
        for conc, flux in (("SS_mgl", "Msus_kg/day"), ("TDP_mgl", "TDP_kg/day"), ("PP_mgl", "PP_kg/day")):
            df_summed[conc] = (df_summed[flux] / df_summed["Q_cumecs"]) * (1000. / 86400.)
        return derived_P_species(df_summed, f_TDP)
    print("One or fewer reaches were selected to be included in the sum, check your reach structure parameters")
    return None
