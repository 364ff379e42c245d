"""Parity of the CUDA path against the oracle port (LSODA at rtol=1e-10, the reference's own solver) on a LARGER
sample of Latin-hypercube members than the committed fixtures hold: worst relative error of every daily flow and
concentration over `n` members x 366 days.  Oracle members run in parallel on the host cores.

    python scripts/validate_ensemble.py [n_members] [seed] [2004|full]
"""
import json, multiprocessing as mp, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
COLS = ("Q_cumecs", "SS_mgl", "TDP_mgl", "PP_mgl", "TP_mgl", "SRP_mgl", "Qr", "Msus_kg/day", "TDP_kg/day", "PP_kg/day")


PERIODS = {"2004": ("2004-01-01", "2004-12-31"), "full": ("1981-01-01", "2010-12-31")}


def _oracle(args):
    i, n, seed, period = args
    sys.path.insert(0, ROOT)
    from oracle import simplyp_oracle as orc
    from simplyp_b200 import ensemble as ens, tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(*PERIODS[period], dynamic="y")
    samples = ens.latin_hypercube(n, seed=seed)
    pi, pLUi, pSCi = ens.apply_member_to_pandas(samples, i, p, p_LU, p_SC)
    _tc, R, _kf, _ = orc.run_simply_p(met, p_struc, p_SU, pLUi, pSCi, pi, dyn, rtol=1e-10, atol=1e-13, mxstep=500000)
    return i, {c: R[1][c].to_numpy() for c in COLS}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 424242
    period = sys.argv[3] if len(sys.argv) > 3 else "2004"
    from simplyp_b200 import _cabi, ensemble as ens, model as spm, packing as pk, tarland
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(*PERIODS[period], dynamic="y")
    topo = pk.build_topology(p_struc, p["SC_list"])
    opt = spm.make_options(p_SU, p, dyn, topo)
    samples = ens.latin_hypercube(n, seed=seed)
    member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo.sc_ids), samples)
    out, dg = _cabi.run_host(pk.forcing_matrix(met), member, sc, topo.parent_offsets, topo.parent_ids, opt)
    t0 = time.time()
    with mp.get_context("spawn").Pool(os.cpu_count() or 1) as pool:
        res = dict(pool.map(_oracle, [(i, n, seed, period) for i in range(n)]))
    worst, where = 0.0, None
    per_member = np.zeros(n)
    for i in range(n):
        _tc, r = spm.raw_to_frames(out[i, 0], met.index, float(sc[i, 0, pk.SC_INDEX["A_catch"]]), p["Msoil_m2"],
                                   float(member[i, pk.MEMBER_INDEX["f_TDP"]]), "None", None)
        for c in COLS:
            a, b = r[c].to_numpy(), res[i][c]
            e = float(np.max(np.abs(a - b) / np.abs(b)))
            per_member[i] = max(per_member[i], e)
            if e > worst:
                worst, where = e, (i, c)
    print(json.dumps({"members": n, "seed": seed, "days": len(met), "worst_rel_err": worst, "where": where,
                      "median_member_worst": float(np.median(per_member)), "p95_member_worst": float(np.percentile(per_member, 95)),
                      "status_bits": int(dg[..., 3].max()), "oracle_seconds": time.time() - t0,
                      "rtol": opt.rtol, "atol": opt.atol}), flush=True)


if __name__ == "__main__":
    main()
