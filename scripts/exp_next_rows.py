"""Timing of the "next" rows on one GPU with CUDA events:
  1. calibration with rank statistics (Spearman's r) vs without, bench ensemble (10^4 members, 2004) and the
     30-year record (4,645 daily Q observations);
  2. sum_to_waterbody reduction on a full-output run (HBM-bound: algorithmic bytes = 4 doubles read per flagged
     reach-day + 11 doubles written per member-day);
  3. per-member snow on the device vs host pre-processing (same kernel, one extra recursion per day).
"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from simplyp_b200 import _cabi, model as spm, packing as pk
from simplyp_b200.engine import Engine

eng = Engine(0)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for period, M in (("2004", 10000), ("full", 2000)):
    w = bench.build_workload(period, M)
    d = {k: eng.to_device(w[k]) for k in ("forcing", "member", "sc", "obs_m", "desc")}
    po, pid = w["topo"].parent_offsets, w["topo"].parent_ids
    res = {}
    for ranks in (0, 1):
        opt = spm.make_options(w["p_SU"], w["p"], w["dyn"], w["topo"])
        opt.rank_stats = ranks
        res[ranks] = timed(lambda: eng.calibrate(d["forcing"], d["member"], d["sc"], po, pid, d["obs_m"], d["desc"], opt))
    n_obs = int(np.sum(~np.isnan(w["obs_m"]), axis=1).max())
    print(json.dumps({"row": "rank statistics", "period": period, "members": M, "days": int(w["forcing"].shape[0]),
                      "max_obs_per_series": n_obs, "ms_without": res[0], "ms_with": res[1],
                      "rank_stats_overhead_ms": res[1] - res[0]}), flush=True)

# waterbody sums on a 64-reach network, 8 members, 10 years
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from simplyp_b200 import synthetic, tarland, inputs as spi
p_SU, dyn, p, p_LU, p_SC0, p_struc0, met, obs = tarland.load(dynamic="y")
p, p_SC, p_struc = synthetic.random_network(p, p_SC0[1], n_sc=64, seed=3)
met = spi.snow_hydrol_inputs(p["D_snow_0"], p["f_DDSM"], synthetic.synthetic_met(3652, seed=11))
pk.validate_land_use(p_SC, p["SC_list"])
topo = pk.build_topology(p_struc, p["SC_list"])
opt = spm.make_options(p_SU, p, dyn, topo)
M = 64
member = np.repeat(pk.member_vector(p, p_LU)[None], M, axis=0)
member[:, pk.MEMBER_INDEX["a_Q"]] *= np.linspace(0.8, 1.2, M)
d_mem, d_sc = eng.to_device(member), eng.to_device(pk.sc_matrix(p_SC, topo.sc_ids)[None])
out, _ = eng.run(eng.to_device(pk.forcing_matrix(met)), d_mem, d_sc, topo.parent_offsets, topo.parent_ids, opt)
torch.cuda.synchronize()
reaches = list(range(topo.n_sc - 16, topo.n_sc))
ms = timed(lambda: eng.sum_to_waterbody(out, d_mem, d_sc, reaches), reps=10)
D = out.shape[2]
alg = M * D * (len(reaches) * 4 * 8 + 11 * 8)
sect = M * D * (len(reaches) * 64 + 96)     # 32-byte sectors actually touched: columns 5..11 of a row span 2 sectors
print(json.dumps({"row": "sum_to_waterbody", "members": M, "days": D, "reaches_summed": len(reaches), "ms": ms,
                  "algorithmic_GB": alg / 1e9, "algorithmic_GBps": alg / 1e9 / (ms * 1e-3),
                  "sector_GBps": sect / 1e9 / (ms * 1e-3), "hbm_peak_GBps": peaks.get("hbm_gbs"),
                  "frac_of_measured_peak_sectors": sect / 1e9 / (ms * 1e-3) / peaks.get("hbm_gbs", 6545.3)}), flush=True)

# snow on the device
w = bench.build_workload("2004", 10000)
opt = spm.make_options(w["p_SU"], w["p"], w["dyn"], w["topo"])
d = {k: eng.to_device(w[k]) for k in ("forcing", "member", "sc", "obs_m", "desc")}
po, pid = w["topo"].parent_offsets, w["topo"].parent_ids
t_host = timed(lambda: eng.calibrate(d["forcing"], d["member"], d["sc"], po, pid, d["obs_m"], d["desc"], opt))
raw = eng.to_device(pk.forcing_matrix(w["met"], raw_snow=True))
opt.snow_on_device = 1
t_dev = timed(lambda: eng.calibrate(raw, d["member"], d["sc"], po, pid, d["obs_m"], d["desc"], opt))
print(json.dumps({"row": "snow on device", "members": 10000, "ms_host_preprocessed": t_host, "ms_snow_on_device": t_dev}), flush=True)
