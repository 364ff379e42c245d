"""Experiment: cycles per step attempt of ONE warp (single-warp latency T1) and of small ensembles."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from simplyp_b200 import model as spm, packing as pk
from simplyp_b200.engine import Engine
eng = Engine(0)
for M in [int(x) for x in sys.argv[1:]] or [1, 8, 4736, 9472]:
    w = bench.build_workload("2004", max(M, 8))
    opt = spm.make_options(w["p_SU"], w["p"], w["dyn"], w["topo"], 1.0, None, None)
    opt.pilot_days = -1
    d_forc = eng.to_device(w["forcing"]); d_mem = eng.to_device(w["member"][:M]); d_sc = eng.to_device(w["sc"][:M])
    d_obs = eng.to_device(w["obs_m"]); d_desc = eng.to_device(w["desc"])
    V = w["obs_m"].shape[0]
    stats = torch.empty((M, V, pk.NSTAT), dtype=torch.float64, device=eng.device)
    diag = torch.zeros((M, 1, pk.NDIAG), dtype=torch.int64, device=eng.device)
    po, pid = w["topo"].parent_offsets, w["topo"].parent_ids
    for _ in range(3):
        eng.calibrate(d_forc, d_mem, d_sc, po, pid, d_obs, d_desc, opt, stats=stats, diag=diag)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.calibrate(d_forc, d_mem, d_sc, po, pid, d_obs, d_desc, opt, stats=stats, diag=diag)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    st = diag[:, 0, 0].cpu().numpy()
    print("M=%6d  %.3f ms  max member steps %d  -> %.0f cycles per step of the heaviest member (1.965 GHz)" %
          (M, ms, st.max(), ms * 1e-3 * 1.965e9 / st.max()))
