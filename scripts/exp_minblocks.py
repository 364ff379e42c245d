"""Experiment: register variant of the quad kernel (SIMPLYP_QUAD_MINBLOCKS = 2/3/4) at large ensemble sizes."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from simplyp_b200 import model as spm, packing as pk
from simplyp_b200.engine import Engine
eng = Engine(0)
for M in [int(x) for x in sys.argv[1:]] or [20000, 40000, 125000]:
    w = bench.build_workload("2004", M)
    opt = spm.make_options(w["p_SU"], w["p"], w["dyn"], w["topo"], 1.0, None, None)
    d_forc = eng.to_device(w["forcing"]); d_mem = eng.to_device(w["member"][:M]); d_sc = eng.to_device(w["sc"][:M])
    d_obs = eng.to_device(w["obs_m"]); d_desc = eng.to_device(w["desc"])
    V = w["obs_m"].shape[0]
    po, pid = w["topo"].parent_offsets, w["topo"].parent_ids
    stats = torch.empty((M, V, pk.NSTAT), dtype=torch.float64, device=eng.device)
    diag = torch.zeros((M, 1, pk.NDIAG), dtype=torch.int64, device=eng.device)
    for mb in ("default", "2", "3", "4"):
        if mb == "default": os.environ.pop("SIMPLYP_QUAD_MINBLOCKS", None)
        else: os.environ["SIMPLYP_QUAD_MINBLOCKS"] = mb
        for _ in range(2):
            eng.calibrate(d_forc, d_mem, d_sc, po, pid, d_obs, d_desc, opt, stats=stats, diag=diag)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            eng.calibrate(d_forc, d_mem, d_sc, po, pid, d_obs, d_desc, opt, stats=stats, diag=diag)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print("M=%6d minblocks %-7s %.2f ms  %.3e member-SC-days/s" % (M, mb, ms, M * 366 / (ms * 1e-3)), flush=True)
