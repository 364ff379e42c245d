"""Experiment: cycles per step attempt of identical warps (one member replicated) at different fillings of the GPU."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from simplyp_b200 import model as spm, packing as pk
from simplyp_b200.engine import Engine
eng = Engine(0)
w = bench.build_workload("2004", 64)
for M in [int(x) for x in sys.argv[1:]] or [8, 32, 64, 1184, 2368, 4736, 7104, 9472, 14208, 18944]:
    opt = spm.make_options(w["p_SU"], w["p"], w["dyn"], w["topo"], 1.0, None, None)
    opt.pilot_days = -1
    mem = np.repeat(w["member"][:1], M, axis=0); sc = np.repeat(w["sc"][:1], M, axis=0)
    d_forc = eng.to_device(w["forcing"]); d_mem = eng.to_device(mem); d_sc = eng.to_device(sc)
    d_obs = eng.to_device(w["obs_m"]); d_desc = eng.to_device(w["desc"])
    V = w["obs_m"].shape[0]
    stats = torch.empty((M, V, pk.NSTAT), dtype=torch.float64, device=eng.device)
    diag = torch.zeros((M, 1, pk.NDIAG), dtype=torch.int64, device=eng.device)
    po, pid = w["topo"].parent_offsets, w["topo"].parent_ids
    for mb in ("2", "3", "4"):
        os.environ["SIMPLYP_QUAD_MINBLOCKS"] = mb
        for _ in range(2):
            eng.calibrate(d_forc, d_mem, d_sc, po, pid, d_obs, d_desc, opt, stats=stats, diag=diag)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            eng.calibrate(d_forc, d_mem, d_sc, po, pid, d_obs, d_desc, opt, stats=stats, diag=diag)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 4
        st = int(diag[0, 0, 0])
        print("M=%6d (%4d blocks, %5d warps) minblocks %s: %.3f ms, %d attempts -> %.0f cycles per attempt (1.965 GHz)" %
              (M, (M + 31) // 32, (M + 7) // 8, mb, ms, st, ms * 1e-3 * 1.965e9 / st), flush=True)
