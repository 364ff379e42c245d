"""Check: the register variants of the quad kernel (SIMPLYP_QUAD_MINBLOCKS 2/3/4) give bit-identical results."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from simplyp_b200 import model as spm, packing as pk
from simplyp_b200.engine import Engine
eng = Engine(0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
w = bench.build_workload("2004", M)
opt = spm.make_options(w["p_SU"], w["p"], w["dyn"], w["topo"], 1.0, None, None)
args = (eng.to_device(w["forcing"]), eng.to_device(w["member"]), eng.to_device(w["sc"]), w["topo"].parent_offsets,
        w["topo"].parent_ids, eng.to_device(w["obs_m"]), eng.to_device(w["desc"]), opt)
ref = None
for mb in ("2", "3", "4"):
    os.environ["SIMPLYP_QUAD_MINBLOCKS"] = mb
    stats, diag = eng.calibrate(*args)
    torch.cuda.synchronize()
    st, dg = stats.cpu().numpy(), diag.cpu().numpy()
    if ref is None: ref = (st, dg)
    print("minblocks %s: bitwise equal to minblocks 2: stats %s diag %s" %
          (mb, np.array_equal(ref[0], st, equal_nan=True), np.array_equal(ref[1], dg)), flush=True)
