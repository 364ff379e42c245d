"""Experiment: where and when every warp of the bench kernel ran (build with -DSP_TIMELINE).
  nvcc ... -DSP_TIMELINE -o build/exp/libsimplyp_timeline.so simplyp_b200/csrc/simplyp_kernels.cu
  SIMPLYP_B200_LIB=build/exp/libsimplyp_timeline.so python scripts/exp_timeline.py 10000
Writes gpurun_out/timeline_<M>[_solo<K>].npz: per member steps, start/end (ns), SM, virtual block, warp."""
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
d_forc = eng.to_device(w["forcing"]); d_mem = eng.to_device(w["member"][:M]); d_sc = eng.to_device(w["sc"][:M])
d_obs = eng.to_device(w["obs_m"]); d_desc = eng.to_device(w["desc"])
V = w["obs_m"].shape[0]
po, pid = w["topo"].parent_offsets, w["topo"].parent_ids
stats = torch.empty((M, V, pk.NSTAT), dtype=torch.float64, device=eng.device)
diag = torch.zeros((M, 1, pk.NDIAG), dtype=torch.int64, device=eng.device)
for _ in range(3):
    eng.calibrate(d_forc, d_mem, d_sc, po, pid, d_obs, d_desc, opt, stats=stats, diag=diag)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.calibrate(d_forc, d_mem, d_sc, po, pid, d_obs, d_desc, opt, stats=stats, diag=diag)
e1.record(); torch.cuda.synchronize()
dg = diag.cpu().numpy()[:, 0, :]
tag = os.environ.get("SIMPLYP_SOLO_WARPS")
out = os.path.join(ROOT, "gpurun_out", "timeline_%d%s.npz" % (M, "_solo" + tag if tag else ""))
os.makedirs(os.path.dirname(out), exist_ok=True)
np.savez_compressed(out, steps=dg[:, 0], t0=dg[:, 1], t1=dg[:, 2], smid=dg[:, 3] & 0xffff, vblock=(dg[:, 3] >> 16) & 0xffffff,
                    warp=(dg[:, 3] >> 40) & 0xf, lockstep=dg[:, 3] >> 44, ms=e0.elapsed_time(e1))
print("M=%d: %.3f ms; kernel span %.3f ms" % (M, e0.elapsed_time(e1), (dg[:, 2].max() - dg[:, 1].min()) * 1e-6))
