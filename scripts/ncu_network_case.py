"""One full-output launch of a network configuration, for ncu application replay (the network kernel waits on
progress flags written by other blocks of the same launch, so kernel replay cannot profile it).

    python scripts/ncu_network_case.py <cfg 3|5> <members> <days>
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from simplyp_b200 import packing as pk, synthetic
from simplyp_b200.engine import Engine

cfg, M, D = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
w = synthetic.scale_config(cfg, M, n_days=D)
eng = Engine(0)
topo = w["topo"]
d_f, d_m, d_s = eng.to_device(w["forcing"]), eng.to_device(w["member"]), eng.to_device(w["sc"])
out = torch.empty((M, topo.n_sc, D, pk.NOUT), dtype=torch.float64, device=eng.device)
diag = torch.zeros((M, topo.n_sc, pk.NDIAG), dtype=torch.int64, device=eng.device)
for _ in range(2):
    eng.run(d_f, d_m, d_s, topo.parent_offsets, topo.parent_ids, w["opt"], out=out, diag=diag)
torch.cuda.synchronize()
print("ok", int(diag[..., 3].max().item()), float(out[0, -1, -1, 5].item()))
