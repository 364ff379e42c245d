"""Experiment: epoch length of the network sweep (SIMPLYP_EPOCH_DAYS; 0 = one piece) — or any other environment
switch of the library, given as KEY=VALUE — on BASELINE configs 3 and 5.
    python scripts/exp_epochs.py 3:64 5:8 -- 0 256 512 1024
    python scripts/exp_epochs.py 3:64 5:8 -- SIMPLYP_NET_MINBLOCKS=2 SIMPLYP_NET_MINBLOCKS=3"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from simplyp_b200 import packing as pk, synthetic
from simplyp_b200.engine import Engine
args = sys.argv[1:]
cut = args.index("--") if "--" in args else len(args)
cases = [tuple(int(x) for x in a.split(":")) for a in args[:cut]] or [(3, 64), (5, 8)]
lengths = args[cut + 1:] or ["0", "512"]
eng = Engine(0)
for cfg, M in cases:
    w = synthetic.scale_config(cfg, M)
    topo, opt = w["topo"], w["opt"]
    S, D = topo.n_sc, w["forcing"].shape[0]
    d_f, d_m, d_s = eng.to_device(w["forcing"]), eng.to_device(w["member"]), eng.to_device(w["sc"])
    out = torch.empty((M, S, D, pk.NOUT), dtype=torch.float64, device=eng.device)
    diag = torch.zeros((M, S, pk.NDIAG), dtype=torch.int64, device=eng.device)
    po, pid = topo.parent_offsets, topo.parent_ids
    ref = None
    for E in lengths:
        if "=" in E:
            os.environ[E.split("=")[0]] = E.split("=")[1]
        else:
            os.environ["SIMPLYP_EPOCH_DAYS"] = E
        eng.run(d_f, d_m, d_s, po, pid, opt, out=out, diag=diag)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            eng.run(d_f, d_m, d_s, po, pid, opt, out=out, diag=diag)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        chk = (float(out[:, :, ::97, 5].sum().item()), int(diag[..., 0].sum().item()), int(diag[..., 3].max().item()))
        same = "" if ref is None else (" same checksum" if chk == ref else " DIFFERENT %r vs %r" % (chk, ref))
        ref = ref or chk
        print("config %d, %d members, setting %-26s: %8.1f ms  %.3e member-SC-days/s  status %d%s"
              % (cfg, M, E, ms, M * S * D / (ms * 1e-3), chk[2], same), flush=True)
    del out
    torch.cuda.empty_cache()
