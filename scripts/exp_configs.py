"""BASELINE configs 3 and 5 at full size on one GPU: timing (CUDA events) and oracle-free properties.

    python scripts/exp_configs.py [3] [5] [--members M]

config 3: 256-sub-catchment branching network, 30-year daily forcing, both dynamic options on
config 5: 4096 sub-catchments x 3 land-use classes, 50-year daily run, full daily output kept in HBM
Output stays on the device (config 5 writes 15 GB per member).
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from simplyp_b200 import inputs as spi, model as spm, packing as pk, synthetic, tarland
from simplyp_b200.engine import Engine


def build(cfg, members=1):
    return synthetic.scale_config(cfg, members)


def main():
    cfgs = [int(a) for a in sys.argv[1:] if a in ("3", "5")] or [3, 5]
    members = int(sys.argv[sys.argv.index("--members") + 1]) if "--members" in sys.argv else 1
    eng = Engine(0)
    for cfg in cfgs:
        w = build(cfg, members)
        topo, opt = w["topo"], w["opt"]
        S, D, M = topo.n_sc, w["forcing"].shape[0], members
        d_f, d_m, d_s = eng.to_device(w["forcing"]), eng.to_device(w["member"]), eng.to_device(w["sc"])
        out = torch.empty((M, S, D, pk.NOUT), dtype=torch.float64, device=eng.device)
        diag = torch.zeros((M, S, pk.NDIAG), dtype=torch.int64, device=eng.device)
        po, pid = topo.parent_offsets, topo.parent_ids
        eng.run(d_f, d_m, d_s, po, pid, opt, out=out, diag=diag)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 2
        e0.record()
        for _ in range(reps):
            eng.run(d_f, d_m, d_s, po, pid, opt, out=out, diag=diag)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        dg = diag.sum(dim=(0, 1)).cpu().numpy()
        units = M * S * D
        from simplyp_b200 import _cabi
        nl = int(_cabi.topology_levels(po, pid)[0])
        print(json.dumps({"config": cfg, "members": M, "sub_catchments": S, "days": D, "levels": nl, "ms": ms,
                          "member_sc_days_per_s": units / (ms * 1e-3), "output_GB": out.numel() * 8 / 1e9,
                          "write_GBps": out.numel() * 8 / 1e9 / (ms * 1e-3),
                          "steps_per_item_day": float(dg[0]) / units, "status_bits": int(diag[..., 3].max().item()),
                          "finite": bool(torch.isfinite(out).all().item())}), flush=True)
        spd = (diag[0, :, 0].cpu().numpy() / D)
        lv = _cabi.topology_levels(po, pid)[1]
        top = np.argsort(-spd)[:5]
        print("  steps/day per reach: median %.1f, p90 %.1f, max %.1f; heaviest reaches (index, level, steps/day, rejected frac): %s"
              % (np.median(spd), np.percentile(spd, 90), spd.max(),
                 [(int(i), int(lv[i]), round(float(spd[i]), 1), round(float(diag[0, i, 1].item() / max(diag[0, i, 0].item(), 1)), 2)) for i in top]))
        qr = out[0, :, :, 5].mean(dim=1).cpu().numpy()
        print("  mean daily flow Qr (mm/d over own area): median %.2f max %.1f" % (np.median(qr), qr.max()))
        del out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
