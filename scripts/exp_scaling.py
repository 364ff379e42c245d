"""Throughput of the calibration kernel vs ensemble size / block size (device-resident, CUDA events)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from simplyp_b200 import _cabi, model as spm, packing as pk
from simplyp_b200.engine import Engine

def main():
    sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1024,10000,40000,160000").split(",")]
    blocks = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0").split(",")]
    eng = Engine(0)
    for M in sizes:
        w = bench.build_workload("2004", M)
        d = [eng.to_device(w[k]) for k in ("forcing", "member", "sc", "obs_m", "desc")]
        for tb in blocks:
            opt = spm.make_options(w["p_SU"], w["p"], w["dyn"], w["topo"])
            for _ in range(2):
                st, dg = eng.calibrate(d[0], d[1], d[2], w["topo"].parent_offsets, w["topo"].parent_ids, d[3], d[4], opt)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 3
            for _ in range(n):
                st, dg = eng.calibrate(d[0], d[1], d[2], w["topo"].parent_offsets, w["topo"].parent_ids, d[3], d[4], opt)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            steps = dg[..., 0].double()
            print(json.dumps({"M": M, "block": tb, "ms": ms, "msd_per_s": M * 366 / ms * 1e3,
                              "steps_per_day_mean": float(steps.mean() / 366), "steps_per_day_max": float(steps.max() / 366),
                              "steps_per_day_min": float(steps.min() / 366)}), flush=True)

if __name__ == "__main__":
    main()
