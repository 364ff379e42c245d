"""Multi-GPU invariance check: run under torchrun with N ranks.
Every rank integrates its contiguous shard of a Latin-hypercube ensemble and the statistics are all-gathered (NCCL);
rank 0 then integrates the WHOLE ensemble alone and asserts that the gathered statistics are bitwise identical.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        scripts/check_sharding.py [members_total]
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
from simplyp_b200 import ensemble as ens, model as spm, packing as pk
from simplyp_b200.engine import Engine

M = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = bench.build_workload("2004", M)
opt = spm.make_options(w["p_SU"], w["p"], w["dyn"], w["topo"])
eng = Engine(local)
lo, hi = ens.shard_bounds(M, world, rank)
po, pid = w["topo"].parent_offsets, w["topo"].parent_ids
d = {k: eng.to_device(w[k]) for k in ("forcing", "obs_m", "desc")}


def run(a, b):
    st, dg = eng.calibrate(d["forcing"], eng.to_device(w["member"][a:b]), eng.to_device(w["sc"][a:b]), po, pid,
                           d["obs_m"], d["desc"], opt)
    return st


for _ in range(2):
    g = ens.all_gather_stats(run(lo, hi), M)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
gathered = ens.all_gather_stats(run(lo, hi), M)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=eng.device, dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ok = True
if rank == 0:
    whole = run(0, M)
    torch.cuda.synchronize()
    ok = bool(torch.equal(torch.nan_to_num(gathered, nan=-7.0), torch.nan_to_num(whole, nan=-7.0)))
    print(json.dumps({"ranks": world, "members_total": M, "days": 366, "ms_sharded_incl_all_gather": float(ms.item()),
                      "member_sc_days_per_s": M * 366 / (float(ms.item()) * 1e-3),
                      "gathered_equals_single_gpu_bitwise": ok}), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
