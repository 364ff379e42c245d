"""Where do the ~0.45 ms per step go that a multi-rank bench step takes longer than the single-GPU one?
Run under torchrun; every rank times (a) the calibration launch alone, (b) launch + all-gather, for its own shard of the
M_total-member ensemble and for the SAME 10^4 members on every rank (no sample variation between ranks)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
from simplyp_b200 import ensemble as ens, model as spm, packing as pk
from simplyp_b200.engine import Engine
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = Engine(local)
M = 10000
res = {}
for mode in ("own_shard", "same_members"):
    w = bench.build_workload("2004", M * world if mode == "own_shard" else M)
    lo, hi = ens.shard_bounds(M * world, world, rank) if mode == "own_shard" else (0, M)
    opt = spm.make_options(w["p_SU"], w["p"], w["dyn"], w["topo"])
    d = [eng.to_device(x) for x in (w["forcing"], w["member"][lo:hi], w["sc"][lo:hi], w["obs_m"], w["desc"])]
    gb = ens.GatherBuffers(M * world, (2, pk.NSTAT), eng.device)
    diag = torch.zeros((M, 1, pk.NDIAG), dtype=torch.int64, device=eng.device)
    po, pid = w["topo"].parent_offsets, w["topo"].parent_ids
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)
    for gather in (False, True):
        for _ in range(3):
            eng.calibrate(d[0], d[1], d[2], po, pid, d[3], d[4], opt, stats=gb.local, diag=diag)
            if gather: gb.gather()
        torch.cuda.synchronize(); dist.barrier()
        tot = 0.0
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.calibrate(d[0], d[1], d[2], po, pid, d[3], d[4], opt, stats=gb.local, diag=diag)
            if gather: gb.gather()
            e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        res["%s/%s" % (mode, "launch+gather" if gather else "launch")] = tot / 10
    res[mode + "/max_member_attempts"] = int(diag[:, 0, 0].max().item())
allres = [None] * world
dist.all_gather_object(allres, res)
if rank == 0:
    for k in res:
        print(k, ["%.3f" % r[k] if isinstance(r[k], float) else r[k] for r in allres])
dist.destroy_process_group()
