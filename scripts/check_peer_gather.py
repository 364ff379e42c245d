"""Fused all-gather check (run under torchrun with N ranks of one box): the calibration kernel that stores every member's
statistics straight into all ranks' buffers over peer memory (ensemble.PeerGather, simplyp_calibrate_gather_device) against
the NCCL all-gather of the same statistics — bitwise equal on every rank, ragged shards, several passes in a row (the two
buffer sets alternate) — and the time per pass of both.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
        scripts/check_peer_gather.py [members_total] [passes]
"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
from simplyp_b200 import ensemble as ens, model as spm, packing as pk
from simplyp_b200.engine import Engine

M = int(sys.argv[1]) if len(sys.argv) > 1 else 20003
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = bench.build_workload("2004", M)
opt = spm.make_options(w["p_SU"], w["p"], w["dyn"], w["topo"])
eng = Engine(local)
lo, hi = ens.shard_bounds(M, world, rank)
po, pid = w["topo"].parent_offsets, w["topo"].parent_ids
d = {k: eng.to_device(w[k]) for k in ("forcing", "obs_m", "desc")}
d_mem, d_sc = eng.to_device(w["member"][lo:hi]), eng.to_device(w["sc"][lo:hi])
V = w["obs_m"].shape[0]
gb = ens.GatherBuffers(M, (V, pk.NSTAT), eng.device)
pg = ens.PeerGather(M, (V, pk.NSTAT), eng.device)
diag = torch.zeros((hi - lo, 1, pk.NDIAG), dtype=torch.int64, device=eng.device)


def nccl_pass():
    eng.calibrate(d["forcing"], d_mem, d_sc, po, pid, d["obs_m"], d["desc"], opt, stats=gb.local, diag=diag)
    return gb.gather()


def peer_pass():
    g, _ = eng.calibrate(d["forcing"], d_mem, d_sc, po, pid, d["obs_m"], d["desc"], opt, diag=diag, peer_gather=pg)
    return g


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(passes):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / passes], device=eng.device, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


ref = nccl_pass().clone()
ok = True
for k in range(5):                      # both buffer sets, several steps
    got = peer_pass()
    torch.cuda.synchronize()
    ok = ok and bool(torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(ref, nan=-7.0)))
    dist.barrier()
status = int(diag[..., 3].max().item())
ms_nccl = timed(nccl_pass)
ms_peer = timed(peer_pass)
ms_nccl2 = timed(nccl_pass)
ms_peer2 = timed(peer_pass)
flag = torch.tensor([1 if ok else 0], device=eng.device)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"ranks": world, "members_total": M, "passes": passes, "peer_gather_equals_nccl_bitwise_on_all_ranks": bool(flag.item()),
                      "status_bits": status, "ms_per_pass_nccl": [ms_nccl, ms_nccl2], "ms_per_pass_peer": [ms_peer, ms_peer2]}), flush=True)
dist.barrier()
pg.close()
dist.destroy_process_group()
sys.exit(0 if bool(flag.item()) else 1)
