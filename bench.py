#!/usr/bin/env python
"""Benchmark of the SimplyP daily mass-balance integration path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--members M_PER_GPU] [--period 2004|full]

Workload (BASELINE.json configs[1]): the Tarland set-up (1 sub-catchment, shipped 2004 period, both
dynamic options on), a Latin-hypercube ensemble of 10^4 parameter sets PER GPU, fused goodness-of-fit
statistics (NSE, log-NSE, Gaussian log-likelihood, r2, bias, nRMSD) against observed Q and TDP.
One "step" = one pass of the whole ensemble over the whole period.  Metric: member-sub-catchment-days
per second, whole job.  Members are sharded over ranks with no data-path collective; the only
collective is the all-gather of the per-member statistics (inside the timed step).

Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for how each field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "member-sub-catchment-days/sec"
UNIT = "member-SC-days/s"
F_RHS, F_STEP_EXTRA, F_DAY = 140.0, 700.0, 120.0      # SURVEY.md §8(d) algorithmic flop counts


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--members", type=int, default=10000, help="ensemble members per GPU")
    ap.add_argument("--period", default="2004", choices=["2004", "full"])
    ap.add_argument("--rtol", type=float, default=None)
    ap.add_argument("--atol", type=float, default=None)
    ap.add_argument("--pilot-days", type=int, default=0, help="cost pilot: 0 = library default, -1 = off")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def period_dates(period):
    return ("2004-01-01", "2004-12-31") if period == "2004" else ("1981-01-01", "2010-12-31")


# ------------------------------------------------------------------------------------------ workload
def build_workload(period, n_members, seed=20260101, member_offset=0):
    """Arrays of the calibration call for `n_members` members (deterministic in the global member index)."""
    from simplyp_b200 import ensemble as ens, model as spm, packing as pk, tarland
    st, end = period_dates(period)
    p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(st, end, dynamic="y")
    topo = pk.build_topology(p_struc, p["SC_list"])
    samples = ens.latin_hypercube(n_members, seed=seed + member_offset)
    member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo.sc_ids), samples)
    obs_m, desc, labels = pk.obs_arrays(obs, topo, met.index, ("Q", "TDP"))
    return dict(p_SU=p_SU, dyn=dyn, p=p, p_LU=p_LU, p_SC=p_SC, p_struc=p_struc, met=met, obs=obs, topo=topo,
                samples=samples, member=member, sc=sc, forcing=pk.forcing_matrix(met), obs_m=obs_m, desc=desc,
                labels=labels)


# ------------------------------------------------------------------------------------------ CPU legs (oracle port)
def _cpu_worker(args):
    """Integrate `count` members with the oracle port; returns (member_sc_days, seconds)."""
    period, first, count, rtol, atol, seed, n_total = args
    sys.path.insert(0, ROOT)
    from oracle import simplyp_oracle as orc
    from simplyp_b200 import ensemble as ens
    w = build_workload(period, n_total, seed=seed)
    t0 = time.perf_counter()
    days = 0
    for i in range(first, first + count):
        pi, pLUi, pSCi = ens.apply_member_to_pandas(w["samples"], i, w["p"], w["p_LU"], w["p_SC"])
        orc.run_simply_p(w["met"], w["p_struc"], w["p_SU"], pLUi, pSCi, pi, w["dyn"], rtol=rtol, atol=atol,
                         mxstep=50000)
        days += len(w["met"])
    return days, time.perf_counter() - t0


def cpu_port_throughput(period, members_per_core, cores, rtol, atol):
    """Oracle port (scipy LSODA, one member per task) on `cores` processes; member-SC-days/s aggregate."""
    import multiprocessing as mp
    tasks = [(period, c * members_per_core, members_per_core, rtol, atol, 20260101, members_per_core * cores)
             for c in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_worker(tasks[0])]
    else:
        with mp.get_context("spawn").Pool(cores) as pool:
            res = pool.map(_cpu_worker, tasks)
    wall = time.perf_counter() - t0
    days = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return days / busy, days, busy, wall


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU algorithm (oracle port: per-day scipy odeint/LSODA at the
    reference's rtol=0.01 with a Python RHS callback) on all host cores.  The reference is pure Python and
    cannot travel to the GPU box, so the port stands in for it (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    days_per_member = 366 if args.period == "2004" else 10957
    per_core = 2 if args.period == "2004" else 1
    n_days_eff = None
    times = []
    days = 0
    for it in range(args.warmup + args.steps):
        v, d, busy, wall = cpu_port_throughput(args.period, per_core, cores, 0.01, None)
        if it >= args.warmup:
            times.append(busy)
            days = d
    ms = 1e3 * float(np.mean(times))
    value = days / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "Tarland forcing/obs fixtures; "
        "Latin-hypercube parameter sets (seed 20260101)",
        "config": {"workload": "Tarland %s (D=%d days, S=1), %d-member LHS ensemble per GPU, fused NSE/log-NSE/"
                               "log-likelihood vs observed Q and TDP, Dynamic_EPC0/erodibility on"
                               % (args.period, days_per_member, args.members),
                   "sample": "each step integrates a %d-member Latin-hypercube sample drawn by the same generator and seed "
                             "(bounded sample; member-SC-days/s does not depend on the ensemble size on the CPU)"
                             % (per_core * cores),
                   "members_per_step": per_core * cores, "days": days_per_member, "sub_catchments": 1,
                   "rtol": 0.01, "atol": "odeint default (the reference's own solver settings, model.py:640)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d members x %d days per step, one member per task on %d processes, "
                                   "scipy odeint (LSODA) rtol=0.01 as the reference calls it" % (per_core * cores, days_per_member, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.gpu_index = gpu_index
        self.path = "/tmp/simplyp_clocks_%d_%d.csv" % (os.getpid(), gpu_index)

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, smax, reasons = [], [], set()
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons), samples=len(sm),
                       window="warm-up + timed steps")
        return out


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from simplyp_b200 import _cabi, ensemble as ens, model as spm, packing as pk
    from simplyp_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    M_local = args.members
    M_total = M_local * world
    w = build_workload(args.period, M_total)
    lo, hi = ens.shard_bounds(M_total, world, rank)
    opt = spm.make_options(w["p_SU"], w["p"], w["dyn"], w["topo"], 1.0, args.rtol, args.atol)
    opt.pilot_days = args.pilot_days
    kernel_name = "simplyp_quad_kernel<cal>"
    eng = Engine(local_rank)
    S, D, V = w["topo"].n_sc, w["forcing"].shape[0], w["obs_m"].shape[0]

    # ---- device-resident leg ("value"): inputs already in HBM
    d_forc = eng.to_device(w["forcing"])
    d_mem = eng.to_device(w["member"][lo:hi])
    d_sc = eng.to_device(w["sc"][lo:hi] if w["sc"].shape[0] > 1 else w["sc"])
    d_obs = eng.to_device(w["obs_m"])
    d_desc = eng.to_device(w["desc"])
    stats = torch.empty((hi - lo, V, pk.NSTAT), dtype=torch.float64, device=eng.device)
    diag = torch.zeros((hi - lo, S, pk.NDIAG), dtype=torch.int64, device=eng.device)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=eng.device)   # > 126 MB L2
    po, pid = w["topo"].parent_offsets, w["topo"].parent_ids

    def step():
        st, _ = eng.calibrate(d_forc, d_mem, d_sc, po, pid, d_obs, d_desc, opt, stats=stats, diag=diag)
        return ens.all_gather_stats(st, M_total) if world > 1 else st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks are sampled from before the warm-up steps (nvidia-smi needs ~0.1 s to start and a timed region of 10
    # steps lasts 0.15 s) to the end of the timed steps: every sample is taken under the bench load
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        step()
    barrier()

    launches0 = _cabi.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    gathered = None
    for k in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations (outside the events)
        ev[k][0].record()
        gathered = step()
        ev[k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = _cabi.launch_count() - launches0
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([float(sum(step_ms))], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    value = M_total * S * D / (ms_per_step * 1e-3)

    # ---- integrator work counters for the roofline (device-counted)
    dg = diag.sum(dim=(0, 1)).cpu().numpy().astype(float)
    n_steps, n_rej, n_rhs = dg[0], dg[1], dg[2]
    status_any = int(diag[..., 3].max().item())
    flops_local = n_rhs * F_RHS + n_steps * F_STEP_EXTRA + (hi - lo) * S * D * F_DAY
    kernel_ms = float(np.mean(step_ms))
    achieved_tf = flops_local / (kernel_ms * 1e-3) / 1e12
    fp64_peak = _cabi.measure_fp64_peak(local_rank, 3) if rank == 0 else None

    # ---- end-to-end leg: host buffers through the C-ABI, H2D + D2H inside the timed region
    h = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() for k, v in
         (("forcing", w["forcing"]), ("member", w["member"][lo:hi]),
          ("sc", w["sc"][lo:hi] if w["sc"].shape[0] > 1 else w["sc"]), ("obs", w["obs_m"]))}
    h2d = sum(v.nbytes for v in h.values()) + w["desc"].nbytes
    d2h = (hi - lo) * V * pk.NSTAT * 8 + (hi - lo) * S * pk.NDIAG * 8
    for _ in range(2):
        _cabi.calibrate_host(h["forcing"], h["member"], h["sc"], po, pid, h["obs"], w["desc"], opt, device=local_rank)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st_host, _dg = _cabi.calibrate_host(h["forcing"], h["member"], h["sc"], po, pid, h["obs"], w["desc"], opt,
                                            device=local_rank)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = M_total * S * D * args.steps / float(e2e_s.item())
    same = bool(np.array_equal(st_host, stats.cpu().numpy(), equal_nan=True))

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic, traffic_src = None, None
        try:     # measured DRAM traffic of this kernel at this size, from the committed ncu capture (null if none)
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ent = tj.get("%s|%d|%d" % (kernel_name, hi - lo, D))
            if ent:
                traffic, traffic_src = ent["bytes"], ent["source"]
        except Exception:
            pass
        alg_bytes = (d_forc.numel() + d_mem.numel() + d_sc.numel() + d_obs.numel()) * 8 + stats.numel() * 8 * 2
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "Tarland forcing/obs fixtures (tests/golden); "
            "Latin-hypercube parameter sets, seed 20260101",
            "config": {"workload": "Tarland %s (D=%d days, S=1), %d-member LHS ensemble per GPU, fused NSE/log-NSE/"
                                   "log-likelihood vs observed Q and TDP, Dynamic_EPC0/erodibility on" % (args.period, D, M_local),
                       "members_total": M_total, "days": D, "sub_catchments": S, "obs_series": [list(l) for l in w["labels"]],
                       "rtol": opt.rtol, "atol": opt.atol, "parallelism": "ensemble members sharded over %d GPU(s)" % world,
                       "l2": "256 MB buffer written between timed steps (outside the per-step CUDA events)",
                       "launches_per_step": "pilot pass (days 0-7, member order) + counting sort (2) + observation "
                                            "constants + main pass (days 8-end, cost order, blocks placed by SM)"},
            "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "simplyp_calibrate_host (C-ABI, pinned host buffers)", "matches_device_leg": same},
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": (achieved_tf / fp64_peak) if fp64_peak else None, "traffic": traffic,
                         "traffic_source": traffic_src,
                         "peak_source": "simplyp_measure_fp64_peak (DFMA probe, this run); MEASURED_PEAKS.json has no FP64 figure",
                         "kernel": kernel_name, "kernel_ms": kernel_ms,
                         "algorithmic_flops_per_launch": flops_local,
                         "steps_per_member_day": n_steps / ((hi - lo) * S * D), "rejected_frac": n_rej / max(n_steps, 1.0),
                         "rhs_per_member_day": n_rhs / ((hi - lo) * S * D),
                         "hbm": {"achieved_gbs": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                 "frac": alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                                 "algorithmic_bytes_per_launch": int(alg_bytes),
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
            "clocks": clocks, "integrator_status_bits": status_any, "wall_s_timed_region": t_wall,
        }
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            per_core = 1 if args.period == "full" else max(1, int(round(args.cpu_seconds / 1.2)))
            try:
                v, days, busy, wall = cpu_port_throughput(args.period, per_core, cores, opt.rtol, opt.atol)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                        "sample": "%d members x %d days (same LHS members), oracle port = per-day scipy "
                                                  "odeint/LSODA with a Python RHS at the GPU run's rtol/atol, one member "
                                                  "per task on %d processes (%.1f s)" % (per_core * cores, D, cores, busy)}
            except Exception as e:  # keep the bench line even if the CPU leg cannot run
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": cores, "kind": "port",
                                        "sample": "failed: %r" % (e,)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
