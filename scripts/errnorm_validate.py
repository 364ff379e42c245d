"""Analysis tool: the parity checks of the CPU tier + N members of the bench ensemble against the oracle (LSODA at
rtol=1e-10), run on a host build of the quad program compiled with the given error-norm weights.

  python scripts/errnorm_validate.py N "-DSP_W_B=0.1 -DSP_W_ACC=0.1" ["-D..." ...]
"""
import ctypes as C
import hashlib
import multiprocessing as mp
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from simplyp_b200 import model as spm, packing as pk  # noqa: E402
from tests import hostemu, parity  # noqa: E402
from tests.util import max_rel  # noqa: E402
from tests.test_gpu_parity import _oracle_member  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
COLS = ["Q_cumecs", "SS_mgl", "TDP_mgl", "PP_mgl", "Qr", "Msus_kg/day", "TDP_kg/day", "PP_kg/day"]


def runner_for(flags):
    tag = hashlib.md5(flags.encode()).hexdigest()[:10]
    lib = os.path.join(ROOT, "build", "libhostemu_%s.so" % tag)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas"] + flags.split() +
                          ["-o", lib, hostemu.SRC])
    hostemu._lib = C.CDLL(lib)
    return hostemu.run_quad


def truth(N, M=10000):
    f = os.path.join(ROOT, "build", "bench_truth_%d.npz" % N)
    picked = list(range(0, M, M // N))[:N]
    if os.path.exists(f):
        z = np.load(f)
        return picked, z["want"]
    with mp.get_context("spawn").Pool(os.cpu_count()) as pool:
        res = pool.map(_oracle_member, [(i, M, 366) for i in picked], chunksize=1)
    want = np.stack([r[1] for r in res])
    np.savez_compressed(f, want=want)
    return picked, want


def main():
    N = int(sys.argv[1])
    picked, want = truth(N)
    w = bench.build_workload("2004", 10000)
    opt = spm.make_options(w["p_SU"], w["p"], w["dyn"], w["topo"])
    for flags in sys.argv[2:] or [""]:
        run = runner_for(flags)
        line = []
        for name, fn in (("tarland n", lambda: parity.check_tarland(run, GOLDEN, "n")),
                         ("tarland y", lambda: parity.check_tarland(run, GOLDEN, "y")),
                         ("network5", lambda: parity.check_network(run, GOLDEN)),
                         ("lhs12", lambda: parity.check_ensemble_series(run, GOLDEN)),
                         ("stiff 0.05", lambda: parity.check_stiff_chain(run, 0.05)),
                         ("stiff 5", lambda: parity.check_stiff_chain(run, 5.0))):
            try:
                r = fn()
                line.append("%s ok%s" % (name, (" %.2e" % r) if isinstance(r, float) else ""))
            except AssertionError as e:
                line.append("%s FAIL %s" % (name, str(e)[:80]))
        out, dg = run(w["forcing"], w["member"][picked], w["sc"][picked], w["topo"].parent_offsets, w["topo"].parent_ids, opt)
        worst = []
        for j, i in enumerate(picked):
            A = float(w["sc"][i, 0, pk.SC_INDEX["A_catch"]])
            _tc, r = spm.raw_to_frames(out[j, 0], w["met"].index, A, w["p"]["Msoil_m2"], w["p"]["f_TDP"], "None", None)
            got = r[COLS].to_numpy(float)
            worst.append([max_rel(got[:, k], want[j][:, k]) for k in range(len(COLS))])
        worst = np.array(worst)
        per = worst.max(1)
        print("[%s]\n  %s\n  bench %d members: attempts/day %.2f (max member %.1f, rejected %.1f%%) worst %.2e p99 %.2e p95 %.2e median %.2e; "
              "worst by column %s" % (flags, "; ".join(line), N, dg[:, 0, 0].mean() / 366.0, dg[:, 0, 0].max() / 366.0,
                                      100.0 * dg[:, 0, 1].sum() / dg[:, 0, 0].sum(), per.max(), np.percentile(per, 99),
                                      np.percentile(per, 95), np.median(per),
                                      " ".join("%.1e" % x for x in worst.max(0))), flush=True)


if __name__ == "__main__":
    main()
