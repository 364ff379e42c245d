"""Analysis tool: lock-step cost of the bench ensemble from per-member per-day attempt counts of the quad program on the
host (scripts/steps_quad_day.cpp compiled with the given flags): members ordered by the cost of the first PILOT days,
eight to a warp.   python scripts/lockstep_model.py 10000 "" "-DSP_TOL_GMAX=1.0"
"""
import ctypes as C, hashlib, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
M = int(sys.argv[1])
w = bench.build_workload("2004", M)
f = np.ascontiguousarray(w["forcing"]); mp = np.ascontiguousarray(w["member"]); sc = np.ascontiguousarray(w["sc"][0])
D = f.shape[0]
for flags in sys.argv[2:] or [""]:
    tag = hashlib.md5(flags.encode()).hexdigest()[:10]
    so = os.path.join(ROOT, "build", "libsteps_quad_day_%s.so" % tag)
    subprocess.check_call(["g++", "-O2", "-fopenmp", "-std=c++17", "-shared", "-fPIC"] + flags.split() + ["-o", so, os.path.join(ROOT, "scripts", "steps_quad_day.cpp")])
    lib = C.CDLL(so)
    steps = np.zeros((M, D), dtype=np.uint16)
    lib.steps_quad_day(C.c_int(M), C.c_int(D), f.ctypes.data_as(C.c_void_p), mp.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p),
                       C.c_double(1e-7), C.c_double(1e-10), steps.ctypes.data_as(C.c_void_p))
    s = steps.astype(np.int64)
    tot = s.sum(1)
    print("[%s] attempts/member-day %.2f, heaviest member %.1f/day (%d attempts)" % (flags, tot.mean() / D, tot.max() / D, tot.max()))
    for pilot in (8, 16, D):
        order = np.argsort(-s[:, :pilot].sum(1), kind="stable")
        n = (M // 8) * 8
        g = s[order[:n]].reshape(-1, 8, D)
        lock = g.max(1).sum(1)
        print("   pilot %3d days: lane efficiency %.3f, warp attempts total %.3e, heaviest warp %d, mean warp %.0f, corr(pilot cost, total) %.3f"
              % (pilot, s[:n].sum() / (8.0 * lock.sum()), lock.sum(), lock.max(), lock.mean(),
                 np.corrcoef(s[:, :pilot].sum(1), tot)[0, 1]))
