"""Analysis tool: warp scheduling policies from per-member per-day step counts (scripts/steps_per_day.cpp)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
M = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
w = bench.build_workload("2004", M)
lib = C.CDLL(os.path.join(ROOT, "build", "libsteps.so"))
D = w["forcing"].shape[0]
steps = np.zeros((M, D), dtype=np.uint16)
f = np.ascontiguousarray(w["forcing"]); mp = np.ascontiguousarray(w["member"]); sc = np.ascontiguousarray(w["sc"][0] if w["sc"].ndim == 3 else w["sc"])
t0 = time.time()
lib.steps_per_day(C.c_int(M), C.c_int(D), f.ctypes.data_as(C.c_void_p), mp.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p),
                  C.c_double(1e-7), C.c_double(1e-10), steps.ctypes.data_as(C.c_void_p))
print("emulated %d members in %.1f s" % (M, time.time() - t0))
np.save(os.path.join(ROOT, "build", "steps_%d.npy" % M), steps)
tot = steps.sum(1).astype(float)
print("steps/day: mean %.2f min %.2f max %.2f" % (tot.mean() / D, tot.min() / D, tot.max() / D))
def policies(order, G, label):
    s = steps[order].astype(float)
    n = (M // G) * G
    s = s[:n].reshape(-1, G, D)
    flat = s.sum(2).max(1)            # flattened loop: warp time = slowest member's total
    lock = s.max(1).sum(1)            # day lock-step: sum over days of the slowest member that day
    ideal = s.sum()                   # total member-steps
    print("%-28s G=%2d  flattened eff %.3f  lockstep eff %.3f   (warp-steps flat %.0f lock %.0f)" %
          (label, G, ideal / (flat.sum() * G), ideal / (lock.sum() * G), flat.sum(), lock.sum()))
for G in (8, 32):
    policies(np.arange(M), G, "unsorted")
    policies(np.argsort(tot), G, "sorted by total steps")
    pilot = steps[:, :20].sum(1)
    policies(np.argsort(pilot, kind="stable"), G, "sorted by 20-day pilot")
    pilot = steps[:, :40].sum(1)
    policies(np.argsort(pilot, kind="stable"), G, "sorted by 40-day pilot")
