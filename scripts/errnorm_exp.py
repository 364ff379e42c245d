"""Analysis tool: attempts per day and output accuracy of the quad program (host build) under error-norm weights."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
w = bench.build_workload("2004", 10000)
sel = np.arange(0, 10000, 10000 // M)[:M]
f = np.ascontiguousarray(w["forcing"]); mp = np.ascontiguousarray(w["member"][sel]); sc = np.ascontiguousarray(w["sc"][0])
D = f.shape[0]
lib = C.CDLL(os.path.join(ROOT, "build", "libsteps_series.so"))
def run(rtol, atol, wts):
    steps = np.zeros(M, dtype=np.int64); rej = np.zeros(M, dtype=np.int64); out = np.zeros((M, D, 12))
    wv = np.array(wts, dtype=np.float64)
    lib.steps_series(C.c_int(M), C.c_int(D), f.ctypes.data_as(C.c_void_p), mp.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p),
                     C.c_double(rtol), C.c_double(atol), wv.ctypes.data_as(C.c_void_p), steps.ctypes.data_as(C.c_void_p),
                     rej.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return steps, rej, out
def series(out):
    q = out[..., 0]
    return np.stack([q, out[..., 1] / q, out[..., 2] / q, out[..., 3] / q], -1)      # Q and the three concentrations
truth_file = os.path.join(ROOT, "build", "errnorm_truth_%d.npy" % M)
if os.path.exists(truth_file):
    truth = np.load(truth_file)
else:
    _, _, truth = run(1e-11, 1e-14, [1, 1, 1, 1, 1000]); np.save(truth_file, truth)
ts = series(truth)
def report(label, rtol, atol, wts):
    steps, rej, out = run(rtol, atol, wts)
    s = series(out)
    rel = np.abs(s - ts) / np.abs(ts)
    per_member = rel.reshape(M, -1).max(1)
    st = np.abs(out[..., 4:] - truth[..., 4:]) / np.maximum(np.abs(truth[..., 4:]), 1e-300)
    print("%-44s attempts/day %.2f (max member %.1f) rej %.1f%%  flows+conc: worst %.2e median-member %.2e p95 %.2e | states worst %.1e"
          % (label, steps.mean() / D, steps.max() / D, 100.0 * rej.sum() / steps.sum(), per_member.max(), np.median(per_member),
             np.percentile(per_member, 95), st.max()), flush=True)
report("current (1,1,1,1,1000) rtol 1e-7", 1e-7, 1e-10, [1, 1, 1, 1, 1000])
for spec in sys.argv[2:]:
    parts = spec.split(",")
    rtol = float(parts[0]); wts = [float(x) for x in parts[1:]]
    report("rtol %g wB %g wACC %g wU %g wVG %g wSOIL %g" % tuple([rtol] + wts), rtol, rtol * 1e-3, wts)
