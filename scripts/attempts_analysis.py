"""Analysis tool: where the step attempts of a day go (host build of the quad program, scripts/steps_attempts.cpp)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
CAP = 8
w = bench.build_workload("2004", M)
f = np.ascontiguousarray(w["forcing"]); mp = np.ascontiguousarray(w["member"]); sc = np.ascontiguousarray(w["sc"][0])
D = f.shape[0]
lib = C.CDLL(os.path.join(ROOT, "build", "libsteps_attempts.so"))
rec_t = np.dtype([("t", "f4"), ("h", "f4"), ("en2", "f4"), ("k", "i2"), ("acc", "i1"), ("pad", "i1")])
rec = np.zeros((M, D, CAP), dtype=rec_t); nrec = np.zeros((M, D), dtype=np.int32)
lib.steps_attempts(C.c_int(M), C.c_int(D), C.c_int(CAP), f.ctypes.data_as(C.c_void_p), mp.ctypes.data_as(C.c_void_p),
                   sc.ctypes.data_as(C.c_void_p), C.c_double(1e-7), C.c_double(1e-10), rec.ctypes.data_as(C.c_void_p),
                   nrec.ctypes.data_as(C.c_void_p))
print("attempts/day mean %.2f" % nrec.mean())
P = f[:, 0]
wet = P > 0
for k in range(CAP):
    r = rec[:, 1:, k]
    have = r["k"] == k + 1
    acc = r["acc"][have]
    en = r["en2"][have]
    # ideal factor for this attempt: 0.9 * en2^-0.1
    fac = 0.9 * np.maximum(en, 1e-30) ** -0.1
    print("attempt %d of the day: present %.3f, rejected %.3f, median ideal factor %.2f, p10 %.2f p90 %.2f, median h %.4f"
          % (k + 1, have.mean(), 1 - acc.mean(), np.median(fac), np.percentile(fac, 10), np.percentile(fac, 90), np.median(r["h"][have])))
# rejected attempts by position
tot_rej = 0
first = rec[:, 1:, 0]
print("first attempt: rejected on wet days %.3f, on dry days %.3f" % (1 - first["acc"][:, wet[1:]].mean(), 1 - first["acc"][:, ~wet[1:]].mean()))
fac1 = 0.9 * np.maximum(first["en2"], 1e-30) ** -0.1
print("first-attempt ideal factor: wet median %.2f, dry median %.2f" % (np.median(fac1[:, wet[1:]]), np.median(fac1[:, ~wet[1:]])))
print("attempts per day: wet %.2f dry %.2f ; wet days %d of %d" % (nrec[:, wet].mean(), nrec[:, ~wet].mean(), wet.sum(), D))
np.savez_compressed(os.path.join(ROOT, "build", "attempts.npz"), rec=rec, nrec=nrec, P=P)
