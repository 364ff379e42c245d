"""Analysis tool: step attempts of the quad program under controller variants (compile-time macros)."""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
M = 512
w = bench.build_workload("2004", M)
f = np.ascontiguousarray(w["forcing"]); mp = np.ascontiguousarray(w["member"]); sc = np.ascontiguousarray(w["sc"][0])
base = None
for label, flags in [("current (0.9, 5)", []), ("safety 0.95", ["-DSP_CTRL_SAFETY=0.95", "-DSP_CTRL_MAXGROW=5.0"]),
                     ("safety 0.85", ["-DSP_CTRL_SAFETY=0.85", "-DSP_CTRL_MAXGROW=5.0"]),
                     ("max growth 10", ["-DSP_CTRL_SAFETY=0.9", "-DSP_CTRL_MAXGROW=10.0"]),
                     ("max growth 3", ["-DSP_CTRL_SAFETY=0.9", "-DSP_CTRL_MAXGROW=3.0"]),
                     ("day start 0.1", ["-DSP_DAYSTART_FAC=0.1"]), ("day start 0.3", ["-DSP_DAYSTART_FAC=0.3"])] + [(" ".join(sys.argv[1:]), sys.argv[1:])] * (len(sys.argv) > 1):
    so = os.path.join(ROOT, "build", "libsteps_quad_%d.so" % abs(hash(label)))
    subprocess.check_call(["g++", "-O2", "-fopenmp", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas"] + flags + ["-o", so, os.path.join(ROOT, "scripts", "steps_quad.cpp")])
    lib = C.CDLL(so)
    steps = np.zeros(M, dtype=np.int64); rej = np.zeros(M, dtype=np.int64); chk = np.zeros(M)
    lib.steps_quad(C.c_int(M), C.c_int(f.shape[0]), f.ctypes.data_as(C.c_void_p), mp.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p),
                   C.c_double(1e-7), C.c_double(1e-10), steps.ctypes.data_as(C.c_void_p), rej.ctypes.data_as(C.c_void_p), chk.ctypes.data_as(C.c_void_p))
    if base is None: base = chk.copy()
    print("%-28s attempts/day mean %.2f max %.2f  rejected %.1f %%  max rel change of sum(Qr) vs current %.1e" %
          (label, steps.mean() / f.shape[0], steps.max() / f.shape[0], 100.0 * rej.sum() / steps.sum(), np.max(np.abs(chk - base) / base)))
