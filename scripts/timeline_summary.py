"""Summary of a timeline written by scripts/exp_timeline.py: SM end times, the SMs that end last and their blocks
(virtual block, start-end in ms, lock-step attempts of its slowest warp).  usage: timeline_summary.py gpurun_out/timeline_10000.npz"""
import sys, numpy as np
z = np.load(sys.argv[1])
steps, t0, t1, smid, vb, warp = (z[k].astype(np.int64) for k in ("lockstep", "t0", "t1", "smid", "vblock", "warp"))
T0 = t0.min()
print("ms", float(z["ms"]), "span %.3f ms" % ((t1.max() - T0) * 1e-6), "SMs", len(np.unique(smid)))
# per block: start, end (max over its warps), SM
blocks = {}
for b in np.unique(vb):
    sel = vb == b
    blocks[b] = (int(smid[sel][0]), (t0[sel].min() - T0) * 1e-6, (t1[sel].max() - T0) * 1e-6, steps[sel].max())
starts = np.array([blocks[b][1] for b in sorted(blocks)]); ends = np.array([blocks[b][2] for b in sorted(blocks)])
print("blocks", len(blocks), "started after 0.1 ms:", int((starts > 0.1).sum()), "latest start %.2f ms" % starts.max())
# per SM: end time, blocks
sm_end = {}
for b, (s, a, e, st) in blocks.items():
    sm_end.setdefault(s, []).append((b, a, e, st))
ends_sm = np.array([max(x[2] for x in v) for v in sm_end.values()])
print("SM end times (ms): min %.2f  p25 %.2f  median %.2f  p75 %.2f  max %.2f" % tuple(np.percentile(ends_sm, [0, 25, 50, 75, 100])))
worst = sorted(sm_end.items(), key=lambda kv: -max(x[2] for x in kv[1]))[:4] + sorted([kv for kv in sm_end.items() if len(kv[1]) == 2], key=lambda kv: -max(x[2] for x in kv[1]))[:5]
for s, v in worst:
    print("SM", s, [(int(b), "%.2f-%.2f" % (a, e), int(st)) for b, a, e, st in sorted(v, key=lambda x: x[1])])
# warp-level: rate of the last-finishing warps: time / lock-step attempts is unknown; use max member steps in warp as a proxy
wkey = vb * 8 + warp
last = np.argsort(-t1)[:5]
for i in last:
    sel = wkey == wkey[i]
    print("late warp: block %d warp %d SM %d ends %.2f ms, member steps max %d  -> >= %.0f cycles/attempt" %
          (vb[i], warp[i], smid[i], (t1[i] - T0) * 1e-6, steps[sel].max(), (t1[i] - t0[i]) * 1.965 / steps[sel].max()))
