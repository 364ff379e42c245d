"""Instruction mix of the hot loop from an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iN, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
body = rows[2:]
n = [int(r[iN] or 0) for r in body]
top = max(n)
mix = collections.Counter(); samp = collections.Counter()
tot = 0
for r, k in zip(body, n):
    op = r[iS].split()[0] if not r[iS].strip().startswith("@") else r[iS].split()[1]
    op = op.split(".")[0]
    mix[op] += k; samp[op] += int(r[iSamp] or 0); tot += k
print("total warp instructions %d ; hottest line executed %d times" % (tot, top))
ts = sum(samp.values())
for op, k in mix.most_common(24):
    print("  %-10s %6.2f %% of instructions   %6.2f %% of stall samples" % (op, 100.0 * k / tot, 100.0 * samp[op] / ts))
# per-iteration static instruction count of the step loop = lines executed at (about) the top frequency
hot = [r for r, k in zip(body, n) if k > 0.5 * top]
print("static instructions at >50%% of top frequency: %d" % len(hot))
hm = collections.Counter()
for r in hot:
    src = r[iS].strip()
    op = src.split()[1] if src.startswith("@") else src.split()[0]
    hm[op.split(".")[0]] += 1
print("  " + ", ".join("%s %d" % kv for kv in hm.most_common(20)))
