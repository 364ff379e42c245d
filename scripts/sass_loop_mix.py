"""Opcode histogram of the step loop (backward branch span) of a kernel in a built library.
usage: python scripts/sass_loop_mix.py <lib.so> <kernel-name-substring> [--dump]"""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs, cur = {}, None
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name:
        continue
    for i, (a, t) in enumerate(ins):      # blocks after the first EXIT are the cold divergent-collective paths
        if t.startswith("EXIT"):
            ins = ins[:i + 1]
            break
    best = None   # the smallest backward-branch span that holds at least 25 % of the function's DFMAs
    n_shfl = sum(1 for a, s in ins if "DFMA" in s)
    for a, s in ins:
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\w+,\s*)?(0x[0-9a-f]+)", s)
        if m:
            t = int(m.group(1), 16)
            if t < a:
                k = sum(1 for b, s2 in ins if t <= b <= a and "DFMA" in s2)
                if 4 * k >= n_shfl and (best is None or a - t < best[1] - best[0]):
                    best = (t, a)
    body = [s for a, s in ins if best and best[0] <= a <= best[1]]
    c = collections.Counter()
    for s in body:
        op = s.split()[1] if s.startswith("@") else s.split()[0]
        c[op.split(".")[0]] += 1
    fp64 = sum(c[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
    print("%s\n  total %d instrs in function; step loop %d instrs, fp64-pipe %d" % (name[-60:], len(ins), len(body), fp64))
    print("  " + ", ".join("%s %d" % kv for kv in c.most_common(18)))
    if "--dump" in sys.argv:
        print("\n".join(body))
