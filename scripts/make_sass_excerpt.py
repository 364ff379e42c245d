"""Writes profiles/r02_step_loop.sass from the built library: opcode mix of the step loop of every build of the quad
kernel, library-wide counts of the TMA / convergence opcodes, and the full listing of the headline build's loop."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "simplyp_b200", "lib", "libsimplyp_b200.so")
mix = os.path.join(ROOT, "scripts", "sass_loop_mix.py")
out = ["# Step loop (one lock-step Runge-Kutta attempt of a warp = 8 quads) of the quad kernel, round 2 final build.",
       "# Produced by scripts/make_sass_excerpt.py (scripts/sass_loop_mix.py on the built library, i.e. cuobjdump -sass).",
       "# Round 1 (commit c43ba5c): 673 instructions in this loop for <MODE_CAL,2,false> (IMAD 96 + MOV 15, mostly register",
       "# copies behind the shuffles; 88 SHFL; 2 BRA.DIV divergence checks per iteration), 739 for the 128-register build,",
       "# 1385 for the network build.  Round 2: no BRA.DIV anywhere in the library (routing loops compile-time per build,",
       "# warp-uniform trip counts, votes instead of thread-varying conditions), six quad broadcasts per evaluation (76 SHFL),",
       "# exponent add on the high word, no division subroutine in any day loop."]
for name, pat in (("<MODE_CAL, MINB=2, STIFF=false>  (ensembles of at most 2 blocks per SM: <= 9,472 members)", "kernelILi1ELi2ELb0"),
                  ("<MODE_CAL, MINB=3, STIFF=false>  (headline: 10^4-member ensembles, 3 resident blocks on 17 SMs; up to 9 blocks per SM)", "kernelILi1ELi3ELb0"),
                  ("<MODE_CAL, MINB=4, STIFF=false>  (larger ensembles)", "kernelILi1ELi4ELb0"),
                  ("<MODE_RUN, MINB=2, STIFF=false>  (full output, one sub-catchment)", "kernelILi0ELi2ELb0"),
                  ("<MODE_RUN, MINB=2, STIFF=true>   (networks: explicit + Rosenbrock paths)", "kernelILi0ELi2ELb1"),
                  ("<MODE_RUN, MINB=3, STIFF=true>   (networks whose epochs fill the machine six times over)", "kernelILi0ELi3ELb1"),
                  ("<MODE_CAL, MINB=2, STIFF=true>", "kernelILi1ELi2ELb1")):
    r = subprocess.run([sys.executable, mix, lib, pat], capture_output=True, text=True).stdout.strip().splitlines()
    out.append("\n## " + name)
    out += r[1:]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
out.append("\n## whole library: UBLKCP (cp.async.bulk) %d, SYNCS.ARRIVE.TRANS64 %d, SYNCS.PHASECHK %d, BRA.DIV %d, SHFL %d, STG %d"
           % (sass.count("UBLKCP"), sass.count("SYNCS.ARRIVE.TRANS64"), sass.count("SYNCS.PHASECHK"), sass.count("BRA.DIV"),
              sass.count("SHFL."), sass.count("STG.")))
r = subprocess.run([sys.executable, mix, lib, "kernelILi1ELi3ELb0", "--dump"], capture_output=True, text=True).stdout.splitlines()
out.append("\n## full listing of the step loop of <MODE_CAL, 3, false>")
out += r[3:]
open(os.path.join(ROOT, "profiles", "r02_step_loop.sass"), "w").write("\n".join(out) + "\n")
print("\n".join(out[7:30]))
