"""Full-output runs of synthetic reach networks (BASELINE configs 3 and 5, reduced): time and throughput."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from simplyp_b200 import _cabi, model as spm, packing as pk, synthetic, tarland, inputs
from simplyp_b200.engine import Engine

def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    years = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    M = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    p_SU, dyn, p, p_LU, p_SC0, p_struc0, met, obs = tarland.load(dynamic="y")
    p, p_SC, p_struc = synthetic.random_network(p, p_SC0[1], n_sc=S, seed=3, all_land_uses=(S >= 1024))
    met = inputs.snow_hydrol_inputs(0.0, 2.74, synthetic.synthetic_met(int(365.25 * years)))
    pk.validate_land_use(p_SC, p["SC_list"])
    topo = pk.build_topology(p_struc, p["SC_list"])
    nlev, lv = _cabi.topology_levels(topo.parent_offsets, topo.parent_ids)
    opt = spm.make_options(p_SU, p, dyn, topo)
    eng = Engine(0)
    forcing = eng.to_device(pk.forcing_matrix(met))
    member = eng.to_device(np.repeat(pk.member_vector(p, p_LU)[None], M, axis=0))
    sc = eng.to_device(pk.sc_matrix(p_SC, topo.sc_ids)[None])
    D = forcing.shape[0]
    out = torch.empty((M, S, D, 25), dtype=torch.float64, device=eng.device)
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out, dg = eng.run(forcing, member, sc, topo.parent_offsets, topo.parent_ids, opt, out=out)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    steps = dg[..., 0].double().sum().item()
    print(json.dumps({"S": S, "D": D, "M": M, "levels": nlev, "seconds": dt, "sc_days_per_s": M * S * D / dt,
                      "out_GB": out.numel() * 8 / 1e9, "write_GBps": out.numel() * 8 / dt / 1e9,
                      "steps_per_sc_day": steps / (M * S * D), "status": int(dg[..., 3].max().item()),
                      "finite": bool(torch.isfinite(out).all().item())}), flush=True)

if __name__ == "__main__":
    main()
