"""Small case for compute-sanitizer (racecheck / memcheck): 5-reach network x 24 members x 300 days swept in 3 epochs of
128 days (forcing tiles, epoch hand-over), full output + calibration with rank statistics, a one-sub-catchment ensemble
with the cost pilot, and a 9,600-member ensemble x 70 days under the placement plan with 3 resident blocks per SM."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simplyp_b200 import _cabi, ensemble as ens, model as spm, packing as pk, tarland
from tests.golden.networks import network5_inputs
p_SU, dyn, p, p_LU, p_SC, p_struc, met, obs = tarland.load(dynamic="y")
met = met.iloc[:300]
os.environ["SIMPLYP_EPOCH_DAYS"] = "128"
p5, p_LU5, p_SC5, p_struc5 = network5_inputs(p, p_LU, p_SC, p_struc)
pk.validate_land_use(p_SC5, p5["SC_list"])
topo = pk.build_topology(p_struc5, p5["SC_list"])
opt = spm.make_options(p_SU, p5, dyn, topo)
samples = ens.latin_hypercube(24, seed=2)
member, sc = ens.pack_members(pk.member_vector(p5, p_LU5), pk.sc_matrix(p_SC5, topo.sc_ids), samples)
forcing = pk.forcing_matrix(met)
out, dg = _cabi.run_host(forcing, member, sc, topo.parent_offsets, topo.parent_ids, opt)
assert np.all(np.isfinite(out)) and not np.any(dg[..., 3])
obs_m, desc, labels = pk.obs_arrays({5: obs[1]}, topo, met.index, ("Q", "TDP"))
opt.rank_stats = 1
st, _ = _cabi.calibrate_host(forcing, member, sc, topo.parent_offsets, topo.parent_ids, obs_m, desc, opt)
assert np.all(np.isfinite(st[..., :9]))
topo1 = pk.build_topology(p_struc, p["SC_list"])
opt1 = spm.make_options(p_SU, p, dyn, topo1)
samples = ens.latin_hypercube(600, seed=4)
member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo1.sc_ids), samples)
obs_m, desc, labels = pk.obs_arrays(obs, topo1, met.index, ("Q", "TDP"))
st, dg = _cabi.calibrate_host(forcing, member, sc, topo1.parent_offsets, topo1.parent_ids, obs_m, desc, opt1)
assert np.all(np.isfinite(st[..., :8])) and not np.any(dg[..., 3])
samples = ens.latin_hypercube(9600, seed=6)
member, sc = ens.pack_members(pk.member_vector(p, p_LU), pk.sc_matrix(p_SC, topo1.sc_ids), samples)
st, dg = _cabi.calibrate_host(forcing[:70], member, sc, topo1.parent_offsets, topo1.parent_ids, obs_m[:, :70], desc, opt1)
assert np.all(np.isfinite(st[..., :8])) and not np.any(dg[..., 3])
print("sanitizer case OK")
