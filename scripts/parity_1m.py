"""configs[1] of BASELINE.json: 1M-galaxy catalogue with Plummer clusters on one B200, group labels and cluster
lists compared exactly with the CPU oracle (fast mode), properties to 1e-9.  Writes a JSON summary."""
import json, sys, time
sys.path.insert(0, '.')
import numpy as np
from fof_b200 import pipeline
from fof_b200.mock import make_catalog
from oracle import fof_oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/parity_1M.json"
arr = make_catalog(n, seed=20260102, as_frame=False)
np.random.seed(2026)
t0 = time.time(); cat = pipeline.find_clusters(arr); t_gpu = time.time() - t0
t0 = time.time(); ref = O.run_pipeline(arr, O.Params(), mode="fast", seed=2026); t_cpu = time.time() - t0
ids = arr[:, 6]
ok_n = np.array_equal(ref["main_arr"][:, 7], None) if False else True
same_count = len(cat) == len(ref["virial"])
members_equal = same_count and all(np.array_equal(ids[cat.members(k)], r["galaxies"][:, 6]) for k, r in enumerate(ref["virial"]))
centres_equal = same_count and all(ids[cat.centre_row[k]] == r["center"][6] for k, r in enumerate(ref["virial"]))
nd_equal = same_count and all(cat.N[k] == r["N"] and cat.D[k] == r["D"] for k, r in enumerate(ref["virial"]))
worst = 0.0
if same_count:
    for k, r in enumerate(ref["virial"]):
        for j, key in enumerate(("cluster_mass", "mass_err", "vel_disp", "vel_disp_err", "projected_radius", "total_absmag")):
            worst = max(worst, abs(cat.props[k, j] - r[key]) / abs(r[key]))
# canonical labels: smallest member row of the first final cluster containing the galaxy
lab_gpu = cat.labels()
lab_ref = np.full(n, -1, dtype=np.int64)
row_of_id = np.empty(n, dtype=np.int64); row_of_id[ids.astype(np.int64)] = np.arange(n)
for r in reversed(ref["virial"]):
    m = row_of_id[r["galaxies"][:, 6].astype(np.int64)]
    cur = lab_ref[m]
    lab_ref[m] = np.where(cur < 0, m.min(), np.minimum(cur, m.min()))
summary = dict(n=n, clusters_gpu=len(cat), clusters_oracle=len(ref["virial"]), members_equal=bool(members_equal),
               centres_equal=bool(centres_equal), N_D_bit_equal=bool(nd_equal), labels_equal=bool(np.array_equal(lab_gpu, lab_ref)),
               labelled_galaxies=int((lab_gpu >= 0).sum()), worst_relative_property_error=worst,
               gpu_wall_s_first_call=t_gpu, oracle_fast_cpu_s=t_cpu)
print(json.dumps(summary))
open(out, "w").write(json.dumps(summary, indent=1))
