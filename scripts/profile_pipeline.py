"""One whole-path step at n galaxies with kernel event profiling; prints per-kernel times and size stats."""
import sys
sys.path.insert(0, '.')
import numpy as np
from fof_b200 import _lib
from fof_b200.mock import make_catalog
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
h = _lib.get_handle(0)
p = _lib.default_params()
arr = make_catalog(n, seed=20260103, as_frame=False)
np.random.seed(1)
for it in range(reps):
    h.profile(2)
    Kf, tot = h.pipeline_run_np_random(arr, p, 0.5, 2.52, 2.5)
    h.profile(0)
    res = h.pipeline_fetch(Kf, tot)
prof = h.profile_read()
tot_ms = sum(v[0] for v in prof.values())
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print(f"{v[0]:9.3f} ms {100*v[0]/tot_ms:5.1f}% x{v[1]:3d} {k}")
print("device ms", h.last_timing(), h.pipeline_stats())
sizes = np.diff(res[0])
print("final clusters", len(sizes), "size mean/max/p99", sizes.mean(), sizes.max(), np.percentile(sizes, 99))
