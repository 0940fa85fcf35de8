import sys, time; sys.path.insert(0, ".")
import numpy as np
from fof_b200 import pipeline, _lib
from fof_b200.mock import make_catalog
import os
for n, blobs in ((50_000, 3), (200_000, 4)) + (((1_000_000, 8),) if os.environ.get("DENSE_FULL") else ()):
    d = make_catalog(n, seed=5, kind="dense", n_blobs=blobs, as_frame=False)
    np.random.seed(3)
    h = _lib.get_handle(0)
    h.profile(2)
    t0 = time.perf_counter()
    cat = pipeline.find_clusters(d)
    dt = time.perf_counter() - t0
    h.profile(0)
    prof = h.profile_read()
    top = sorted(prof.items(), key=lambda kv: -kv[1][0])[:4]
    sizes = np.diff(cat.offsets) if len(cat) else np.zeros(1)
    print(n, blobs, "ms", round(1e3 * dt, 1), "clusters", len(cat), "largest", int(sizes.max()), [(k, round(v[0], 1)) for k, v in top], flush=True)
    print("   stats", h.pipeline_stats(), {k: (round(v[0], 1), v[1]) for k, v in prof.items() if k.startswith("k_hop2") or k.startswith("k_virial")}, flush=True)
