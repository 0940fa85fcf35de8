"""configs[2] (linking_length_factor sweep at 10M) and configs[4] (dense percolating blobs) timings."""
import json, sys, time
sys.path.insert(0, '.')
import numpy as np
from fof_b200 import _lib, transitive
from fof_b200.mock import make_catalog
h = _lib.get_handle(0)
res = {"sweep_10M": [], "dense": [], "transitive": []}
arr = make_catalog(10_000_000, seed=20260103, as_frame=False)
for b in (0.1, 0.2, 0.3, 0.4):
    p = _lib.default_params(linking_length_factor=b)
    ts = []
    for it in range(3):
        np.random.seed(1)
        t0 = time.perf_counter(); Kf, tot = h.pipeline_run_np_random(arr, p, 0.5, 2.52, 2.5); ts.append(time.perf_counter() - t0)
    st = h.pipeline_stats()
    res["sweep_10M"].append(dict(b=b, ms_host_input=1e3 * min(ts), clusters=Kf, members=tot, **st))
    print(res["sweep_10M"][-1], flush=True)
for ll in (0.3, 0.5):
    ts = []
    for it in range(3):
        t0 = time.perf_counter(); lab, g, l = transitive.friends_of_friends(arr, linking_length=ll); ts.append(time.perf_counter() - t0)
    res["transitive"].append(dict(catalogue="10M cosmos", ll=ll, ms_host_input=1e3 * min(ts), device_ms=h.last_timing()[0], groups=g, links=l,
                                  largest=int(np.unique(lab, return_counts=True)[1].max())))
    print(res["transitive"][-1], flush=True)
del arr
for n, blobs in ((2_000_000, 12), (10_000_000, 20)):
    d = make_catalog(n, seed=77, kind="dense", n_blobs=blobs, as_frame=False)
    ts = []
    for it in range(2):
        t0 = time.perf_counter(); lab, g, l = transitive.friends_of_friends(d, linking_length=1.0); ts.append(time.perf_counter() - t0)
    res["transitive"].append(dict(catalogue=f"{n} dense, {blobs} blobs", ll=1.0, ms_host_input=1e3 * min(ts), device_ms=h.last_timing()[0],
                                  groups=g, links=l, largest=int(np.unique(lab, return_counts=True)[1].max())))
    print(res["transitive"][-1], flush=True)
    if n <= 2_000_000:
        p = _lib.default_params()
        ts = []
        for it in range(2):
            np.random.seed(1)
            t0 = time.perf_counter(); Kf, tot = h.pipeline_run_np_random(d, p, 0.5, 2.52, 2.5); ts.append(time.perf_counter() - t0)
        res["dense"].append(dict(n=n, blobs=blobs, ms_host_input=1e3 * min(ts), clusters=Kf, members=tot, **h.pipeline_stats()))
        print(res["dense"][-1], flush=True)
open("gpurun_out/sweep_and_stress.json", "w").write(json.dumps(res, indent=1))
