import sys, time
sys.path.insert(0, '.')
import numpy as np
from fof_b200 import _lib, params as P
from fof_b200.mock import make_catalog
h = _lib.get_handle(0)
p = _lib.default_params()
reps = 3
for n in [int(x) for x in sys.argv[1:]]:
    arr = make_catalog(n, seed=5, as_frame=False)
    for it in range(reps):
        t0 = time.time()
        K = h.pipeline_begin(arr, p, 0.5, 2.52, 2.5)
        ms1, k1 = h.last_timing()
        t1 = time.time()
        u = np.random.random_sample((K, 2, 300))
        t2 = time.time()
        Kf, tot = h.pipeline_finish(u)
        ms2, k2 = h.last_timing()
        t3 = time.time()
        res = h.pipeline_fetch(Kf, tot)
        t4 = time.time()
        print(f"n={n} begin={1e3*(t1-t0):.1f}ms(dev {ms1:.1f}) rng={1e3*(t2-t1):.1f}ms finish={1e3*(t3-t2):.1f}ms(dev {ms2-ms1:.1f}) fetch={1e3*(t4-t3):.1f}ms total={1e3*(t4-t0):.1f}ms K'={K} Kf={Kf} members={tot} launches={k2} stats={h.pipeline_stats()}", flush=True)
