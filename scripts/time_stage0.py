import sys, time
sys.path.insert(0, '.')
import numpy as np
from fof_b200 import _lib
from fof_b200.mock import make_catalog
h = _lib.get_handle(0)
p = _lib.default_params()
for n in [int(x) for x in sys.argv[1:]]:
    t = time.time(); arr = make_catalog(n, seed=5, as_frame=False); tg = time.time() - t
    for it in range(3):
        t = time.time(); N, rows = h.number_counts(arr, p); tw = time.time() - t
        ms, k = h.last_timing()
        print(f"n={n} gen={tg:.1f}s wall={tw*1e3:.1f}ms device={ms:.2f}ms launches={k} cand={len(rows)} Nmax={N.max()} gal/s(device)={n/ms*1e3:.3e}", flush=True)
