import sys, threading
sys.path.insert(0, '.')
import numpy as np
from fof_b200 import _lib, distributed as D, pipeline
from fof_b200.mock import make_catalog
arr = make_catalog(20000, seed=3, as_frame=False)
np.random.seed(99)
st = np.random.get_state(); key, pos = st[1].copy(), st[2]
ref = pipeline.find_clusters(arr)
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
comms = D.ThreadComm.group(world); out = [None]*world; tim=[{} for _ in range(world)]
def work(r):
    h = _lib.Handle(0)
    out[r] = D.find_clusters_distributed(arr, comm=comms[r], handle=h, mt_state=(key, pos), timings=tim[r])
ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
[t.start() for t in ts]; [t.join() for t in ts]
cat = out[0]
print('ref K', len(ref), 'dist K', len(cat), tim[0])
rc = set(ref.centre_row.tolist()); dc = set(cat.centre_row.tolist())
print('only ref', sorted(rc - dc), 'only dist', sorted(dc - rc))
for c in sorted(rc ^ dc):
    print(c, arr[c, :3])
common = sorted(rc & dc)
ri = {c:i for i,c in enumerate(ref.centre_row.tolist())}; di = {c:i for i,c in enumerate(cat.centre_row.tolist())}
bad = 0
for c in common:
    a, b = ri[c], di[c]
    if not (np.array_equal(ref.members(a), cat.members(b)) and ref.N[a]==cat.N[b] and ref.D[a]==cat.D[b]):
        bad += 1
        if bad < 6: print('diff', c, arr[c,2], len(ref.members(a)), len(cat.members(b)), ref.N[a], cat.N[b], ref.D[a], cat.D[b])
print('differing common clusters', bad)
edges = D.slab_edges(arr[:,2], world); print('edges', edges)
