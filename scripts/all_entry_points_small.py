"""Small end-to-end run of every device entry point (pipeline, sweep, unsorted rows, both union-find kernels, the
generator, catalogue matching, the giant-group paths): a quick whole-surface check on a GPU box.
    python scripts/all_entry_points_small.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
from fof_b200 import _lib, pipeline, transitive, catalog_matching, fof, mass_analysis
from fof_b200.cluster import Cluster
from fof_b200.mock import make_catalog

arr = make_catalog(4000, seed=3, as_frame=False)
np.random.seed(7)
cat = pipeline.find_clusters(arr, richness=8, overdensity=0.0)
print("pipeline clusters", len(cat))
cats = pipeline.sweep(arr, [0.1, 0.3], richness=8, overdensity=0.0)
print("sweep", {b: len(c) for b, c in cats.items()})
shuffled = arr[np.random.default_rng(1).permutation(len(arr))]
np.random.seed(7)
print("shuffled", len(pipeline.find_clusters(shuffled, richness=8, overdensity=0.0)))
lab, g, l = transitive.friends_of_friends(arr, linking_length=0.5)
lab2, g2, _ = transitive.friends_of_friends(arr, linking_length=0.5, count_links=False)
assert np.array_equal(lab, lab2)
dense = make_catalog(6000, seed=4, kind="dense", n_blobs=2, as_frame=False)
a, ga, _ = transitive.friends_of_friends(dense, linking_length=1.0)
b, gb, _ = transitive.friends_of_friends(dense, linking_length=1.0, count_links=False)
assert np.array_equal(a, b)
print("transitive", g, ga)
h = _lib.get_handle(0)
w, st = h.mt19937_words(3 * 2048 * 624 + 11, state=np.random.RandomState(5).get_state())
print("mt words", len(w))
sample = np.column_stack([arr[:300, 0], arr[:300, 1], arr[:300, 2], np.full(300, 20.0)])
rows, sep, n = catalog_matching.match_rows(sample, sample[::3] + [1e-4, 0, 0.01, 0], 0.5)
print("matched", n)
# a group with long bins and more than 4096 members: batched clipping and the grid-wide pair sum
rng = np.random.default_rng(9)
n = 5000
z = 1.0 + rng.normal(0, 800, n) / 299792.458 * 2
gg = np.column_stack([150 + rng.normal(0, 0.01, n), 2 + rng.normal(0, 0.01, n), z, z - 0.01 - 1e-7 * np.arange(n),
                      z + 0.01 + 1e-7 * np.arange(n), rng.normal(-21, 1, n), np.arange(n, dtype=float), np.full(n, 30.0)])
gg[0, :3] = (150, 2, 1.0)
cleaned, removed = fof.interloper_removal([Cluster(gg[0], gg.copy())])
v = mass_analysis.find_masses(cleaned, "virial")[0]
print("giant group", removed, float(v.cluster_mass) > 0)
# photo-z style columns (duplicated z_upper: the by-value tracker) and a field dense enough that strip neighbourhoods
# exceed the shared-memory tile (the global-memory path of the tiled counts, the larger hop-2 queues)
pz = make_catalog(4000, seed=5, photoz_decimals=2, as_frame=False)
np.random.seed(7)
print("photo-z", len(pipeline.find_clusters(pz, richness=8, overdensity=0.0)))
crowded = make_catalog(12000, seed=6, density=3.0e6, richness_range=(200, 900), cluster_fraction=0.6, as_frame=False)
np.random.seed(7)
print("crowded", len(pipeline.find_clusters(crowded, linking_length_factor=0.05, richness=8, overdensity=0.0)))
print("all entry points ran")
