"""Catalogue matching on the GPU (fofb_match_catalog behind fof_b200.catalog_matching.compare_clusters) against the
reference's golden vectors and the oracle (catalog_matching.py:21-111)."""
import glob
import os

import numpy as np
import pytest

from fof_b200 import catalog_matching as cm
from oracle import fof_oracle as fo

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "matching", "*.npz")))


class _C:
    def __init__(self, row):
        self.ra, self.dec, self.z, self.richness = row


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_gpu_matches_reference_golden(path):
    g = np.load(path)
    clusters = [_C(r) for r in g["sample"]]
    matched, idx = cm.compare_clusters(clusters, g["catalog"], float(g["matching_distance"]))
    assert np.array_equal(idx, g["catalog_matching_idx"])
    assert np.array_equal(matched, g["sample_matched"])


def _random_case(seed, ns, nc):
    rng = np.random.default_rng(seed)
    sample = np.column_stack([150 + rng.uniform(-3, 3, ns), 2 + rng.uniform(-3, 3, ns), rng.uniform(0.5, 2.0, ns),
                              rng.integers(12, 90, ns).astype(float)])
    cat = np.column_stack([150 + rng.uniform(-3, 3, nc), 2 + rng.uniform(-3, 3, nc), rng.uniform(0.3, 2.3, nc),
                           rng.integers(5, 100, nc).astype(float)])
    k = min(ns, nc) // 2
    cat[:k, :2] = sample[:k, :2] + rng.normal(0, 0.005, (k, 2))
    cat[:k, 2] = sample[:k, 2] + rng.normal(0, 0.04, k)
    return sample, cat


@pytest.mark.parametrize("seed,ns,nc", [(1, 3000, 2000), (2, 17, 900), (3, 5000, 33)])
def test_gpu_matches_oracle(seed, ns, nc):
    sample, cat = _random_case(seed, ns, nc)
    want_m, want_idx, kept, want_row, want_sep = fo.compare_clusters(sample, cat, 0.5, detail=True)
    rows, sep, n = cm.match_rows(sample, cat, 0.5)
    assert np.array_equal(np.flatnonzero(rows != -2), kept)
    assert np.array_equal(rows[kept], want_row)          # index work: bit-exact
    assert n == int((want_row >= 0).sum())
    both = ~np.isnan(want_sep)
    assert np.array_equal(np.isnan(sep[kept]), ~both)
    assert np.allclose(sep[kept][both], want_sep[both], rtol=1e-9, atol=0)   # float tolerance stated: 1e-9 relative
    got_m, got_idx = cm.compare_clusters(sample, cat, 0.5)
    assert np.array_equal(got_idx, want_idx) and np.array_equal(got_m, want_m)


def test_duplicate_sample_rows_pick_lowest_row():
    sample = np.array([[150.0, 2.0, 1.0, 20.0]] * 5 + [[150.2, 2.1, 1.0, 30.0]])
    cat = np.array([[150.0001, 2.0, 1.01, 9.0]])
    rows, sep, n = cm.match_rows(sample, cat, 0.5)
    assert rows.tolist() == [0] and n == 1      # pinned order D5


def test_large_sample_properties():
    # 2M found clusters against 50k catalogue rows: every catalogue row planted on a sample row must find it
    rng = np.random.default_rng(9)
    ns, nc = 2_000_000, 50_000
    sample = np.column_stack([rng.uniform(100, 150, ns), rng.uniform(-20, 20, ns), rng.uniform(0.5, 2.0, ns),
                              rng.integers(12, 90, ns).astype(float)])
    pick = rng.choice(ns, nc, replace=False)
    cat = sample[pick].copy()
    cat[:, 2] += rng.uniform(-0.05, 0.05, nc)
    rows, sep, n = cm.match_rows(sample, cat, 0.5)
    inside = rows != -2
    assert inside.sum() >= nc - 10
    assert np.array_equal(rows[inside], pick[inside]) and n == inside.sum()
    assert np.all(sep[inside] < 1e-12)


def test_empty_inputs():
    sample, cat = _random_case(4, 50, 20)
    rows, sep, n = cm.match_rows(sample, cat[:0], 0.5)
    assert rows.size == 0 and n == 0
    with pytest.raises(ValueError):
        cm.compare_clusters(sample[:0], cat, 0.5)
