"""Generate tests/golden/matching/*.npz by running the UNMODIFIED reference compare_clusters
(/root/reference/FoF/catalog_matching.py:21-111) against the third-party shims.  Build-container only.

    python -m oracle.make_golden_matching
"""
import importlib
import logging
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.run_reference import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "matching")


class SampleCluster:
    """The four attributes compare_clusters reads from a found cluster (catalog_matching.py:56)."""

    def __init__(self, ra, dec, z, richness):
        self.ra, self.dec, self.z, self.richness = ra, dec, z, richness


def make_case(seed, n_sample, n_catalogue, zero_richness=False):
    rng = np.random.default_rng(seed)
    sample = np.column_stack([150 + rng.uniform(-.7, .7, n_sample), 2.2 + rng.uniform(-.7, .7, n_sample),
                              rng.uniform(0.5, 2.0, n_sample), rng.integers(12, 90, n_sample).astype(float)])
    cat = np.column_stack([150 + rng.uniform(-.7, .7, n_catalogue), 2.2 + rng.uniform(-.7, .7, n_catalogue),
                           rng.uniform(0.2, 2.4, n_catalogue), rng.integers(5, 100, n_catalogue).astype(float)])
    # plant near-coincident pairs at every third catalogue row, at separations around the matching angle
    for k in range(0, min(n_catalogue, n_sample), 3):
        cat[k, 0] = sample[k, 0] + rng.normal(0, 0.006)
        cat[k, 1] = sample[k, 1] + rng.normal(0, 0.006)
        cat[k, 2] = sample[k, 2] + rng.normal(0, 0.05)
    if zero_richness:  # a matched row containing a zero is dropped by catalog_matching.py:95
        sample[0, 3] = 0.0
        cat[0, :3] = sample[0, :3]
    return sample, cat


CASES = {"match_400x300": (3, 400, 300, False), "match_60x500_zero": (11, 60, 500, True), "match_1x40": (5, 1, 40, False)}


def main():
    ref = load_reference()
    cwd = os.getcwd()
    os.chdir(ref.scratch)  # logging.basicConfig at catalog_matching.py:12 opens a file relative to the cwd
    cm = importlib.import_module("catalog_matching")
    os.chdir(cwd)
    logging.disable(logging.CRITICAL)
    u = sys.modules["astropy"].units
    os.makedirs(OUT, exist_ok=True)
    for name, (seed, ns, nc, zero) in CASES.items():
        sample, cat = make_case(seed, ns, nc, zero)
        clusters = [SampleCluster(*row) for row in sample]
        matched, idx = cm.compare_clusters(clusters, cat, 0.5 * u.Mpc / u.littleh, richness_plot=False)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), sample=sample, catalog=cat, matching_distance=0.5,
                            sample_matched=matched, catalog_matching_idx=idx)
        print(name, "catalogue rows kept", len(idx), "matched", int(idx.sum()), "rows returned", len(matched))


if __name__ == "__main__":
    main()
