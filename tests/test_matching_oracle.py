"""Catalogue matching oracle (oracle/fof_oracle.compare_clusters) against the golden vectors produced by the
unmodified reference compare_clusters (catalog_matching.py:21-111; generator: oracle/make_golden_matching.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import fof_oracle as fo

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "matching", "*.npz")))


def test_fixtures_present():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_golden(path):
    g = np.load(path)
    matched, idx = fo.compare_clusters(g["sample"], g["catalog"], float(g["matching_distance"]))
    assert np.array_equal(idx, g["catalog_matching_idx"])
    assert np.array_equal(matched, g["sample_matched"])


def test_matching_angle_known_answer():
    # (0.5/h)(1+z)/D_A(z) radians for a flat cosmology: catalog_matching.py:85-88 collapses to this
    for z in (0.5, 1.0, 2.0):
        want = np.rad2deg((0.5 / 0.7) * (1 + z) / fo.angular_diameter_distance(z))
        assert abs(fo.matching_angle(0.5, z) - want) <= 1e-13 * want


def test_redshift_filter_and_empty_cut():
    sample = np.array([[150.0, 2.0, 1.0, 20.0], [150.01, 2.0, 1.05, 30.0]])
    cat = np.array([[150.0, 2.0, 0.85, 9.0],      # dropped by the range filter (z < 1.0 - 0.1)
                    [150.0, 2.0, 1.16, 9.0],      # dropped (z > 1.05 + 0.1)
                    [150.0005, 2.0, 1.02, 9.0],   # nearest = sample 0, within the angle
                    [151.0, 2.0, 1.02, 9.0]])     # kept, nothing within the angle
    matched, idx, kept, row, sep = fo.compare_clusters(sample, cat, 0.5, detail=True)
    assert kept.tolist() == [2, 3] and idx.tolist() == [1.0, 0.0] and row.tolist() == [0, -1]
    assert np.array_equal(matched, sample[[0]])
