"""Stage artefacts in the reference's formats (fof_b200/io.py): pickles written by the unmodified reference classes load
into this package's classes (pipeline.py:103-128), the FITS property export of helper.py:224-237 round-trips, and the
columnar property table equals the per-object one."""
import os

import numpy as np
import pytest

from fof_b200 import io as fio
from fof_b200.cluster import CleanedCluster, Cluster
from fof_b200.pipeline import Catalogue
from fof_b200.units import Q

IO_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "io")


def test_reference_pickle_loads_into_our_classes():
    clusters = fio.load_clusters(os.path.join(IO_DIR, "ref_cleaned_candidates.dat"))
    want = np.load(os.path.join(IO_DIR, "ref_cleaned_candidates_expected.npz"))
    assert len(clusters) == len(want["centre"]) > 0
    assert all(type(c) is Cluster for c in clusters)
    assert np.array_equal(np.array([c.center_attributes for c in clusters]), want["centre"])
    assert np.array_equal(np.array([c.D for c in clusters]), want["D"])
    assert np.array_equal(np.array([c.flag_poor for c in clusters]), want["flag_poor"])
    off = want["offsets"]
    for k, c in enumerate(clusters):
        assert np.array_equal(c.galaxies, want["members"][off[k]:off[k + 1]])
        assert c.richness == off[k + 1] - off[k]


def _cleaned(k, rng):
    n = 12 + k
    gal = np.column_stack([150 + rng.normal(0, .01, n), 2 + rng.normal(0, .01, n), 1 + rng.normal(0, .002, n),
                           np.zeros(n), np.zeros(n), rng.normal(-21, 1, n), np.arange(n) + 100. * k, np.full(n, 9.)])
    c = CleanedCluster(gal[0], gal, Q(1e14 * (k + 1), "Msun/littleh"), Q(1e13, "Msun/littleh"), Q(4e5, "km2/s2"),
                       Q(3e4, "km2/s2"), Q(0.8, "Mpc/littleh"), -24.5 - k)
    c.D = 2.5 + k
    return c


def test_dump_and_load_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    clusters = [_cleaned(k, rng) for k in range(4)]
    path = str(tmp_path / "clusters.dat")
    fio.dump_clusters(clusters, path)
    back = fio.load_clusters(path)
    assert [type(c) for c in back] == [CleanedCluster] * 4
    for a, b in zip(clusters, back):
        assert np.array_equal(a.properties, b.properties) and np.array_equal(a.galaxies, b.galaxies)
        assert b.cluster_mass.value == a.cluster_mass.value


def test_fits_export_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    clusters = [_cleaned(k, rng) for k in range(7)]
    arr = fio.export(clusters, "sample", path=str(tmp_path) + os.sep)
    assert arr.shape == (7, 11) and np.array_equal(arr[3], clusters[3].properties)
    path = str(tmp_path / "sample.fits")
    raw = open(path, "rb").read()
    assert len(raw) % 2880 == 0 and raw.startswith(b"SIMPLE  =                    T")
    assert np.array_equal(fio.read_fits_array(path), arr)
    with pytest.raises(OSError):
        fio.export(clusters, "sample", path=str(tmp_path) + os.sep)   # astropy's writeto refuses to overwrite


def test_properties_table_equals_per_object_properties():
    rng = np.random.default_rng(2)
    n, K = 200, 5
    g = np.column_stack([150 + rng.uniform(-1, 1, n), 2 + rng.uniform(-1, 1, n), rng.uniform(.5, 2, n), np.zeros(n), np.zeros(n),
                         rng.normal(-21, 1, n), np.arange(n, dtype=float)])
    sizes = rng.integers(12, 30, K)
    off = np.concatenate([[0], np.cumsum(sizes)])
    rows = rng.integers(0, n, off[-1])
    cat = Catalogue(g, off, rows, rng.integers(0, n, K), rng.integers(8, 40, K).astype(float), rng.uniform(2, 9, K),
                    rng.uniform(1, 10, (K, 6)), N_row=rng.integers(0, 30, n).astype(np.int32))
    assert np.array_equal(cat.cluster(1).galaxies[:, 7], cat.N_row[cat.members(1)])   # rows carry their own N(0.5)
    table = fio.properties_table(cat)
    want = np.array([c.properties for c in cat.clusters()])
    assert np.array_equal(table, want)
    assert fio.properties_table(Catalogue(g, np.zeros(1, dtype=np.int64), rows[:0], rows[:0], np.zeros(0), np.zeros(0),
                                          np.zeros((0, 6)))).shape == (0, 11)
