"""Whole path at the BASELINE.json sizes against the committed oracle fixtures (scripts/make_parity_fixtures.py):
configs[0] 100k galaxies, configs[1] 1M galaxies with Plummer clusters, and a 100k table with photo-z style duplicated
z_upper values.  Bit-exact cluster lists, member order, centres, N, D and canonical labels; properties to 1e-9.
The table is regenerated from the stored seed; the oracle itself does not run here (it needs 14 minutes at 1M)."""
import glob
import os
import sys

import numpy as np
import pytest

from fof_b200.mock import make_catalog

pytestmark = pytest.mark.gpu
SIZES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sizes")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))


def _labels(n, off, rows):
    """Canonical label per row: smallest member row of the final clusters containing it, -1 outside."""
    lab = np.full(n, n, dtype=np.int64)
    sizes = np.diff(off)
    mins = np.minimum.reduceat(rows, off[:-1]) if len(sizes) else np.zeros(0, dtype=np.int64)
    np.minimum.at(lab, rows, np.repeat(mins, sizes))
    lab[lab == n] = -1
    return lab


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(SIZES, "*.npz"))), ids=lambda p: os.path.basename(p)[:-4])
def test_whole_path_matches_oracle_fixture(lib, handle, path):
    from fof_b200 import pipeline
    from make_parity_fixtures import n_checksum
    f = np.load(path)
    n, dec = int(f["n"]), int(f["photoz_decimals"])
    arr = make_catalog(n, seed=int(f["seed"]), as_frame=False, **({"photoz_decimals": dec} if dec >= 0 else {}))
    np.random.seed(int(f["rng_seed"]))
    cat = pipeline.find_clusters(arr)
    # stage 0: the N(0.5) column and the number of candidate centres
    assert int(cat.N_row.sum()) == int(f["N_sum"])
    assert n_checksum(cat.N_row.astype(np.float64)) == int(f["N_checksum"])
    st = handle.pipeline_stats()
    assert st["candidate_centres"] == int(f["n_candidates"])
    assert st["stage4_clusters"] >= int(f["n_stage4"])   # the fixture counts FoF()'s output, i.e. after the N / D cuts
    # final catalogue: clusters in order, members in the reference's order
    ids = arr[:, 6].astype(np.int64)
    assert np.array_equal(cat.offsets, f["off"])
    assert np.array_equal(ids[cat.member_rows], f["member_ids"].astype(np.int64))
    assert np.array_equal(ids[cat.centre_row], f["centre_ids"])
    assert np.array_equal(cat.N, f["N"])
    assert np.array_equal(cat.D, f["D"], equal_nan=True)
    np.testing.assert_allclose(cat.props[:, :6], f["props"], rtol=1e-9, atol=0)
    # group labels after canonicalisation (min member row)
    row_of_id = np.empty(n, dtype=np.int64)
    row_of_id[ids] = np.arange(n)
    want = _labels(n, f["off"], row_of_id[f["member_ids"].astype(np.int64)])
    assert np.array_equal(cat.labels(), want)
    assert np.array_equal(_labels(n, cat.offsets, cat.member_rows.astype(np.int64)), want)
