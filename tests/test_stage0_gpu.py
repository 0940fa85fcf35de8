"""GPU parity of stage 0 (candidate_center_search, fof.py:24-95) against the oracle."""
import numpy as np
import pytest

from fof_b200.mock import make_catalog
from oracle import fof_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,seed,kind", [(3000, 1, "cosmos"), (20000, 2, "cosmos"), (20000, 3, "wide"),
                                          (100000, 4, "cosmos")])
def test_number_counts_match_oracle(lib, handle, n, seed, kind):
    df = make_catalog(n, seed=seed, kind=kind)
    arr = df.values
    p = lib.default_params()
    N, rows = handle.number_counts(arr, p)
    main_arr, cand = O.candidate_center_search(arr, O.Params(), mode="fast")
    assert np.array_equal(N.astype(np.float64), main_arr[:, 7])
    assert np.array_equal(arr[rows, 6], cand[:, 6])  # same centres in the same (N desc, row desc) order
