"""Redshift-slab decomposition: the same catalogue on 1, 2, 3 and 4 ranks gives identical output (SURVEY.md
test 6).  Ranks are emulated by threads with one handle each on the single test GPU; the collectives go through
distributed.ThreadComm (the multi-process NCCL path is exercised by bench.py --decomposition slabs)."""
import threading

import numpy as np
import pytest

from fof_b200.mock import make_catalog

pytestmark = pytest.mark.gpu


def _run_ranks(arr, world, state):
    from fof_b200 import _lib, distributed as D
    comms = D.ThreadComm.group(world)
    out = [None] * world
    err = []

    def work(r):
        try:
            h = _lib.Handle(0)
            out[r] = D.find_clusters_distributed(arr, comm=comms[r], handle=h, mt_state=state)
            h.close()
        except Exception as e:  # pragma: no cover
            err.append(e)
            comms[r].s.barrier.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    if err:
        raise err[0]
    return out


@pytest.mark.parametrize("n,seed", [(20000, 3), (60000, 17)])
def test_slabs_equal_single_gpu(lib, n, seed):
    from fof_b200 import pipeline
    arr = make_catalog(n, seed=seed, as_frame=False)
    np.random.seed(99)
    state0 = np.random.get_state()
    key, pos = state0[1].copy(), state0[2]
    ref = pipeline.find_clusters(arr)
    ref_state = np.random.get_state()
    assert len(ref) > 3
    for world in (2, 3, 4):
        cats = _run_ranks(arr, world, (key, pos))
        for cat in cats:
            assert np.array_equal(cat.offsets, ref.offsets)
            assert np.array_equal(cat.member_rows, ref.member_rows)
            assert np.array_equal(cat.centre_row, ref.centre_row)
            assert np.array_equal(cat.N, ref.N) and np.array_equal(cat.D, ref.D)
            np.testing.assert_allclose(cat.props, ref.props, rtol=1e-12)
            assert np.array_equal(cat.mt_state[0], ref_state[1]) and cat.mt_state[1] == ref_state[2]
