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
            h._mt_share_min_words = 0   # share the stream however short it is
            first = D.find_clusters_distributed(arr, comm=comms[r], handle=h, mt_state=state)
            # the second call on a handle knows the cluster count and generates the MT19937 stream in shares that the
            # ranks all-gather: the result must not change
            out[r] = D.find_clusters_distributed(arr, comm=comms[r], handle=h, mt_state=state)
            assert np.array_equal(first.member_rows, out[r].member_rows) and np.array_equal(first.D, out[r].D)
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


NCCL_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import torch
import torch.distributed as dist
rank = int(os.environ["RANK"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
from fof_b200 import _lib, distributed as D
from fof_b200.mock import make_catalog
arr = make_catalog(int(sys.argv[2]), seed=17, as_frame=False)
np.random.seed(99)
st = np.random.get_state()
handle = _lib.Handle(rank)
handle._mt_share_min_words = 0   # share the MT19937 stream on the second call however short it is
cat = D.find_clusters_distributed(arr, comm=D.TorchComm(), handle=handle, mt_state=(st[1].copy(), st[2]))
cat = D.find_clusters_distributed(arr, comm=D.TorchComm(), handle=handle, mt_state=(st[1].copy(), st[2]))  # MT stream in shares
np.savez(sys.argv[3] + f".rank{rank}.npz", off=cat.offsets, rows=cat.member_rows, centre=cat.centre_row, N=cat.N, D=cat.D,
         props=cat.props, key=cat.mt_state[0], pos=cat.mt_state[1])
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_process_nccl_slabs_equal_single_gpu(lib, tmp_path):
    """One process per GPU over NCCL (the production layout): the catalogue both ranks assemble must be byte-identical to
    the single-GPU run.  Needs two GPUs on the box (``gpurun --gpus 2``); skipped on a single-GPU box, where the
    thread-emulated ranks above cover the same driver code."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from fof_b200 import pipeline
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n = 200000
    arr = make_catalog(n, seed=17, as_frame=False)
    np.random.seed(99)
    ref = pipeline.find_clusters(arr)
    ref_state = np.random.get_state()
    script = tmp_path / "worker.py"
    script.write_text(NCCL_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), root, str(n), str(tmp_path / "out")],
                              env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert len(ref) > 100
    for r in range(2):
        got = np.load(str(tmp_path / "out") + f".rank{r}.npz")
        assert np.array_equal(got["off"], ref.offsets) and np.array_equal(got["rows"], ref.member_rows)
        assert np.array_equal(got["centre"], ref.centre_row)
        assert np.array_equal(got["N"], ref.N) and np.array_equal(got["D"], ref.D)
        np.testing.assert_allclose(got["props"], ref.props, rtol=1e-12)
        assert np.array_equal(got["key"], ref_state[1]) and int(got["pos"]) == ref_state[2]


def _run_uf_ranks(arr, world, ll):
    from fof_b200 import _lib, distributed as D
    comms = D.ThreadComm.group(world)
    edges = D.slab_edges(arr[:, 2], world)
    out = [None] * world
    err = []

    def work(r):
        try:
            h = _lib.Handle(0)
            own, valid, local = D.plan_slab(edges, r, 2000.0)
            rows = np.nonzero((arr[:, 2] >= valid[0]) & (arr[:, 2] <= valid[1]))[0]
            out[r] = D.friends_of_friends_slabs(arr[rows], rows, own, comms[r], h, linking_length=ll)
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


@pytest.mark.parametrize("n,seed,kind,ll", [(60000, 5, "cosmos", 0.4), (40000, 6, "dense", 0.8)])
def test_union_find_forests_merge_across_slabs(lib, n, seed, kind, ll):
    """Transitive mode over slabs: per-rank union-find forests, boundary (row, label) pairs exchanged, replicated merge.
    The labels of the owned rows of 2, 3 and 4 ranks together must equal the single-GPU labels (min row of the component)."""
    from fof_b200 import transitive
    arr = make_catalog(n, seed=seed, kind=kind, as_frame=False)
    want, n_groups, _ = transitive.friends_of_friends(arr, linking_length=ll, count_links=False)
    assert np.unique(want, return_counts=True)[1].max() > 20
    for world in (2, 3, 4):
        parts = _run_uf_ranks(arr, world, ll)
        got = np.full(n, -1, dtype=np.int64)
        for lab, rows in parts:
            assert (got[rows] == -1).all()          # every row is owned by exactly one rank
            got[rows] = lab
        assert np.array_equal(got, want.astype(np.int64))
