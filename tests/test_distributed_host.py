"""Host-side logic of the slab decomposition on CPU: slab planning invariants, and the collectives of
distributed.TorchComm under a real 2-process gloo group."""
import os
import subprocess
import sys

import numpy as np

from fof_b200 import distributed as D
from fof_b200.mock import make_catalog
from oracle import fof_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_plan_invariants():
    arr = make_catalog(30000, seed=5, as_frame=False)
    z = arr[:, 2]
    vmax = 2000.0
    for world in (1, 2, 3, 8):
        edges = D.slab_edges(z, world)
        owned = np.zeros(len(z), dtype=int)
        for r in range(world):
            rows, own, valid = D.select_local_rows(z, edges, r, vmax)
            own_mask = (z >= own[0]) & (z < own[1])
            owned += own_mask
            zl = z[rows]
            # every owned centre's velocity window lies inside the valid interval ...
            for zc in z[own_mask][:: max(1, own_mask.sum() // 50)]:
                win = z[np.abs(O.redshift_to_velocity(z, zc)) <= vmax]
                assert win.min() >= valid[0] and win.max() <= valid[1]
            # ... and every valid row's window is held locally
            vmask = (z >= valid[0]) & (z <= valid[1])
            for zc in z[vmask][:: max(1, vmask.sum() // 50)]:
                win = z[np.abs(O.redshift_to_velocity(z, zc)) <= vmax]
                assert win.min() >= zl.min() and win.max() <= zl.max()
            assert np.all(np.diff(rows) > 0)
        assert np.all(owned == 1)  # the owned intervals partition the table


def test_global_priority_is_the_single_table_order():
    rng = np.random.default_rng(1)
    N = rng.integers(8, 30, 500).astype(float)
    rows = rng.permutation(100000)[:500]
    order = D.global_priority(N, rows)
    table = np.zeros(100000)
    table[rows] = N
    want = np.nonzero(table >= 8)[0]
    want = want[table[want].argsort(kind="stable")[::-1]]
    assert np.array_equal(rows[order], want)


WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch.distributed as dist
from fof_b200 import distributed as D
dist.init_process_group("gloo")
c = D.TorchComm()
r = c.rank
x = np.arange((r + 1) * 6, dtype=np.float64).reshape(-1, 3) + 100 * r
parts = c.allgatherv(x)
assert len(parts) == 2 and parts[0].shape == (2, 3) and parts[1].shape == (4, 3)
assert np.array_equal(parts[r], x) and parts[1 - r][0, 0] == 100 * (1 - r)
e = c.allgatherv(np.zeros((0, 3)) if r == 0 else np.ones((2, 3)))
assert e[0].shape == (0, 3) and e[1].shape == (2, 3)
i32 = c.allgatherv(np.array([r, 7], dtype=np.int32))
assert i32[0].dtype == np.int32 and i32[1][0] == 1
assert c.allreduce(np.array([float(r + 1)]), "sum")[0] == 3.0
assert c.allreduce(np.array([float(r + 1)]), "min")[0] == 1.0
import torch
a = torch.tensor([1, r, 1], dtype=torch.uint8); d = torch.tensor([0, r, 1 - r], dtype=torch.uint8)
c.allreduce_bytes(a, d)
assert a.tolist() == [1, 0, 1] and d.tolist() == [0, 1, 1]
# one collective per round: states 0 alive < 1 switched off < 2 evaluated, merged by MAX; last byte = anybody active
buf = torch.tensor([0, 2 * r, 1 - r, r], dtype=torch.uint8)
c.allreduce_max_bytes(buf)
assert buf.tolist() == [0, 2, 1, 1]
# ragged tensors (exact sizes, no padding in the result) and the equal shares of the MT19937 stream
t = c.allgatherv_t(torch.arange(3 * r + 1, dtype=torch.int64) + 100 * r)
assert [x.tolist() for x in t] == [[0], [100, 101, 102, 103]]
full = torch.zeros(8, dtype=torch.int32)
work = c.allgather_shares(full, torch.arange(4, dtype=torch.int32) + 10 * r)
if work is not None:
    work.wait()
assert full.tolist() == [0, 1, 2, 3, 10, 11, 12, 13]
c.barrier()
dist.destroy_process_group()
print("ok", r)
"""


def test_torch_comm_under_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_thread_comm_collectives_world3():
    """ThreadComm (the ranks-as-threads communicator the slab parity tests use) implements the same collectives as the
    torch one: ragged all-gathers, sum/min/max all-reduce, and the MIN / MAX merge of the tracker state bytes."""
    import threading
    import torch
    from fof_b200.distributed import ThreadComm
    world = 3
    comms = ThreadComm.group(world)
    results = [None] * world
    errors = []

    def worker(r):
        try:
            c = comms[r]
            rows = np.arange(r * 2 + 1, dtype=np.float64).reshape(-1, 1) + 10 * r       # ragged: 1, 3, 5 rows
            g = c.allgatherv(rows)
            t = c.allgatherv_t(torch.arange(r + 1, dtype=torch.int64) + 100 * r)          # ragged tensors
            s = c.allreduce(np.array([r + 1.0, -r]), "sum")
            mn = c.allreduce(np.array([r + 1.0, -r]), "min")
            mx = c.allreduce(np.array([r + 1.0, -r]), "max")
            alive = torch.tensor([1, 1, 1, 1], dtype=torch.uint8)
            decided = torch.tensor([0, 0, 0, 0], dtype=torch.uint8)
            alive[r] = 0                                                                   # rank r kills centre r
            decided[3 - r] = 1                                                             # and decides centre 3 - r
            c.allreduce_bytes(alive, decided)
            state = torch.zeros(5, dtype=torch.uint8)
            state[r] = 1                                                                   # rank r switches centre r off
            state[3 - r] = 2                                                               # and evaluates centre 3 - r
            state[4] = 1 if r == 1 else 0                                                  # only rank 1 still has work
            c.allreduce_max_bytes(state)
            assert state.tolist() == [1, 2, 2, 2, 1], state.tolist()
            full = torch.zeros(6, dtype=torch.int32)                                       # MT19937 stream in equal shares
            assert c.allgather_shares(full, torch.tensor([10 * r, 10 * r + 1], dtype=torch.int32)) is None
            assert full.tolist() == [0, 1, 10, 11, 20, 21], full.tolist()
            c.barrier()
            results[r] = (g, t, s, mn, mx, alive.clone(), decided.clone())
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=60)
    assert not errors, errors
    for r in range(world):
        g, t, s, mn, mx, alive, decided = results[r]
        assert [len(x) for x in g] == [1, 3, 5] and g[2][0, 0] == 20.0
        assert [x.tolist() for x in t] == [[0], [100, 101], [200, 201, 202]]
        assert s.tolist() == [6.0, -3.0] and mn.tolist() == [1.0, -2.0] and mx.tolist() == [3.0, 0.0]
        assert alive.tolist() == [0, 0, 0, 1] and decided.tolist() == [0, 1, 1, 1]


def test_balanced_slab_edges_cover_the_table_and_even_out_the_held_rows():
    """slab_edges_balanced: increasing edges, every row owned by exactly one rank, and the heaviest rank holds clearly
    fewer (weighted) rows than under equal-count slabs."""
    from fof_b200 import distributed as D
    from fof_b200.mock import make_catalog
    z = make_catalog(200000, seed=3, as_frame=False)[:, 2]
    for world in (2, 4, 8):
        e = D.slab_edges_balanced(z[::4], world, 2000.0)
        assert len(e) == world + 1 and np.all(np.diff(e[1:-1]) > 0) and e[0] == -np.inf and e[-1] == np.inf
        owned = sum(int(((z >= e[r]) & (z < e[r + 1])).sum()) for r in range(world))
        assert owned == len(z)
        held = lambda edges: [int(((z >= D.plan_slab(edges, r, 2000.0)[2][0]) & (z <= D.plan_slab(edges, r, 2000.0)[2][1])).sum())
                              for r in range(world)]
        interior_equal = held(D.slab_edges(z, world))
        assert max(held(e)[1:-1] or [0]) <= max(interior_equal[1:-1] or [1 << 60])
