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
