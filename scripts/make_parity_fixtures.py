"""Compact whole-path fixtures at the BASELINE.json sizes, produced by the CPU oracle (fast mode) in the build container:

    python scripts/make_parity_fixtures.py            # configs[0] 100k and configs[1] 1M (about 10 minutes on one core)
    python scripts/make_parity_fixtures.py 100000     # one size

Each fixture holds what tests/test_parity_sizes_gpu.py compares with the GPU path: the mock's seed and size (the table
is regenerated from them, it is not stored), a checksum of the N(0.5) column, the candidate count, and the final
catalogue in CSR form -- member ids in the reference's order, centre ids, N, D and the six virial properties.
The oracle itself is pinned to the unmodified reference by tests/golden/*.npz (oracle/make_golden.py).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from fof_b200.mock import make_catalog  # noqa: E402
from oracle import fof_oracle as O  # noqa: E402

KEYS = ("cluster_mass", "mass_err", "vel_disp", "vel_disp_err", "projected_radius", "total_absmag")
# name: (n, mock seed, mock kwargs, rng seed)
CASES = {
    "parity_100k": (100_000, 20260101, {}, 2025),
    "parity_1M": (1_000_000, 20260102, {}, 2026),
    "parity_100k_photoz": (100_000, 20260105, {"photoz_decimals": 3}, 2027),
}


def n_checksum(N):
    """Order-sensitive 64-bit checksum of the N(0.5) column (exact small integers)."""
    v = N.astype(np.uint64)
    w = (np.arange(len(v), dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) ^ np.uint64(0xD1B54A32D192ED03)
    return int(np.bitwise_xor.reduce((v + np.uint64(1)) * w))


def build(name):
    n, seed, kw, rng_seed = CASES[name]
    arr = make_catalog(n, seed=seed, as_frame=False, **kw)
    t0 = time.time()
    ref = O.run_pipeline(arr, O.Params(), mode="fast", seed=rng_seed)
    dt = time.time() - t0
    vir = ref["virial"]
    off = np.zeros(len(vir) + 1, dtype=np.int64)
    if vir:
        np.cumsum([len(v["galaxies"]) for v in vir], out=off[1:])
    ids = np.concatenate([v["galaxies"][:, 6] for v in vir]).astype(np.int32) if vir else np.zeros(0, np.int32)
    out = os.path.join(ROOT, "tests", "golden", "sizes", name + ".npz")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    np.savez_compressed(
        out, n=n, seed=seed, rng_seed=rng_seed, photoz_decimals=kw.get("photoz_decimals", -1),
        N_checksum=np.uint64(n_checksum(ref["main_arr"][:, 7])), N_sum=int(ref["main_arr"][:, 7].sum()),
        n_candidates=len(ref["candidate_centers"]), n_stage4=len(ref["clusters"]),
        off=off, member_ids=ids, centre_ids=np.array([v["center"][6] for v in vir], dtype=np.int64),
        N=np.array([v["N"] for v in vir]), D=np.array([v["D"] for v in vir]),
        props=np.array([[v[k] for k in KEYS] for v in vir]).reshape(-1, len(KEYS)), oracle_seconds=dt)
    print(f"{name}: n={n} candidates={len(ref['candidate_centers'])} stage4={len(ref['clusters'])} final={len(vir)} "
          f"members={off[-1]} oracle {dt:.1f} s -> {out} ({os.path.getsize(out) / 1e6:.2f} MB)", flush=True)


if __name__ == "__main__":
    want = [a for a in sys.argv[1:]]
    for name in CASES:
        if not want or name in want or str(CASES[name][0]) in want:
            build(name)
