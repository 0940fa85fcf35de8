"""CPU-path scaling study: the oracle's literal mode (the reference's own O(n^2) control flow, one
core) and its grid-accelerated fast mode on COSMOS-like mocks of growing size, with a power-law fit and the labelled
extrapolation to the bench size.  Runs without a GPU:  python scripts/cpu_scaling.py > profiles/r01_cpu_scaling.json"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fof_b200.mock import make_catalog  # noqa: E402
from oracle import fof_oracle as O  # noqa: E402


def timed(n, mode, seed=777):
    df = make_catalog(n, seed=seed, as_frame=False)
    t0 = time.perf_counter()
    res = O.run_pipeline(df, O.Params(), mode=mode, seed=12345)
    dt = time.perf_counter() - t0
    return dt, len(res["cleaned"]) if isinstance(res, dict) and "cleaned" in res else None


def main():
    out = {"host": {"cores_used": 1, "nproc": os.cpu_count()}, "literal": [], "fast": []}
    for n in (5000, 10000, 20000, 40000):
        dt, k = timed(n, "literal")
        out["literal"].append({"n": n, "seconds": round(dt, 3), "galaxies_per_s": round(n / dt, 1), "clusters": k})
        print(out["literal"][-1], file=sys.stderr, flush=True)
    for n in (20000, 100000, 300000):
        dt, k = timed(n, "fast")
        out["fast"].append({"n": n, "seconds": round(dt, 3), "galaxies_per_s": round(n / dt, 1), "clusters": k})
        print(out["fast"][-1], file=sys.stderr, flush=True)
    for mode in ("literal", "fast"):
        ns = np.array([r["n"] for r in out[mode]], dtype=float)
        ts = np.array([r["seconds"] for r in out[mode]])
        slope, icpt = np.polyfit(np.log(ns), np.log(ts), 1)
        out[mode + "_fit"] = {"seconds = a * n^k": {"k": round(float(slope), 3), "a": float(np.exp(icpt))},
                              "EXTRAPOLATED_seconds_at_1e7": float(np.exp(icpt) * 1e7 ** slope)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
