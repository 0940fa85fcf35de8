"""Helpers shared by the golden-fixture tests (fixtures are written by oracle/make_golden.py)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fixtures():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(path):
    z = np.load(path)
    g = {k: z[k] for k in z.files}
    b, rich, D, seed = g["params"]
    g["b"], g["richness"], g["overdensity"], g["rng_seed"] = float(b), int(rich), float(D), int(seed)
    return g


def fof_inputs(g, zcut=(0.5, 2.52, 2.5)):
    """main_arr / candidate_centers exactly as the reference's pipeline hands them to FoF(): the golden N column
    and the golden candidate ORDER (the reference's default argsort leaves equal-N ties unspecified)."""
    main_arr = np.column_stack([g["table"], g["N"]])
    row_of_id = {v: i for i, v in enumerate(main_arr[:, 6].tolist())}
    cand = main_arr[[row_of_id[v] for v in g["candidate_ids"].tolist()]]
    gal = main_arr[(main_arr[:, 2] >= zcut[0]) & (main_arr[:, 2] <= zcut[1])]
    cen = cand[(cand[:, 2] >= zcut[0]) & (cand[:, 2] <= zcut[2])]
    return main_arr, np.ascontiguousarray(gal), np.ascontiguousarray(cen)


def split(off, ids):
    return [ids[off[k]:off[k + 1]] for k in range(len(off) - 1)]
