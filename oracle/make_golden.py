"""Generate tests/golden/*.npz by running the UNMODIFIED reference files (/root/reference/FoF) against the
third-party shims on small seeded mocks.  Build-container only (the reference tree is not on the GPU box).

    python -m oracle.make_golden

Each fixture stores the input table and every stage artefact the parity tests compare: the N(0.5) column, the
candidate-centre order the reference produced on this machine (its default argsort leaves ties unspecified), the
FoF() cluster lists (member rows in order, N, D), the cleaned lists and the virial properties.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from fof_b200.mock import make_catalog  # noqa: E402
from oracle.run_reference import load_reference, run_pipeline  # noqa: E402

CASES = {
    # name: (n, mock seed, mock kwargs, b, richness, overdensity, rng seed)
    "cosmos_3000_default": (3000, 2, {}, 0.2, 12, 2.0, 12345),
    "cosmos_4000_b005": (4000, 4, {}, 0.05, 8, 0.5, 2468),
    "dense_2500_b002": (2500, 6, {"density": 200000.0}, 0.02, 5, -5.0, 1357),
    "wide_3000_default": (3000, 7, {"kind": "wide"}, 0.2, 12, 2.0, 97531),
    # z_lower / z_upper quoted to 3 (2) decimals: dozens (hundreds) of rows share each column-4 value, the key of the
    # reference's tracker (fof.py:238-246: np.intersect1d switches off the FIRST candidate row carrying the value)
    "photoz3_4000_default": (4000, 12, {"photoz_decimals": 3}, 0.2, 12, 2.0, 4321),
    "photoz2_4000_b01": (4000, 13, {"photoz_decimals": 2, "density": 100000.0}, 0.1, 10, 1.0, 8642),
}


def pack(clusters_members):
    off = np.zeros(len(clusters_members) + 1, dtype=np.int64)
    if clusters_members:
        np.cumsum([len(m) for m in clusters_members], out=off[1:])
        ids = np.concatenate([m[:, 6] for m in clusters_members]) if off[-1] else np.zeros(0)
    else:
        ids = np.zeros(0)
    return off, ids


def main():
    import logging
    logging.disable(logging.CRITICAL)
    ref = load_reference()
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    for name, (n, seed, kw, b, rich, D, rng_seed) in CASES.items():
        df = make_catalog(n, seed=seed, **kw)
        R = run_pipeline(ref, df, richness=rich, overdensity=D, linking_length_factor=b, seed=rng_seed)
        c_off, c_ids = pack([c[1] for c in R["clusters"]])
        k_off, k_ids = pack([c[1] for c in R["cleaned"]])
        keys = ("cluster_mass", "mass_err", "vel_disp", "vel_disp_err", "projected_radius", "total_absmag")
        np.savez_compressed(
            os.path.join(outdir, name + ".npz"),
            table=df.values, params=np.array([b, rich, D, rng_seed], dtype=np.float64),
            N=R["main_arr"][:, 7], candidate_ids=R["candidate_centers"][:, 6],
            cluster_off=c_off, cluster_member_ids=c_ids,
            cluster_centre_ids=np.array([c[0][6] for c in R["clusters"]]),
            cluster_N=np.array([c[2] for c in R["clusters"]]), cluster_D=np.array([c[3] for c in R["clusters"]]),
            cleaned_off=k_off, cleaned_member_ids=k_ids,
            cleaned_centre_ids=np.array([c[0][6] for c in R["cleaned"]]),
            number_removed=np.array(R["number_removed"], dtype=np.int64),
            virial=np.array([[v[k] for k in keys] for v in R["virial"]]).reshape(-1, len(keys)),
            projected=np.array(R["projected"]).reshape(-1, 2),
            r200=np.array([[v["r200"], v["r500"]] for v in R["virial"]]).reshape(-1, 2))
        print(name, "clusters", len(R["clusters"]), "cleaned", len(R["cleaned"]), "candidates", len(R["candidate_centers"]))


if __name__ == "__main__":
    main()
