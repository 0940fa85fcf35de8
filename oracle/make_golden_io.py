"""Generate tests/golden/io/: a cluster list pickled by the UNMODIFIED reference classes (cluster.py, module name
``cluster``) the way pipeline.py:103-120 does, plus the attribute values a loader must recover.  Build-container only.

    python -m oracle.make_golden_io
"""
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from fof_b200.mock import make_catalog  # noqa: E402
from oracle.run_reference import load_reference, run_pipeline  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "io")


def main():
    import logging
    logging.disable(logging.CRITICAL)
    ref = load_reference()
    df = make_catalog(1500, seed=21)
    res = run_pipeline(ref, df, richness=8, overdensity=0.5, linking_length_factor=0.2, seed=99)
    cleaned = res["cleaned_objects"]
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "ref_cleaned_candidates.dat"), "wb") as f:
        pickle.dump(cleaned, f)
    off = np.zeros(len(cleaned) + 1, dtype=np.int64)
    np.cumsum([len(c.galaxies) for c in cleaned], out=off[1:])
    np.savez_compressed(os.path.join(OUT, "ref_cleaned_candidates_expected.npz"),
                        centre=np.array([c.center_attributes for c in cleaned]), D=np.array([c.D for c in cleaned]),
                        flag_poor=np.array([c.flag_poor for c in cleaned]), offsets=off,
                        members=np.concatenate([c.galaxies for c in cleaned]))
    print(len(cleaned), "clusters pickled by", type(cleaned[0]).__module__ + "." + type(cleaned[0]).__name__)


if __name__ == "__main__":
    main()
