"""CPU oracle of the transitive mode (TEST INFRASTRUCTURE ONLY): scipy connected components on the same link
predicate as fofb_transitive_labels.  This is not a restatement of reference code -- the reference has no
connected-components step; the helpers it uses are the oracle's restatements of
analysis/methods.py:12-32,68-87 and of the haversine metric."""
import math

import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components
from scipy.spatial import cKDTree

from . import fof_oracle as O


def labels(galaxy_arr, linking_length_mpc_h, max_velocity=2000.0):
    arr = np.asarray(galaxy_arr, dtype=float)
    n = len(arr)
    ra, dec, z = arr[:, 0], arr[:, 1], arr[:, 2]
    theta = O.linear_to_angular_dist(linking_length_mpc_h, z)
    cosmin = np.cos(np.deg2rad(min(89.0, np.abs(dec).max() + 1.0)))
    reach = float(theta.max()) / cosmin * 1.001 + 1e-9
    pairs = cKDTree(arr[:, :2]).query_pairs(reach * math.sqrt(2.0), p=2.0, output_type="ndarray")
    i, j = pairs[:, 0], pairs[:, 1]
    ok = (np.abs(O.redshift_to_velocity(z[j], z[i])) <= max_velocity) & (np.abs(O.redshift_to_velocity(z[i], z[j])) <= max_velocity)
    i, j = i[ok], j[ok]
    d = O.haversine_deg(ra[i], dec[i], ra[j], dec[j])
    ok = (d <= theta[i]) & (d <= theta[j])
    i, j = i[ok], j[ok]
    graph = coo_matrix((np.ones(len(i), dtype=np.int8), (i, j)), shape=(n, n))
    _, comp = connected_components(graph, directed=False)
    # canonical label: smallest member index
    first = np.full(comp.max() + 1, n, dtype=np.int64)
    np.minimum.at(first, comp, np.arange(n))
    return first[comp].astype(np.int32), len(i)
