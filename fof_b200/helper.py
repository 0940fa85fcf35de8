"""The hot-path helpers of /root/reference/FoF/helper.py with the same names and argument meaning:
``setdiff2d`` (helper.py:138-146), ``find_number_count`` (:149-188), ``center_overdensity`` (:191-221).
Counting is done on the GPU (fofb_bubble_count); ``export`` (helper.py:224-237) lives in fof_b200.io and is re-exported here;
the plotting / LOWESS helpers are out of scope."""
import numpy as np

from . import _lib
from . import params as P
from .analysis.methods import linear_to_angular_dist
from .cluster import Cluster
from .units import Q, as_float
from .io import export  # noqa: F401  (helper.py:224-237)


def setdiff2d(cluster1, cluster2):
    """Rows of cluster1 whose gal_id (column 6) does not occur in cluster2 (helper.py:138-146)."""
    return cluster1[np.isin(cluster1[:, 6], cluster2[:, 6], invert=True), :]


def _lib_params(max_velocity=None):
    return _lib.default_params(H0=P.cosmo.H0, Om0=P.cosmo.Om0, Ode0=P.cosmo.Ode0,
                               max_velocity=as_float(P.max_velocity if max_velocity is None else max_velocity, "km/s"))


def find_number_count(center, galaxies, distance=Q(0.5, "Mpc/littleh"), device=0):
    """N(d): number of rows of ``galaxies[:, :2]`` within the angle of ``distance`` [h^-1 Mpc] of the centre
    (helper.py:149-188; counts the centre itself when it is one of the rows)."""
    if isinstance(center, Cluster) or hasattr(center, "gal_id"):
        coords, z = [center.ra, center.dec], center.z
    else:
        coords, z = center[:2], center[2]
    theta = float(linear_to_angular_dist(distance, z))
    g = np.ascontiguousarray(np.asarray(galaxies, dtype=np.float64)[:, :2])
    return int(_lib.get_handle(device).bubble_count(g, np.array([coords], dtype=np.float64), theta, _lib_params())[0])


def center_overdensity(center, galaxy_data, max_velocity, device=0):
    """D = (N - mean) / rms over the N(0.5) counts at 300 uniform random points (helper.py:191-221); the points
    come from numpy's global generator exactly as in the reference."""
    n = 300
    galaxy_data = np.asarray(galaxy_data, dtype=np.float64)
    ra_random = np.random.uniform(low=min(galaxy_data[:, 0]), high=max(galaxy_data[:, 0]), size=n)
    dec_random = np.random.uniform(low=min(galaxy_data[:, 1]), high=max(galaxy_data[:, 1]), size=n)
    points = np.vstack((ra_random, dec_random)).T
    theta = float(linear_to_angular_dist(Q(0.5, "Mpc/littleh"), center.z))
    h = _lib.get_handle(device)
    N_arr = h.bubble_count(np.ascontiguousarray(galaxy_data[:, :3]), points, theta, _lib_params(max_velocity),
                           window_z=center.z).astype(np.int64)
    N_mean = np.mean(N_arr)
    N_rms = np.sqrt(np.mean(N_arr ** 2))
    return (center.N - N_mean) / N_rms
