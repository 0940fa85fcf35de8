"""Virial mass estimator with the reference's names (/root/reference/FoF/analysis/mass_estimator.py)."""
import numpy as np

from .. import _lib
from ..params import cosmo
from ..units import Q


def _lib_params(**kw):
    return _lib.default_params(H0=cosmo.H0, Om0=cosmo.Om0, Ode0=cosmo.Ode0, **kw)


def velocity_dispersion(v_arr, group_size):
    """mass_estimator.py:21-23."""
    return np.sum(np.asarray(v_arr) ** 2) / (group_size - 1)


def projected_radius(group_size, separation):
    """mass_estimator.py:26-28."""
    return group_size * (group_size - 1) / separation


def virial_mass_estimator(cluster, device=0):
    """mass_estimator.py:36-112: (mass [Msun/h], mass_err, vel_disp [km^2/s^2], vel_disp_err, radius [Mpc/h])."""
    g = np.ascontiguousarray(np.asarray(cluster.galaxies, dtype=np.float64))
    h = _lib.get_handle(device)
    out = h.virial(np.array([0, len(g)], dtype=np.int64), g, _lib_params())[0]
    return (Q(out[0], "Msun/littleh"), Q(out[1], "Msun/littleh"), Q(out[2], "km2/s2"), Q(out[3], "km2/s2"),
            Q(out[4], "Mpc/littleh"))


def projected_mass_estimator(cluster, device=0):
    """mass_estimator.py:115-145: (projected mass [Msun/h], mass_err)."""
    g = np.ascontiguousarray(np.asarray(cluster.galaxies, dtype=np.float64))
    h = _lib.get_handle(device)
    out = h.projected(np.array([0, len(g)], dtype=np.int64), g, _lib_params())[0]
    return Q(out[0], "Msun/littleh"), Q(out[1], "km2/s2")


def mass_error(mass, vel_disp, vel_disp_err):
    """mass_estimator.py:170-171."""
    return mass * (vel_disp_err / vel_disp)


def log_error(error, quantity):
    """mass_estimator.py:174-175."""
    return 0.434 * (error / quantity)
