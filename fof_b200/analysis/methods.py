"""Scalar helpers with the reference's names (/root/reference/FoF/analysis/methods.py:12-87), evaluated
by libfofb200's host code (the same fp64 routines the kernels use)."""
import ctypes as C

import numpy as np

from .. import _lib
from ..params import cosmo  # noqa: F401  (analysis/methods.py:9 defines its own identical cosmo)
from ..units import Q, as_float


def _params(**kw):
    return _lib.default_params(H0=cosmo.H0, Om0=cosmo.Om0, Ode0=cosmo.Ode0, **kw)


def linear_to_angular_dist(distance, photo_z):
    """analysis/methods.py:12-32: proper distance [h^-1 Mpc] at redshift photo_z -> angle [deg]."""
    d = as_float(distance)
    z = np.ascontiguousarray(np.atleast_1d(photo_z), dtype=np.float64)
    out = np.empty_like(z)
    p = _params()
    unit = getattr(distance, "unit", "Mpc/littleh")
    plain_mpc = str(unit) in ("Mpc",)
    if plain_mpc:  # a length already in Mpc: the with_H0 step is a no-op (fof.py:173-175,198-200)
        d = d * (cosmo.H0 / 100.0)
    rc = _lib.load().fofb_linear_to_angular_dist(C.byref(p), d, z.size, z.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise _lib.FofbError(rc, "fofb_linear_to_angular_dist failed")
    return Q(out[0], "deg") if np.ndim(photo_z) == 0 else out


def mean_separation(n, z, max_dist, max_velocity, survey_area):
    """analysis/methods.py:35-65 (``max_dist`` is unused there too). Returns Mpc."""
    p = _params(max_velocity=as_float(max_velocity, "km/s"), survey_area=float(survey_area))
    out = C.c_double()
    rc = _lib.load().fofb_mean_separation(C.byref(p), float(n), float(z), C.byref(out))
    if rc != 0:
        raise _lib.FofbError(rc, "fofb_mean_separation failed")
    return Q(out.value, "Mpc")


def redshift_to_velocity(z, center_z):
    """analysis/methods.py:68-87: v = (c z - c z_c) / (1 + z_c) [km/s]."""
    zz = np.ascontiguousarray(np.atleast_1d(z), dtype=np.float64)
    out = np.empty_like(zz)
    _lib.load().fofb_redshift_to_velocity(zz.size, zz.ctypes.data, float(center_z), out.ctypes.data)
    return Q(out[0], "km/s") if np.ndim(z) == 0 else out
