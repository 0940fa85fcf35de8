"""Parameters of the cluster finder: same names and defaults as /root/reference/FoF/params.py:16-38,
without its import-time side effects (no directory creation, no matplotlib import)."""
from .units import Q


class Cosmology:
    """Stand-in for ``LambdaCDM(H0=70, Om0=0.3, Ode0=0.7)`` (params.py:16) backed by libfofb200's host
    helpers (fofb_comoving_distance etc.)."""

    def __init__(self, H0=70.0, Om0=0.3, Ode0=0.7):
        self.H0, self.Om0, self.Ode0 = float(H0), float(Om0), float(Ode0)

    def _params(self):
        from . import _lib
        return _lib.default_params(H0=self.H0, Om0=self.Om0, Ode0=self.Ode0)

    def _vec(self, name, z):
        import ctypes as C
        import numpy as np
        from . import _lib
        zz = np.ascontiguousarray(np.atleast_1d(z), dtype=np.float64)
        out = np.empty_like(zz)
        p = self._params()
        rc = getattr(_lib.load(), name)(C.byref(p), zz.size, zz.ctypes.data, out.ctypes.data)
        if rc != 0:
            raise _lib.FofbError(rc, "cosmology helper failed")
        return out if np.ndim(z) else float(out[0])

    def comoving_distance(self, z):
        return self._vec("fofb_comoving_distance", z)

    comoving_transverse_distance = comoving_distance  # flat

    def angular_diameter_distance(self, z):
        return self._vec("fofb_angular_diameter_distance", z)


# cosmology (params.py:16)
cosmo = Cosmology(H0=70.0, Om0=0.3, Ode0=0.7)

# redshift range (params.py:19-20)
min_redshift = 0.5
max_redshift = 2.5

# FoF key parameters (params.py:23-27)
max_velocity = Q(2000.0, "km/s")
linking_length_factor = 0.2  # b
max_radius = Q(1.5, "Mpc/littleh")
richness = 12
D = 2  # overdensity

runname = "COSMOS"
