"""astropy.cosmology.LambdaCDM restated for flat, radiation-free parameters.

Call sites: params.py:16, analysis/methods.py:9,31,59, analysis/mass_estimator.py:17,86,
fof.py:169,430, cluster.py:112.  With Tcmb0 = 0 and Ok0 = 0 astropy (>= 3.0) evaluates the
comoving distance in closed form through scipy.special.hyp2f1; that is what is restated here.
"""
import math
import numpy as np
from scipy.special import hyp2f1
from . import units as u
from . import constants as const


class LambdaCDM:
    def __init__(self, H0, Om0, Ode0, Tcmb0=0.0, Neff=3.04, m_nu=0.0, Ob0=None, name=None):
        self.H0 = H0 if isinstance(H0, u.Quantity) else u.Quantity(H0, u.km / u.s / u.Mpc)
        self._H0 = self.H0.to(u.km / u.s / u.Mpc).value
        self.Om0 = self._Om0 = float(Om0)
        self.Ode0 = self._Ode0 = float(Ode0)
        self.Ok0 = self._Ok0 = 1.0 - self._Om0 - self._Ode0
        if Tcmb0 != 0 or self._Ok0 != 0:
            raise NotImplementedError("shim restates only the flat, Tcmb0=0 branch used by params.py:16")
        self.h = self._H0 / 100.0
        self._hubble_distance = const.c.to(u.km / u.s).value / self._H0  # Mpc

    def efunc(self, z):
        zp1 = 1.0 + np.asarray(z, dtype=float)
        return np.sqrt(zp1 ** 2 * (self._Om0 * zp1 + self._Ok0) + self._Ode0)

    @staticmethod
    def _T_hypergeometric(x):
        return 2 * np.sqrt(x) * hyp2f1(1.0 / 6, 1.0 / 2, 7.0 / 6, -(x ** 3))

    def _comoving_distance_z1z2(self, z1, z2):
        s = ((1 - self._Om0) / self._Om0) ** (1.0 / 3)
        prefactor = self._hubble_distance / np.sqrt(s * self._Om0)
        z1 = np.asarray(z1, dtype=float)
        z2 = np.asarray(z2, dtype=float)
        return prefactor * (self._T_hypergeometric(s / (z1 + 1.0)) - self._T_hypergeometric(s / (z2 + 1.0)))

    def comoving_distance(self, z):
        return u.Quantity(self._comoving_distance_z1z2(0.0, z), u.Mpc)

    def comoving_transverse_distance(self, z):
        return self.comoving_distance(z)  # Ok0 == 0

    def angular_diameter_distance(self, z):
        z = np.asarray(z, dtype=float)
        return u.Quantity(self._comoving_distance_z1z2(0.0, z) / (1.0 + z), u.Mpc)

    def kpc_proper_per_arcmin(self, z):
        d = self.angular_diameter_distance(z).to(u.kpc)
        return u.Quantity(d.value * (math.pi / 10800.0), u.kpc / u.arcmin)

    def arcsec_per_kpc_proper(self, z):
        return (1.0 / self.kpc_proper_per_arcmin(z)).to(u.arcsec / u.kpc)

    def differential_comoving_volume(self, z):
        dm = self._comoving_distance_z1z2(0.0, z)
        return u.Quantity(self._hubble_distance * (dm ** 2.0) / self.efunc(z), u.Mpc ** 3 / u.sr)

    def critical_density(self, z):
        H0_si = self.H0.si.value
        rho0 = 3.0 * H0_si ** 2 / (8.0 * math.pi * const.G.value)  # kg / m^3
        return u.Quantity(rho0 * 1e-3 * self.efunc(z) ** 2, u.g / u.cm ** 3)


FlatLambdaCDM = None  # not used by the reference
