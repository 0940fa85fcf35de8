"""Cluster record types with the reference's attributes (/root/reference/FoF/cluster.py:7-136)."""
import numpy as np

from .units import Q


class Cluster:
    """cluster.py:7-85."""

    __slots__ = ["ra", "dec", "z", "z_unc", "bcg_absMag", "gal_id", "N", "galaxies", "D", "flag_poor",
                 "properties", "_richness", "_center_attributes"]

    def __init__(self, center, galaxies):
        self.ra = center[0]
        self.dec = center[1]
        self.z = center[2]
        self.z_unc = [center[3], center[4]]
        self.bcg_absMag = center[5]
        self.gal_id = center[6]
        self.N = center[-1]
        self.galaxies = galaxies
        self.D = 0
        self.flag_poor = False
        if len(center) > 8:
            self.properties = center[7:-1]

    @property
    def richness(self):
        return len(self.galaxies)

    @property
    def center_attributes(self):
        return np.array([self.ra, self.dec, self.z, self.z_unc[0], self.z_unc[1], self.bcg_absMag, self.gal_id,
                         self.N])

    def remove_overlap(self, other):
        """cluster.py:65-82: (self_galaxies, other_galaxies); the loser is emptied."""
        if not isinstance(other, type(self)):
            return NotImplemented
        if self.N > other.N:
            return self.galaxies, np.array([])
        if self.N < other.N:
            return np.array([]), other.galaxies
        if abs(self.bcg_absMag) > abs(other.bcg_absMag):
            return self.galaxies, np.array([])
        return np.array([]), other.galaxies

    def __str__(self):
        return f"Cluster at ({self.ra:.2f},{self.dec:.2f},{self.z:.2f}) with {self.richness} galaxies"

    # pipeline.py pickles the cluster lists between stages; the slot named ``properties`` is shadowed by a
    # read-only property in CleanedCluster, which the default slot pickling cannot restore
    def __getstate__(self):
        state = {}
        for cls in type(self).__mro__:
            for name in getattr(cls, "__slots__", ()):
                if isinstance(getattr(type(self), name, None), property) or name in state:
                    continue
                try:
                    state[name] = object.__getattribute__(self, name)
                except AttributeError:
                    pass
        return state

    def __setstate__(self, state):
        if isinstance(state, tuple):   # default slot pickling (None, {slot: value}), e.g. written by the reference
            merged = {}
            for part in state:
                merged.update(part or {})
            state = merged
        for name, value in state.items():
            setattr(self, name, value)


class CleanedCluster(Cluster):
    """cluster.py:88-136.  Quantity-valued fields are ``units.Q`` floats (``.value`` works)."""

    __slots__ = ["cluster_mass", "mass_err", "vel_disp", "vel_disp_err", "projected_radius", "total_absmag"]

    def __init__(self, center, galaxies, cluster_mass, mass_err, vel_disp, vel_disp_err, projected_radius, absmag):
        super().__init__(center, galaxies)
        self.cluster_mass = cluster_mass      # h^-1 M_sun
        self.mass_err = mass_err
        self.vel_disp = vel_disp              # (km/s)^2
        self.vel_disp_err = vel_disp_err
        self.projected_radius = projected_radius  # h^-1 Mpc
        self.total_absmag = absmag

    def r200(self, cosmo, delta):
        """cluster.py:111-118: radius [h^-1 Mpc] enclosing ``delta`` times the critical density at z."""
        G = 6.6743e-11
        mpc_m = 1.0e6 * 149597870700.0 * 648000.0 / np.pi
        msun_kg = 1.3271244e20 / G
        h = cosmo.H0 / 100.0
        H0_si = cosmo.H0 * 1.0e3 / mpc_m
        zp1 = 1.0 + self.z
        rho = 3.0 * H0_si ** 2 / (8.0 * np.pi * G) * (cosmo.Om0 * zp1 ** 3 + (1 - cosmo.Om0 - cosmo.Ode0) * zp1 ** 2
                                                      + cosmo.Ode0)  # kg / m^3
        mass_kg = float(self.cluster_mass) / h * msun_kg       # M_sun/h -> kg
        radius_m = ((3 / 4) * mass_kg / (delta * np.pi * rho)) ** (1 / 3)
        return Q(radius_m / mpc_m * h, "Mpc/littleh")

    @property
    def properties(self):
        return np.array([self.ra, self.dec, self.z, self.cluster_mass.value, self.vel_disp.value,
                         self.projected_radius.value, self.total_absmag, self.richness, self.D, self.flag_poor,
                         self.gal_id])
