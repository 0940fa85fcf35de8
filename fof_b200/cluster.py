"""Cluster record types exposing the reference's attribute names (/root/reference/FoF/cluster.py:7-136), so that
user code written against the reference (``c.ra``, ``c.N``, ``c.galaxies``, ``c.richness``, ``c.center_attributes``,
``c.properties`` ...) keeps working.  The numerical work they used to carry lives in the CUDA library."""
import numpy as np

from .units import Q

# centre row layout handed to FoF(): ra, dec, z, z_lower, z_upper, RMag, ID[, extra columns ...], N(0.5)
_RA, _DEC, _Z, _ZLO, _ZHI, _MAG, _ID = range(7)
_MASS_FIELDS = ("cluster_mass", "mass_err", "vel_disp", "vel_disp_err", "projected_radius", "total_absmag")


class Cluster:
    """One candidate cluster: its centre galaxy's row and the member rows (cluster.py:7-85)."""

    __slots__ = ("ra", "dec", "z", "z_unc", "bcg_absMag", "gal_id", "N", "galaxies", "D", "flag_poor", "properties",
                 "_richness", "_center_attributes")

    def __init__(self, center, galaxies):
        row = center
        self.ra, self.dec, self.z = row[_RA], row[_DEC], row[_Z]
        self.z_unc = [row[_ZLO], row[_ZHI]]
        self.bcg_absMag, self.gal_id = row[_MAG], row[_ID]
        self.N = row[-1]                 # N(0.5) is always the last column
        self.galaxies = galaxies
        self.D = 0                       # overdensity, filled in by FoF() stage 4
        self.flag_poor = False           # set by hand in the reference's workflow
        if len(row) > 8:
            self.properties = row[7:-1]  # optional extra columns between ID and N

    # -- derived quantities ------------------------------------------------------------------------------------
    @property
    def richness(self):
        return len(self.galaxies)

    @property
    def center_attributes(self):
        """The 8-column centre row again (ra, dec, z, z_lower, z_upper, RMag, ID, N)."""
        return np.array([self.ra, self.dec, self.z, self.z_unc[0], self.z_unc[1], self.bcg_absMag, self.gal_id, self.N])

    def remove_overlap(self, other):
        """Winner-takes-all contest of cluster.py:65-82: returns (own rows, other's rows) with the loser's side
        emptied.  Larger N(0.5) wins; on a tie the brighter centre (larger |RMag|) does, and `other` wins a full tie."""
        if not isinstance(other, type(self)):
            return NotImplemented
        nothing = np.array([])
        self_wins = self.N > other.N or (self.N == other.N and abs(self.bcg_absMag) > abs(other.bcg_absMag))
        return (self.galaxies, nothing) if self_wins else (nothing, other.galaxies)

    def __str__(self):
        return "Cluster at ({:.2f},{:.2f},{:.2f}) with {} galaxies".format(self.ra, self.dec, self.z, self.richness)

    # -- pickling ----------------------------------------------------------------------------------------------
    # pipeline.py pickles the cluster lists between stages.  The slot named ``properties`` is shadowed by a read-only
    # property in CleanedCluster, which the default slot pickling cannot restore, hence the explicit state dict.
    def __getstate__(self):
        state = {}
        for klass in type(self).__mro__:
            for name in getattr(klass, "__slots__", ()):
                if name in state or isinstance(getattr(type(self), name, None), property):
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
    """A cluster after interloper removal, with its mass estimate (cluster.py:88-136).  The Quantity-valued fields of
    the reference are ``units.Q`` floats here (``.value`` works): cluster_mass [h^-1 Msun], vel_disp [(km/s)^2],
    projected_radius [h^-1 Mpc]."""

    __slots__ = _MASS_FIELDS

    def __init__(self, center, galaxies, cluster_mass, mass_err, vel_disp, vel_disp_err, projected_radius, absmag):
        super().__init__(center, galaxies)
        for name, value in zip(_MASS_FIELDS, (cluster_mass, mass_err, vel_disp, vel_disp_err, projected_radius, absmag)):
            setattr(self, name, value)

    def r200(self, cosmo, delta):
        """Radius [h^-1 Mpc] of the sphere whose mean density is ``delta`` times critical at the cluster's redshift
        (cluster.py:111-118), in SI with CODATA-2018 / IAU-2015 constants."""
        G = 6.6743e-11
        mpc_m = 1.0e6 * 149597870700.0 * 648000.0 / np.pi
        msun_kg = 1.3271244e20 / G
        h = cosmo.H0 / 100.0
        H0_si = cosmo.H0 * 1.0e3 / mpc_m
        zp1 = 1.0 + self.z
        e2 = cosmo.Om0 * zp1 ** 3 + (1 - cosmo.Om0 - cosmo.Ode0) * zp1 ** 2 + cosmo.Ode0
        rho = 3.0 * H0_si ** 2 / (8.0 * np.pi * G) * e2          # critical density, kg / m^3
        mass_kg = float(self.cluster_mass) / h * msun_kg           # Msun/h -> kg
        radius_m = ((3 / 4) * mass_kg / (delta * np.pi * rho)) ** (1 / 3)
        return Q(radius_m / mpc_m * h, "Mpc/littleh")

    @property
    def properties(self):
        """The 11-value export vector of cluster.py:120-136 (``io.PROPERTY_COLUMNS`` names the entries)."""
        return np.array([self.ra, self.dec, self.z, self.cluster_mass.value, self.vel_disp.value,
                         self.projected_radius.value, self.total_absmag, self.richness, self.D, self.flag_poor,
                         self.gal_id])
