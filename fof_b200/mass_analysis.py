"""Drop-in for /root/reference/FoF/mass_analysis.py:8-44."""
import numpy as np

from . import _lib
from .cluster import CleanedCluster
from .fof import _pack
from .params import cosmo
from .units import Q


def find_masses(data, estimator, device=0, keep_flags=True):
    """mass_analysis.py:8-44: wrap every cluster in a CleanedCluster carrying its virial mass, velocity
    dispersion (a variance, km^2/s^2), projected radius and total absolute magnitude.

    ``keep_flags``: the reference rebuilds each object through ``Cluster.__init__`` (cluster.py:40-41), so its
    CleanedClusters come back with ``D = 0`` and ``flag_poor = False`` whatever FoF() and the flagging step set
    (``properties[8:10]`` lose them).  True (default) carries the input's D / flag_poor over -- a documented
    deviation; False reproduces the reference's reset bit for bit."""
    if estimator not in ("virial", "projected"):
        raise ValueError("estimator must be 'virial' or 'projected'")
    if not len(data):
        return []
    h = _lib.get_handle(device)
    p = _lib.default_params(H0=cosmo.H0, Om0=cosmo.Om0, Ode0=cosmo.Ode0)
    off, members = _pack(data)
    if estimator == "projected":  # mass_analysis.py:36-40
        pm = h.projected(off, members, p)
        out = []
        for k, c in enumerate(data):
            new_c = CleanedCluster(c.center_attributes, c.galaxies, Q(pm[k, 0], "Msun/littleh"), Q(pm[k, 1], "km2/s2"),
                                   None, None, None, None)
            if keep_flags:
                new_c.D = c.D
                new_c.flag_poor = c.flag_poor
            out.append(new_c)
        return out
    props = h.virial(off, members, p)
    new_cs = []
    for k, c in enumerate(data):
        m, me, vd, vde, r, absmag = props[k, :6]
        new_c = CleanedCluster(c.center_attributes, c.galaxies, Q(m, "Msun/littleh"), Q(me, "Msun/littleh"),
                               Q(vd, "km2/s2"), Q(vde, "km2/s2"), Q(r, "Mpc/littleh"), float(absmag))
        if keep_flags:
            new_c.D = c.D
            new_c.flag_poor = c.flag_poor
        assert new_c.gal_id == c.gal_id, "Different cluster has been added"
        new_cs.append(new_c)
    return new_cs
