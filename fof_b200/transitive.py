"""Transitive friends-of-friends (connected components with a lock-free GPU union-find).

This mode is NOT part of the reference: kencx/FoF groups galaxies around candidate centres (two hops) and never
computes connected components.  It exists because the project brief names a union-find linker; its results must
not be compared with the reference oracle.  See include/fofb.h (fofb_transitive_labels) for the link predicate.
"""
import numpy as np

from . import _lib
from . import params as P
from .fof import _as_matrix
from .units import as_float


def friends_of_friends(galaxy_data, linking_length=0.5, max_velocity=None, device=0, count_links=True):
    """Labels per row: the smallest row index of the row's connected component.

    linking_length: transverse linking length in h^-1 Mpc (proper); max_velocity in km/s (default params.py).
    count_links=False skips pairs that are already connected (the link count is then -1): far faster on percolating
    blobs.  Returns (labels int32[n], number_of_groups, number_of_links).
    """
    g = _as_matrix(galaxy_data)
    h = _lib.get_handle(device)
    p = _lib.default_params(H0=P.cosmo.H0, Om0=P.cosmo.Om0, Ode0=P.cosmo.Ode0,
                            max_velocity=as_float(P.max_velocity if max_velocity is None else max_velocity, "km/s"))
    return h.transitive_labels(g, p, as_float(linking_length), count_links=count_links)


def group_sizes(labels):
    """(unique labels, member counts)."""
    return np.unique(labels, return_counts=True)
