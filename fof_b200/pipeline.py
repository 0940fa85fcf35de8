"""Whole-job entry point: the call sequence of /root/reference/FoF/pipeline.py:62-64,80-99,113,173
(candidate_center_search -> z cuts -> FoF -> interloper_removal -> find_masses("virial")) as ONE call
that keeps every intermediate on the GPU.  Results come back columnar; ``Catalogue.clusters()``
materialises ``CleanedCluster`` objects on demand.
"""
import numpy as np

from . import _lib
from . import params as P
from .cluster import CleanedCluster
from .fof import _as_matrix
from .units import Q, as_float


class Catalogue:
    """Final cluster catalogue in CSR form, indexing rows of the input table."""

    def __init__(self, galaxy_arr, offsets, member_rows, centre_row, N, D, props, N_row=None):
        self.galaxy_arr = galaxy_arr
        self.offsets, self.member_rows, self.centre_row = offsets, member_rows, centre_row
        self.N, self.D, self.props = N, D, props
        self.N_row = N_row   # N(0.5) per input row (column 7 of main_arr, fof.py:61-74); None: not fetched

    def __len__(self):
        return len(self.centre_row)

    def members(self, k):
        return self.member_rows[self.offsets[k]:self.offsets[k + 1]]

    def labels(self):
        """Canonical group label per input row: the smallest member row of the (first) final cluster
        containing it, -1 for field galaxies (labels derive from the final lists)."""
        lab = np.full(len(self.galaxy_arr), -1, dtype=np.int64)
        for k in range(len(self) - 1, -1, -1):
            m = self.members(k)
            if len(m):
                cand = np.full(len(m), m.min(), dtype=np.int64)
                cur = lab[m]
                lab[m] = np.where(cur < 0, cand, np.minimum(cur, cand))
        return lab

    def cluster(self, k):
        g = self.galaxy_arr
        center = np.append(g[self.centre_row[k]], self.N[k])
        if self.N_row is None:
            raise RuntimeError("this Catalogue was built without the per-row N(0.5) column (find_clusters(..., "
                               "with_counts=False)); Cluster.galaxies rows carry it in the reference")
        rows = np.column_stack([g[self.members(k)], self.N_row[self.members(k)].astype(np.float64)])
        m, me, vd, vde, r, absmag = self.props[k, :6]
        c = CleanedCluster(center, rows, Q(m, "Msun/littleh"), Q(me, "Msun/littleh"), Q(vd, "km2/s2"),
                           Q(vde, "km2/s2"), Q(r, "Mpc/littleh"), float(absmag))
        c.D = self.D[k]
        return c

    def clusters(self):
        return [self.cluster(k) for k in range(len(self))]


def find_clusters(galaxy_data, max_velocity=None, linking_length_factor=None, virial_radius=None, richness=None,
                  overdensity=None, min_redshift=None, max_redshift=None, device=0, uniforms=None, grid_cells=None,
                  with_counts=True):
    """Run the whole cluster-finding path on ``galaxy_data`` (DataFrame or (n, 7) array in the column order
    of pipeline.py:51-53).  Defaults come from ``fof_b200.params``.  The overdensity sample points consume
    ``np.random`` (legacy global generator) exactly as the reference does.  ``grid_cells``: GriSPy's N_cells (64, the
    library default the reference runs with).  ``with_counts``: also read back N(0.5) of every row (4 bytes per row), so that
    ``Catalogue.cluster(k).galaxies`` has the reference's 8 columns."""
    g = _as_matrix(galaxy_data)
    h = _lib.get_handle(device)
    p = _lib.default_params(
        H0=P.cosmo.H0, Om0=P.cosmo.Om0, Ode0=P.cosmo.Ode0,
        max_velocity=as_float(P.max_velocity if max_velocity is None else max_velocity, "km/s"),
        linking_length_factor=float(P.linking_length_factor if linking_length_factor is None else linking_length_factor),
        virial_radius=as_float(P.max_radius if virial_radius is None else virial_radius),
        richness=int(P.richness if richness is None else richness),
        overdensity=float(P.D if overdensity is None else overdensity))
    if grid_cells is not None:
        p.grid_cells = int(grid_cells)
    zmin = P.min_redshift if min_redshift is None else min_redshift
    zmax = P.max_redshift if max_redshift is None else max_redshift
    if uniforms is None:
        Kf, total = h.pipeline_run_np_random(g, p, zmin, zmax + 0.02, zmax)
    else:
        h.pipeline_begin(g, p, zmin, zmax + 0.02, zmax)
        Kf, total = h.pipeline_finish(uniforms)
    return Catalogue(g, *h.pipeline_fetch(Kf, total), N_row=h.pipeline_fetch_counts(len(g)) if with_counts else None)


def sweep(galaxy_data, linking_length_factors, richness=None, overdensity=None, max_velocity=None, virial_radius=None,
          min_redshift=None, max_redshift=None, device=0):
    """The thesis' tuning loop (README.md:21; BASELINE.json configs[2]): one catalogue, several linking-length
    factors.  Stage 0, the z-sorted catalogue, the cell lists and the per-centre scalars are built once; every further
    factor reruns only the priority rounds and what follows.  Returns {b: Catalogue}."""
    g = _as_matrix(galaxy_data)
    h = _lib.get_handle(device)
    rich = int(P.richness if richness is None else richness)
    od = float(P.D if overdensity is None else overdensity)
    p = _lib.default_params(
        H0=P.cosmo.H0, Om0=P.cosmo.Om0, Ode0=P.cosmo.Ode0,
        max_velocity=as_float(P.max_velocity if max_velocity is None else max_velocity, "km/s"),
        virial_radius=as_float(P.max_radius if virial_radius is None else virial_radius), richness=rich, overdensity=od,
        linking_length_factor=float(linking_length_factors[0]))
    zmin = P.min_redshift if min_redshift is None else min_redshift
    zmax = P.max_redshift if max_redshift is None else max_redshift
    out = {}
    for i, b in enumerate(linking_length_factors):
        if i == 0:
            Kf, total = h.pipeline_run_np_random(g, p, zmin, zmax + 0.02, zmax)
        else:
            Kf, total = h.pipeline_rerun_np_random(b, rich, od)
        if i == 0:
            N_row = h.pipeline_fetch_counts(len(g))   # stage 0 does not depend on the swept parameters
        out[float(b)] = Catalogue(g, *h.pipeline_fetch(Kf, total), N_row=N_row)
    return out
