"""Drop-in for /root/reference/FoF/fof.py: the same functions, computed by libfofb200 on a B200.

    candidate_center_search  fof.py:24-95     ->  fofb_number_counts / fofb_candidate_order
    FoF                      fof.py:98-392    ->  fofb_fof_run / fofb_fof_finish / fofb_fof_fetch
    three_sigma              fof.py:396-496   ->  fofb_three_sigma
    interloper_removal       fof.py:499-516   ->  fofb_three_sigma (all clusters in one call)

There is no CPU path: without the CUDA library and a GPU every function raises.
"""
import numpy as np

from . import _lib
from . import params as _params_mod
from .cluster import Cluster
from .units import as_float

try:  # pandas is what the reference's pipeline passes in (pipeline.py:38-54)
    import pandas as pd
except Exception:  # pragma: no cover
    pd = None


def _lib_params(**kw):
    c = _params_mod.cosmo
    return _lib.default_params(H0=c.H0, Om0=c.Om0, Ode0=c.Ode0, **kw)


def _as_matrix(x):
    if pd is not None and isinstance(x, pd.DataFrame):
        x = x.values
    a = np.asarray(x)
    if a.dtype != np.float64:
        a = a.astype(np.float64)
    if a.ndim != 2:
        raise ValueError("expected a 2-D array of galaxies")
    return a


def candidate_center_search(galaxy_data, max_velocity, fname=None, path=None, device=0):
    """fof.py:24-95.  Returns (main_arr (n, C+1) with N(0.5) appended, candidate_centers sorted by N desc).

    Equal-N candidates are ordered by descending row index (what ``argsort(kind="stable")[::-1]``
    yields; the reference's default argsort leaves ties unspecified).  The two CSV files of
    fof.py:90-93 are written only when ``path`` is given.
    """
    galaxy_arr = _as_matrix(galaxy_data)
    h = _lib.get_handle(device)
    p = _lib_params(max_velocity=as_float(max_velocity, "km/s"))
    N, rows = h.number_counts(galaxy_arr, p)
    main_arr = np.column_stack([galaxy_arr, N.astype(np.float64)])
    candidate_centers = main_arr[rows]
    if path is not None and fname is not None:
        np.savetxt(path + fname + "_galaxy_data.csv", main_arr, delimiter=",")
        np.savetxt(path + fname + "_candidate_centers.csv", candidate_centers, delimiter=",")
    return main_arr, candidate_centers


def draw_overdensity_points(n_clusters, bbox, n_random=300):
    """The 2 x n_random uniforms per cluster of helper.py:197-202, from numpy's legacy GLOBAL generator
    (cluster by cluster: the RA block, then the Dec block), drawn in one vectorised call that walks the
    same C routine in the same order as the reference's per-cluster calls."""
    low = np.array([bbox[0], bbox[2]])[None, :, None]
    high = np.array([bbox[1], bbox[3]])[None, :, None]
    pts = np.random.uniform(low=low, high=high, size=(int(n_clusters), 2, int(n_random)))
    return np.ascontiguousarray(pts[:, 0, :]), np.ascontiguousarray(pts[:, 1, :])


def _clusters_from_csr(galaxy_data, candidate_centers, off, mem, cen, N, D):
    out = []
    for k in range(len(cen)):
        center = candidate_centers[cen[k]]
        c = Cluster(center, galaxy_data[mem[off[k]:off[k + 1]]])
        n = N[k]
        c.N = int(n) if float(n).is_integer() else n
        c.D = D[k]
        out.append(c)
    return out


def FoF(galaxy_data, candidate_centers, richness, overdensity, max_velocity=2000.0, linking_length_factor=0.1,
        virial_radius=1.5, device=0):
    """fof.py:98-392: candidate clusters around the centres (in priority order), overlap removal, BCG
    stage, N(0.5) / overdensity selection.  Returns ``list[Cluster]``.

    The uniform points of ``center_overdensity`` come from ``np.random`` (global legacy generator) exactly
    as in the reference, so seeding it before the call reproduces a seeded reference run.
    """
    g = _as_matrix(galaxy_data)
    c = _as_matrix(candidate_centers)
    h = _lib.get_handle(device)
    p = _lib_params(max_velocity=as_float(max_velocity, "km/s"), linking_length_factor=float(linking_length_factor),
                    virial_radius=as_float(virial_radius), richness=int(richness), overdensity=float(overdensity))
    h.fof_run(g, c, p)
    h.fof_finish_np_random()  # consumes np.random's MT19937 stream on the device, state written back
    return _clusters_from_csr(g, c, *h.fof_fetch(3))


def _pack(clusters):
    sizes = [len(c.galaxies) for c in clusters]
    off = np.zeros(len(clusters) + 1, dtype=np.int64)
    np.cumsum(sizes, out=off[1:])
    ncols = next((c.galaxies.shape[1] for c in clusters if len(c.galaxies)), 8)
    rows = [np.asarray(c.galaxies, dtype=np.float64).reshape(-1, ncols) for c in clusters]
    members = np.concatenate(rows, axis=0) if rows else np.zeros((0, ncols))
    return off, np.ascontiguousarray(members)


def three_sigma(center, member_galaxies, n, device=0):
    """fof.py:396-496 for one cluster: the cleaned member array (rows reordered by radial bin)."""
    mg = np.ascontiguousarray(np.asarray(member_galaxies, dtype=np.float64))
    h = _lib.get_handle(device)
    p = _lib_params(n_bins=int(n))
    off = np.array([0, len(mg)], dtype=np.int64)
    cen = np.array([[center.ra, center.dec, center.z]], dtype=np.float64)
    _, idx = h.three_sigma(off, mg, cen, p)
    return mg[idx]


def interloper_removal(data, device=0):
    """fof.py:499-516.  Mutates ``c.galaxies`` in place like the reference and keeps the clusters whose
    richness is still >= ``params.richness``.  Returns (cleaned_candidates, number_removed)."""
    print(f"Removing interlopers from {len(data)} clusters")
    if not len(data):
        return [], []
    h = _lib.get_handle(device)
    p = _lib_params(n_bins=10)
    off, members = _pack(data)
    cen = np.array([[c.ra, c.dec, c.z] for c in data], dtype=np.float64)
    out_off, idx = h.three_sigma(off, members, cen, p)
    number_removed = []
    for k, c in enumerate(data):
        initial_size = len(c.galaxies)
        c.galaxies = np.asarray(c.galaxies)[idx[out_off[k]:out_off[k + 1]]]
        number_removed.append(initial_size - len(c.galaxies))
    cleaned_candidates = [c for c in data if c.richness >= _params_mod.richness]
    print(f"Number of clusters: {len(cleaned_candidates)} returned")
    return cleaned_candidates, number_removed
