"""``compare_clusters`` of /root/reference/FoF/catalog_matching.py:21-111 with the same arguments and return values; the
nearest-cluster search runs on the GPU (fofb_match_catalog).  The sqlite ingestion and plotting of that script's
``__main__`` block are out of scope."""
import logging

import numpy as np

from . import _lib
from . import params as P
from .units import as_float

logger = logging.getLogger(__name__)


def match_rows(sample_arr, catalog, matching_distance, device=0, z_cut=0.1):
    """Device call behind compare_clusters: (match_row[C], separation_deg[C], n_matched) over ALL catalogue rows
    (-2 = outside the sample's redshift range, -1 = unmatched, else the matched sample row)."""
    p = _lib.default_params(H0=P.cosmo.H0, Om0=P.cosmo.Om0, Ode0=P.cosmo.Ode0)
    s = np.ascontiguousarray(np.asarray(sample_arr, dtype=np.float64)[:, :3])
    c = np.ascontiguousarray(np.asarray(catalog)[:, :3].astype(np.float64))
    return _lib.get_handle(device).match_catalog(s, c, p, as_float(matching_distance, "Mpc/littleh"), z_cut)


def compare_clusters(sample, catalog, matching_distance, richness_plot=False, device=0):
    """sample: list of Cluster-like objects (ra, dec, z, richness) or an (S, 4) array of those columns; catalog: (C, >=4)
    array [ra, dec, z, richness, ...]; matching_distance: h^-1 Mpc (float or Quantity).

    Returns (sample_matched, catalog_matching_idx) exactly as catalog_matching.py:95-111: the matched sample rows in
    catalogue order (rows containing a zero dropped, :95) and a 0/1 array over the catalogue rows that pass the
    redshift filter of :66-68.  ``richness_plot`` is accepted for signature compatibility; nothing is drawn."""
    if isinstance(sample, np.ndarray):
        sample = np.asarray(sample, dtype=np.float64)[:, :4]
    else:
        sample = np.array([[c.ra, c.dec, c.z, c.richness] for c in sample], dtype=np.float64).reshape(-1, 4)
    if len(sample) == 0:
        raise ValueError("compare_clusters: the sample is empty")
    catalog = np.asarray(catalog)
    rows, _sep, n_matched = match_rows(sample, catalog, matching_distance, device=device)
    kept = rows != -2
    logger.info("No. of galaxies in sample: %d and No. of galaxies in catalog: %d", len(sample), int(kept.sum()))
    rows = rows[kept]
    catalog_matching_idx = (rows >= 0).astype(np.float64)
    sample_matched = sample[rows[rows >= 0]]
    sample_matched = sample_matched[np.all(sample_matched != 0, axis=1)]
    if kept.sum():
        logger.info("%d/%d matched. %s %% of candidates matched", len(sample_matched), int(kept.sum()),
                    len(sample_matched) / kept.sum() * 100)
    return sample_matched, catalog_matching_idx
