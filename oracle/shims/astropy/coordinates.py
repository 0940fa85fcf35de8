"""astropy.coordinates.SkyCoord subset: construction from (ra, dec), indexing, iteration and
``separation`` = Vincenty great-circle formula (astropy ``angular_separation``).

``match_coordinates_sky`` restates astropy's nearest-neighbour match: unit-sphere Cartesian vectors, scipy cKDTree
(the tree astropy itself uses), d2d = Vincenty separation of the matched pair.

Call sites: fof.py:425-429, analysis/mass_estimator.py:74-88,120-131, catalog_matching.py:82-91.
"""
import numpy as np
from . import units as u


def angular_separation(lon1, lat1, lon2, lat2):
    sdlon = np.sin(lon2 - lon1)
    cdlon = np.cos(lon2 - lon1)
    slat1 = np.sin(lat1)
    slat2 = np.sin(lat2)
    clat1 = np.cos(lat1)
    clat2 = np.cos(lat2)
    num1 = clat2 * sdlon
    num2 = clat1 * slat2 - slat1 * clat2 * cdlon
    denominator = slat1 * slat2 + clat1 * clat2 * cdlon
    return np.arctan2(np.hypot(num1, num2), denominator)


class SkyCoord:
    def __init__(self, ra=None, dec=None, **kw):
        self.ra = ra if isinstance(ra, u.Quantity) else u.Quantity(ra, u.deg)
        self.dec = dec if isinstance(dec, u.Quantity) else u.Quantity(dec, u.deg)

    @property
    def shape(self):
        return np.shape(self.ra.value)

    def __len__(self):
        return len(self.ra.value)

    def __bool__(self):  # astropy ShapedLikeNDArray: true unless empty
        return int(np.size(self.ra.value)) > 0

    def unit_xyz(self):
        """UnitSphericalRepresentation.to_cartesian: x = cos(lat) cos(lon), y = cos(lat) sin(lon), z = sin(lat)."""
        lon = np.asarray(self.ra.to(u.rad).value, dtype=np.float64)
        lat = np.asarray(self.dec.to(u.rad).value, dtype=np.float64)
        return np.stack([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)], axis=-1)

    def __getitem__(self, i):
        return SkyCoord(self.ra[i], self.dec[i])

    def __iter__(self):
        for k in range(len(self)):
            yield self[k]

    def separation(self, other):
        lon1 = self.ra.to(u.rad).value
        lat1 = self.dec.to(u.rad).value
        lon2 = other.ra.to(u.rad).value
        lat2 = other.dec.to(u.rad).value
        sep = angular_separation(lon1, lat1, lon2, lat2)
        return u.Quantity(sep, u.rad).to(u.deg)


class Distance(u.Quantity):
    """Imported by catalog_matching.py:5 but never used there."""


def match_coordinates_sky(matchcoord, catalogcoord, nthneighbor=1, storekdtree="kdtree_sky"):
    """(idx, d2d, d3d): for every matchcoord the nearest catalogcoord on the sky (astropy
    coordinates/matching.py: match_coordinates_sky -> match_coordinates_3d on unit-sphere vectors)."""
    from scipy.spatial import cKDTree
    cat = catalogcoord.unit_xyz().reshape(-1, 3)
    if len(cat) == 0:
        raise ValueError("The catalog for coordinate matching cannot be a scalar or length-0.")
    xyz = matchcoord.unit_xyz()
    flat = xyz.reshape(-1, 3)
    dist, idx = cKDTree(cat).query(flat, nthneighbor)
    if nthneighbor > 1:
        dist, idx = dist[:, -1], idx[:, -1]
    idx = idx.reshape(xyz.shape[:-1])
    d2d = catalogcoord[idx].separation(matchcoord)
    return idx, d2d, dist.reshape(xyz.shape[:-1])
