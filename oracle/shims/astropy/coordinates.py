"""astropy.coordinates.SkyCoord subset: construction from (ra, dec), indexing, iteration and
``separation`` = Vincenty great-circle formula (astropy ``angular_separation``).

Call sites: fof.py:425-429, analysis/mass_estimator.py:74-88,120-131.
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
