"""astropy.constants subset: c and G (CODATA 2018, astropy >= 4.0).

Call sites: analysis/methods.py:57,85 (const.c), analysis/mass_estimator.py:32 (const.G).
"""
from . import units as u

c = u.Quantity(299792458.0, u.m / u.s)
G = u.Quantity(6.6743e-11, u.m ** 3 / (u.kg * u.s ** 2))
