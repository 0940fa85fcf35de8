"""Stand-in for the astropy subset used by the reference FoF hot path (oracle test infrastructure).

astropy is absent from this image; see oracle/README.md.  Only the names imported by
/root/reference/FoF/{params,fof,helper,cluster,mass_analysis}.py and analysis/*.py exist here.
"""
__version__ = "0.0-shim (restates astropy>=4.0: CODATA 2018, IAU 2015)"
from . import units, constants, cosmology, coordinates, stats, utils  # noqa: F401
