"""Tiny unit-tagged float used where the reference returns astropy Quantities.

astropy is not a dependency of this package.  ``Q`` is a float subclass with ``.value`` and ``.unit``
so that ``CleanedCluster.properties`` (cluster.py:127-129 reads ``.value``) and plain arithmetic both
work.  Inputs may be floats, ``Q`` or real astropy Quantities (anything with ``.to``/``.value``).
"""


class Q(float):
    __slots__ = ("unit",)

    def __new__(cls, value, unit=""):
        obj = super().__new__(cls, value)
        obj.unit = unit
        return obj

    @property
    def value(self):
        return float(self)

    def __repr__(self):
        return f"<Q {float(self)!r} {self.unit}>"

    def __reduce__(self):
        return (Q, (float(self), self.unit))


def as_float(x, astropy_unit=None):
    """Float value of a number / Q / astropy Quantity (converted to ``astropy_unit`` when possible)."""
    if hasattr(x, "to") and astropy_unit is not None and not isinstance(x, Q):
        try:
            return float(x.to(astropy_unit).value)
        except Exception:
            return float(x.value)
    if hasattr(x, "value") and not isinstance(x, Q):
        return float(x.value)
    return float(x)
