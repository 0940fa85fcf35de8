"""Minimal stand-in for the subset of ``astropy.units`` the reference FoF path touches.

TEST INFRASTRUCTURE ONLY (oracle).  astropy is not installed in this image and its
source is not under /root/reference, so this file *restates* the published behaviour of
astropy >= 4.0 units for exactly the expressions used at these call sites:

  params.py:23-25, analysis/methods.py:30-32,57-65,85-87, fof.py:73,165-175,190,
  analysis/mass_estimator.py:31-33,66-71,90-104, cluster.py:111-118.

A unit is (scale to SI, exponents over the bases m, s, kg, rad, littleh).  A Quantity is
(value ndarray/float, unit).  ``.to()`` multiplies the value by scale_from/scale_to, with
the two equivalencies the reference uses: ``with_H0`` (littleh <-> H0/100) and
``dimensionless_angles`` (rad == 1).  ULP-level rounding may differ from real astropy;
thresholds derived from these numbers are audited for margin in the parity tests.
"""
from fractions import Fraction
import math
import numpy as np

_BASES = ("m", "s", "kg", "rad", "littleh")


def _frac(p):
    if isinstance(p, Fraction):
        return p
    if isinstance(p, (int, np.integer)):
        return Fraction(int(p))
    return Fraction(float(p)).limit_denominator(12)


class Unit:
    __array_priority__ = 20000
    __array_ufunc__ = None

    def __init__(self, scale, dims, name=None):
        self.scale = float(scale)
        self.dims = tuple(_frac(d) for d in dims)
        self.name = name

    # ---- algebra -------------------------------------------------------------------
    def _combine(self, other, sign):
        if isinstance(other, Unit):
            dims = tuple(a + sign * b for a, b in zip(self.dims, other.dims))
            scale = self.scale * other.scale if sign > 0 else self.scale / other.scale
            return Unit(scale, dims)
        return NotImplemented

    def __mul__(self, other):
        if isinstance(other, Unit):
            return self._combine(other, +1)
        if isinstance(other, Quantity):
            return Quantity(other.value, self * other.unit)
        return Quantity(other, self)

    def __rmul__(self, other):
        if isinstance(other, Quantity):
            return Quantity(other.value, other.unit * self)
        return Quantity(other, self)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return self._combine(other, -1)
        if isinstance(other, Quantity):
            return Quantity(1.0 / other.value, self / other.unit)
        return Quantity(1.0 / np.asarray(other, dtype=float), self)

    def __rtruediv__(self, other):
        if isinstance(other, Quantity):
            return Quantity(other.value, other.unit / self)
        return Quantity(other, dimensionless_unscaled / self)

    def __pow__(self, p):
        p = _frac(p)
        return Unit(self.scale ** float(p), tuple(d * p for d in self.dims))

    def is_dimensionless(self):
        return all(d == 0 for d in self.dims)

    def __eq__(self, other):
        return isinstance(other, Unit) and self.dims == other.dims and self.scale == other.scale

    def __hash__(self):
        return hash((self.scale, self.dims))

    def __repr__(self):
        if self.name:
            return f"Unit({self.name})"
        body = " ".join(f"{b}^{d}" for b, d in zip(_BASES, self.dims) if d != 0)
        return f"Unit({self.scale:g} {body})"


def _base(i):
    return tuple(1 if j == i else 0 for j in range(len(_BASES)))


dimensionless_unscaled = Unit(1.0, (0, 0, 0, 0, 0), "")
m = Unit(1.0, _base(0), "m")
s = Unit(1.0, _base(1), "s")
kg = Unit(1.0, _base(2), "kg")
rad = Unit(1.0, _base(3), "rad")
littleh = Unit(1.0, _base(4), "littleh")

cm = Unit(1e-2, _base(0), "cm")
km = Unit(1e3, _base(0), "km")
g = Unit(1e-3, _base(2), "g")
# IAU 2012/2015: au = 149597870700 m; pc = au * 648000/pi  (astropy >= 4.0)
_pc_m = 149597870700.0 * 648000.0 / math.pi
pc = Unit(_pc_m, _base(0), "pc")
kpc = Unit(1e3 * _pc_m, _base(0), "kpc")
Mpc = Unit(1e6 * _pc_m, _base(0), "Mpc")
# IAU 2015 nominal GM_sun / CODATA-2018 G
M_sun = Unit(1.3271244e20 / 6.6743e-11, _base(2), "M_sun")
deg = degree = Unit(math.pi / 180.0, _base(3), "deg")
arcmin = Unit(math.pi / 10800.0, _base(3), "arcmin")
arcsec = Unit(math.pi / 648000.0, _base(3), "arcsec")
radian = rad
sr = steradian = Unit(1.0, (0, 0, 0, 2, 0), "sr")

_NAMES = {
    "m": m, "s": s, "kg": kg, "rad": rad, "km": km, "Mpc": Mpc, "kpc": kpc, "pc": pc,
    "deg": deg, "arcsec": arcsec, "arcmin": arcmin, "sr": sr, "M_sun": M_sun, "g": g,
    "cm": cm, "littleh": littleh,
}


def _parse(unit):
    """Accept a Unit or the simple 'a/b' strings the reference passes to .to()."""
    if isinstance(unit, Unit):
        return unit
    if isinstance(unit, str):
        num, *den = [t.strip() for t in unit.split("/")]
        out = _NAMES[num] if num not in ("", "1") else dimensionless_unscaled
        for d in den:
            out = out / _NAMES[d]
        return out
    raise TypeError(f"cannot interpret {unit!r} as a unit")


# ---- equivalencies (only the two used on the path) ---------------------------------------
class _Equivalency:
    def __init__(self, kind, h=None):
        self.kind = kind
        self.h = h


def with_H0(H0=None):
    """littleh <-> H0/(100 km/s/Mpc)  (analysis/methods.py:31, mass_estimator.py:33)."""
    h = H0.to(km / s / Mpc).value / 100.0
    return _Equivalency("with_H0", h=float(h))


def dimensionless_angles():
    """radian == dimensionless  (fof.py:171,431; mass_estimator.py:92)."""
    return _Equivalency("dimensionless_angles")


class UnitConversionError(ValueError):
    pass


def _conversion_factor(src, dst, equivalencies=None):
    src_d, dst_d = list(src.dims), list(dst.dims)
    factor = src.scale / dst.scale
    eqs = equivalencies if isinstance(equivalencies, (list, tuple)) else [equivalencies]
    for eq in eqs:
        if eq is None:
            continue
        if eq.kind == "with_H0" and src_d[4] != dst_d[4]:
            # value [X littleh^p] == value * h^p [X]
            factor = factor * eq.h ** float(src_d[4] - dst_d[4])
            src_d[4] = dst_d[4]
        if eq.kind == "dimensionless_angles" and src_d[3] != dst_d[3]:
            src_d[3] = dst_d[3]
    if src_d != dst_d:
        raise UnitConversionError(f"{src!r} is not convertible to {dst!r}")
    return factor


class Quantity:
    __array_priority__ = 10000

    def __init__(self, value, unit=dimensionless_unscaled):
        if isinstance(value, Quantity):
            unit = value.unit * unit if unit is not dimensionless_unscaled else value.unit
            value = value.value
        v = np.asarray(value, dtype=float)
        self.value = float(v) if v.ndim == 0 else v
        self.unit = _parse(unit)

    # ---- conversion ----------------------------------------------------------------
    def to(self, unit, equivalencies=None):
        unit = _parse(unit)
        f = _conversion_factor(self.unit, unit, equivalencies)
        return Quantity(self.value * f if f != 1.0 else self.value, unit)

    def to_value(self, unit=None, equivalencies=None):
        return self.value if unit is None else self.to(unit, equivalencies).value

    @property
    def si(self):
        return Quantity(self.value * self.unit.scale, Unit(1.0, self.unit.dims))

    def _same(self, other):
        """other's value expressed in self's unit (0 may carry any unit, like astropy)."""
        if isinstance(other, Quantity):
            if other.unit.dims == self.unit.dims and other.unit.scale == self.unit.scale:
                return other.value
            return other.to(self.unit).value
        if np.ndim(other) == 0 and other == 0:
            return other
        if self.unit.is_dimensionless():
            return np.asarray(other, dtype=float) / self.unit.scale
        raise UnitConversionError("can only combine quantities of equivalent units")

    # ---- arithmetic ------------------------------------------------------------------
    def __add__(self, o):
        return Quantity(self.value + self._same(o), self.unit)

    __radd__ = __add__

    def __sub__(self, o):
        return Quantity(self.value - self._same(o), self.unit)

    def __rsub__(self, o):
        return Quantity(self._same(o) - self.value, self.unit)

    def __mul__(self, o):
        if isinstance(o, Quantity):
            return Quantity(self.value * o.value, self.unit * o.unit)
        if isinstance(o, Unit):
            return Quantity(self.value, self.unit * o)
        return Quantity(self.value * o, self.unit)

    def __rmul__(self, o):
        if isinstance(o, Unit):
            return Quantity(self.value, o * self.unit)
        return Quantity(o * self.value, self.unit)

    def __truediv__(self, o):
        if isinstance(o, Quantity):
            return Quantity(self.value / o.value, self.unit / o.unit)
        if isinstance(o, Unit):
            return Quantity(self.value, self.unit / o)
        return Quantity(self.value / o, self.unit)

    def __rtruediv__(self, o):
        if isinstance(o, Unit):
            return Quantity(1.0 / self.value, o / self.unit)
        return Quantity(o / self.value, dimensionless_unscaled / self.unit)

    def __pow__(self, p):
        return Quantity(self.value ** p, self.unit ** p)

    def __neg__(self):
        return Quantity(-self.value, self.unit)

    def __abs__(self):
        return Quantity(np.abs(self.value), self.unit)

    def __float__(self):
        if not self.unit.is_dimensionless():
            raise TypeError("only dimensionless scalar quantities can be converted to float")
        return float(self.value * self.unit.scale)

    # ---- comparisons -----------------------------------------------------------------
    def __le__(self, o):
        return self.value <= self._same(o)

    def __lt__(self, o):
        return self.value < self._same(o)

    def __ge__(self, o):
        return self.value >= self._same(o)

    def __gt__(self, o):
        return self.value > self._same(o)

    def __eq__(self, o):
        try:
            return self.value == self._same(o)
        except UnitConversionError:
            return False

    def __ne__(self, o):
        return np.logical_not(self.__eq__(o))

    __hash__ = None

    def __array__(self, dtype=None, copy=None):
        # like the ndarray-subclass original: plain-array consumers see the bare values
        return np.asarray(self.value, dtype=dtype)

    # ---- container protocol ------------------------------------------------------------
    def __len__(self):
        return len(self.value)

    def __getitem__(self, i):
        return Quantity(self.value[i], self.unit)

    def __iter__(self):
        for v in self.value:
            yield Quantity(v, self.unit)

    @property
    def shape(self):
        return np.shape(self.value)

    def sum(self, axis=None, dtype=None, out=None, **kw):
        return Quantity(np.sum(self.value, axis=axis), self.unit)

    def mean(self, axis=None, dtype=None, out=None, **kw):
        return Quantity(np.mean(self.value, axis=axis), self.unit)

    # ---- numpy ufuncs ------------------------------------------------------------------
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs.get("out") is not None:
            return NotImplemented
        a = inputs[0]
        b = inputs[1] if len(inputs) > 1 else None
        qa = a if isinstance(a, Quantity) else None
        if ufunc is np.multiply:
            return a * b if qa is not None else b.__rmul__(a)
        if ufunc in (np.divide, np.true_divide):
            return a / b if qa is not None else b.__rtruediv__(a)
        if ufunc is np.add:
            return a + b if qa is not None else b.__radd__(a)
        if ufunc is np.subtract:
            return a - b if qa is not None else b.__rsub__(a)
        if ufunc is np.power:
            return a ** b
        if ufunc is np.sqrt:
            return Quantity(np.sqrt(a.value), a.unit ** Fraction(1, 2))
        if ufunc in (np.absolute, np.fabs):
            return abs(a)
        if ufunc is np.negative:
            return -a
        if ufunc in (np.sin, np.cos, np.tan):
            return ufunc(a.to(rad).value)
        cmp = {np.less_equal: "__le__", np.less: "__lt__", np.greater_equal: "__ge__",
               np.greater: "__gt__", np.equal: "__eq__", np.not_equal: "__ne__"}
        if ufunc in cmp:
            if qa is not None:
                return getattr(a, cmp[ufunc])(b)
            flip = {"__le__": "__ge__", "__lt__": "__gt__", "__ge__": "__le__",
                    "__gt__": "__lt__", "__eq__": "__eq__", "__ne__": "__ne__"}
            return getattr(b, flip[cmp[ufunc]])(a)
        return NotImplemented

    def __repr__(self):
        return f"<Quantity {self.value!r} {self.unit!r}>"
