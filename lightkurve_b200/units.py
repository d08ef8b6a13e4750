"""Minimal unit bookkeeping (astropy is not available in the build image).

The reference carries ``astropy.units.Quantity`` through the hot path only to label
results (``/root/reference/src/lightkurve/periodogram.py:793-794,969-975``).  This module
provides the small subset the shim needs: ``Unit`` algebra, ``Quantity`` (an ndarray
subclass with ``.value``/``.unit``/``.to``) and a ``Time`` array (days).  Objects coming
from a real astropy installation are accepted by duck typing (``.value`` / ``.unit``).
"""
import numpy as np

__all__ = ["Unit", "Quantity", "Time", "UnitConversionError", "Kelvin", "cm", "solRad", "solMass", "dex",
           "day", "d", "hour", "minute", "second", "s",
           "Hz", "hertz", "microhertz", "uHz", "electron", "dimensionless_unscaled", "percent", "ppt", "ppm",
           "K"]


class UnitConversionError(ValueError):
    pass


class Unit:
    """scale * prod(base ** power).  `bases` carries the physics ('s', 'electron', 'K');
    `parts` carries the named factors used for display only (e.g. {'uHz': -1, 'electron': 2})."""

    __array_ufunc__ = None          # ndarray <op> Unit defers to Unit.__r<op>__ (-> one Quantity)

    def __init__(self, bases=None, scale=1.0, name=None, parts=None):
        self.bases = {k: v for k, v in (bases or {}).items() if v != 0}
        self.scale = float(scale)
        if parts is None:
            parts = {} if name in (None, "") else {name: 1}
        self.parts = {k: v for k, v in parts.items() if v != 0}

    @property
    def name(self):
        return self.to_string()

    # -- algebra
    @staticmethod
    def _merge(a, b, sign):
        out = dict(a)
        for k, v in b.items():
            out[k] = out.get(k, 0) + sign * v
        return out

    def __mul__(self, other):
        if isinstance(other, Unit):
            return Unit(self._merge(self.bases, other.bases, +1), self.scale * other.scale,
                        parts=self._merge(self.parts, other.parts, +1))
        return Quantity(other, self)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Unit(self._merge(self.bases, other.bases, -1), self.scale / other.scale,
                        parts=self._merge(self.parts, other.parts, -1))
        return Quantity(1.0 / np.asarray(other, dtype=float), self)

    def __rtruediv__(self, other):
        inv = self ** -1
        if isinstance(other, (int, float)) and other == 1:
            return inv
        return Quantity(other, inv)

    def __pow__(self, p):
        return Unit({k: v * p for k, v in self.bases.items()}, self.scale ** p,
                    parts={k: v * p for k, v in self.parts.items()})

    # -- comparison / conversion
    def is_equivalent(self, other):
        other = _as_unit(other)
        keys = set(self.bases) | set(other.bases)
        return all(abs(self.bases.get(k, 0) - other.bases.get(k, 0)) < 1e-12 for k in keys)

    def factor_to(self, other):
        if other is self:
            return 1.0
        if not self.is_equivalent(other):
            raise UnitConversionError("'%s' and '%s' are not convertible" % (self, other))
        return self.scale / other.scale

    def __eq__(self, other):
        if not isinstance(other, Unit):
            other = _as_unit(other, strict=False)
            if other is None:
                return False
        return self.is_equivalent(other) and bool(np.isclose(self.scale, other.scale, rtol=1e-12))

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash((tuple(sorted(self.bases.items())), round(np.log10(self.scale), 9)))

    def __bool__(self):
        # astropy: bool(dimensionless_unscaled) is False (used at lightcurve.py:1283)
        return not (len(self.bases) == 0 and self.scale == 1.0)

    def to_string(self):
        def fmt(items):
            return " ".join("%s%s" % (k, "" if abs(v) == 1 else _fmt_pow(abs(v))) for k, v in items)
        num = sorted((k, v) for k, v in self.parts.items() if v > 0)
        den = sorted((k, v) for k, v in self.parts.items() if v < 0)
        if not num and not den:
            return ""
        s = fmt(num) if num else "1"
        if den:
            d_ = fmt(den)
            s += " / " + ("(%s)" % d_ if len(den) > 1 else d_)
        return s

    __str__ = to_string

    def __repr__(self):
        return 'Unit("%s")' % self.to_string()


def _fmt_pow(p):
    return "%d" % p if float(p).is_integer() else "%g" % p


second = s = Unit({"s": 1}, 1.0, "s")
minute = Unit({"s": 1}, 60.0, "min")
hour = Unit({"s": 1}, 3600.0, "h")
day = d = Unit({"s": 1}, 86400.0, "d")
Hz = hertz = Unit({"s": -1}, 1.0, "Hz")
microhertz = uHz = Unit({"s": -1}, 1e-6, "uHz")
electron = Unit({"electron": 1}, 1.0, "electron")
K = Kelvin = Unit({"K": 1}, 1.0, "K")
cm = Unit({"m": 1}, 1e-2, "cm")
solRad = Unit({"m": 1}, 6.957e8, "solRad")           # IAU 2015 nominal solar radius
solMass = Unit({"kg": 1}, 1.988409870698051e30, "solMass")
dex = Unit({"dex": 1}, 1.0, "dex")
dimensionless_unscaled = Unit({}, 1.0, "")
percent = Unit({}, 1e-2, "%")
ppt = Unit({}, 1e-3, "ppt")
ppm = Unit({}, 1e-6, "ppm")

_BY_NAME = {"s": second, "second": second, "min": minute, "minute": minute, "h": hour, "hour": hour,
            "d": day, "day": day, "Hz": Hz, "hertz": Hz, "uHz": microhertz, "microhertz": microhertz,
            "electron": electron, "e": electron, "": dimensionless_unscaled,
            "dimensionless": dimensionless_unscaled, "percent": percent, "%": percent, "ppt": ppt, "ppm": ppm,
            "K": K, "Kelvin": K, "cm": cm, "solRad": solRad, "solMass": solMass, "dex": dex, "electron/s": electron / second, "electron/second": electron / second,
            "electron / s": electron / second, "e/s": electron / second, "1/d": 1 / day, "1 / d": 1 / day}


def _as_unit(x, strict=True):
    """Unit from a Unit, a known string, None, or a foreign (astropy) unit via its string form."""
    if x is None:
        return dimensionless_unscaled
    if isinstance(x, Unit):
        return x
    if isinstance(x, str):
        if x in _BY_NAME:
            return _BY_NAME[x]
        if strict:
            raise ValueError("unknown unit string %r" % x)
        return None
    if hasattr(x, "to_string"):          # astropy unit
        st = x.to_string()
        if st in _BY_NAME:
            return _BY_NAME[st]
        if hasattr(x, "decompose"):
            try:
                dec = x.decompose()
                bases = {str(b): p for b, p in zip(dec.bases, dec.powers)}
                return Unit(bases, float(dec.scale), name=st)
            except Exception:
                pass
    if strict:
        raise ValueError("cannot interpret %r as a unit" % (x,))
    return None


_BOOL_UFUNCS = {np.less, np.less_equal, np.greater, np.greater_equal, np.equal, np.not_equal, np.isfinite,
                np.isnan, np.isinf, np.signbit, np.logical_and, np.logical_or, np.logical_not}
_SAME_UNIT = {np.add, np.subtract, np.maximum, np.minimum, np.fmax, np.fmin, np.hypot, np.remainder, np.fmod}


class Quantity(np.ndarray):
    """ndarray + unit.  Only the operations the hot-path shim needs are unit-aware."""

    __array_priority__ = 10000

    def __new__(cls, value, unit=None, dtype=None, copy=True):
        if hasattr(value, "unit") and hasattr(value, "value") and not isinstance(value, Quantity):
            value = Quantity(np.asarray(value.value), _as_unit(value.unit))      # astropy Quantity
        if isinstance(value, Quantity):
            if unit is None:
                unit = value.unit
            else:
                unit = _as_unit(unit)
                value = value.view(np.ndarray) * value.unit.factor_to(unit)
        unit = _as_unit(unit)
        arr = np.array(value, dtype=dtype, copy=copy, subok=False)
        if arr.dtype.kind in "iub":
            arr = arr.astype(float)
        obj = arr.view(cls)
        obj._unit = unit
        return obj

    def __array_finalize__(self, obj):
        self._unit = getattr(obj, "_unit", dimensionless_unscaled)

    @property
    def unit(self):
        return self._unit

    @property
    def value(self):
        v = self.view(np.ndarray)
        return v[()] if v.ndim == 0 else v

    def to(self, unit):
        unit = _as_unit(unit)
        return Quantity(self.view(np.ndarray) * self._unit.factor_to(unit), unit)

    def to_value(self, unit=None):
        return self.value if unit is None else self.to(unit).value

    def copy(self, order="C"):
        return Quantity(self.view(np.ndarray).copy(order), self._unit)

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        units = [x._unit if isinstance(x, Quantity) else None for x in inputs]
        raw = [x.view(np.ndarray) if isinstance(x, Quantity) else x for x in inputs]
        first = next(u for u in units if u is not None)
        unit = first
        if ufunc in _SAME_UNIT and len(inputs) == 2:
            # bring both operands to the unit of the first Quantity
            for i in range(2):
                if units[i] is not None and units[i] is not first and units[i] != first:
                    raw[i] = raw[i] * units[i].factor_to(first)
        elif ufunc is np.multiply:
            unit = (units[0] or dimensionless_unscaled) * (units[1] or dimensionless_unscaled)
        elif ufunc in (np.true_divide, np.divide, np.floor_divide):
            unit = (units[0] or dimensionless_unscaled) / (units[1] or dimensionless_unscaled)
        elif ufunc is np.sqrt:
            unit = first ** 0.5
        elif ufunc is np.square:
            unit = first ** 2
        elif ufunc is np.reciprocal:
            unit = 1 / first
        elif ufunc is np.power:
            unit = (units[0] or dimensionless_unscaled) ** float(np.asarray(raw[1]).ravel()[0])
        elif ufunc in _BOOL_UFUNCS:
            if len(inputs) == 2 and units[0] is not None and units[1] is not None and units[0] != units[1]:
                raw[1] = raw[1] * units[1].factor_to(units[0])
            unit = None
        if out is not None:
            kwargs["out"] = tuple(o.view(np.ndarray) if isinstance(o, Quantity) else o for o in out)
        res = getattr(ufunc, method)(*raw, **kwargs)
        if out is not None:
            res = out[0] if len(out) == 1 else out
            if isinstance(res, Quantity) and unit is not None:
                res._unit = unit
            return res
        if unit is None or not isinstance(res, np.ndarray) and not np.isscalar(res):
            return res
        if isinstance(res, tuple):
            return res
        q = np.asarray(res).view(Quantity)
        q._unit = unit
        return q

    def __mul__(self, other):
        if isinstance(other, Unit):
            return Quantity(self.view(np.ndarray), self._unit * other)
        return np.multiply(self, other)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Quantity(self.view(np.ndarray), self._unit / other)
        return np.true_divide(self, other)

    def __getitem__(self, item):
        out = super().__getitem__(item)
        if not isinstance(out, Quantity):
            out = np.asarray(out).view(type(self))
            out._unit = self._unit
            for attr in ("format", "scale"):
                if hasattr(self, attr):
                    setattr(out, attr, getattr(self, attr))
        return out

    def __repr__(self):
        return "<Quantity %s %s>" % (np.array2string(self.view(np.ndarray), separator=", ", threshold=8),
                                      self._unit.to_string())

    def __str__(self):
        return ("%s %s" % (self.view(np.ndarray), self._unit.to_string())).strip()

    def __format__(self, spec):
        if self.ndim == 0:
            return ("%s %s" % (format(float(self.view(np.ndarray)), spec), self._unit.to_string())).strip()
        return str(self)

    def __reduce__(self):
        base = super().__reduce__()
        return (base[0], base[1], base[2] + (self._unit,))

    def __setstate__(self, state):
        self._unit = state[-1]
        super().__setstate__(state[:-1])


class Time(Quantity):
    """Array of times in days (the reference assumes days: periodogram.py 'Caution' note)."""

    def __new__(cls, value, format="jd", scale="tdb"):
        if hasattr(value, "jd") and hasattr(value, "format") and not isinstance(value, Time):   # astropy Time
            fmt = value.format
            value = np.asarray(value.value, dtype=float) if fmt in ("jd", "bkjd", "btjd", "mjd") else \
                np.asarray(value.jd, dtype=float)
            format = fmt if fmt in ("jd", "bkjd", "btjd", "mjd") else "jd"
        if isinstance(value, Time):
            format, scale = value.format, value.scale
        obj = Quantity.__new__(cls, np.asarray(getattr(value, "value", value), dtype=float), day)
        obj.format = format
        obj.scale = scale
        return obj

    def __array_finalize__(self, obj):
        super().__array_finalize__(obj)
        self._unit = day
        self.format = getattr(obj, "format", "jd")
        self.scale = getattr(obj, "scale", "tdb")

    def copy(self, order="C"):
        return Time(self.view(np.ndarray).copy(order), self.format, self.scale)

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        # differences of times are plain day Quantities (TimeDelta analogue)
        inputs = tuple(Quantity(x.view(np.ndarray), day, copy=False) if isinstance(x, Time) else x for x in inputs)
        return Quantity.__array_ufunc__(inputs[0] if isinstance(inputs[0], Quantity) else inputs[1], ufunc, method,
                                        *inputs, out=out, **kwargs)

    def __repr__(self):
        return "<Time format=%s scale=%s %s>" % (self.format, self.scale,
                                                 np.array2string(self.view(np.ndarray), threshold=8))


def is_quantity(x):
    return hasattr(x, "unit") and hasattr(x, "value")
