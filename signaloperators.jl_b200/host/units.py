"""Units used by the operator API: Hz, kHz, s, ms, frames, kframes, dB, deg, rad.

Stand-in for SignalBase.Units / Unitful as re-exported by the reference at
src/SignalOperators.jl:11-15.  Only what the sink path needs: a quantity is a
(value, kind) pair; conversions follow SURVEY.md Appendix B.5
(`inframes(Int,t,fs) = floor(Int, inseconds(t)*inHz(fs))`, bare numbers are
seconds for times (src/util.jl:23-24) and Hz for rates, `x dB -> 10^(x/20)`
(src/numbers.jl:48-49)).
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np

__all__ = ["Hz", "kHz", "s", "ms", "frames", "kframes", "dB", "deg", "rad",
           "Quantity", "inHz", "inseconds", "inframes", "inradians",
           "maybeseconds", "gain_to_amplitude"]


class Quantity:
    """A number tagged with a dimension; `kind` in time/freq/frames/gain/angle."""
    __slots__ = ("value", "kind", "label")
    __array_ufunc__ = None

    def __init__(self, value, kind, label=""):
        self.value = value
        self.kind = kind
        self.label = label

    def _same(self, other):
        if not isinstance(other, Quantity) or other.kind != self.kind:
            raise TypeError(f"cannot combine {self!r} with {other!r}")

    def __add__(self, other):
        self._same(other)
        return Quantity(self.value + other.value, self.kind, self.label)

    def __sub__(self, other):
        self._same(other)
        return Quantity(self.value - other.value, self.kind, self.label)

    def __neg__(self):
        return Quantity(-self.value, self.kind, self.label)

    def __mul__(self, k):
        if isinstance(k, Quantity):
            raise TypeError("quantity products are not supported")
        return Quantity(self.value * k, self.kind, self.label)

    __rmul__ = __mul__

    def __truediv__(self, k):
        if isinstance(k, Quantity):
            self._same(k)
            return self.value / k.value
        return Quantity(self.value / k, self.kind, self.label)

    def __repr__(self):
        return f"{self.value} [{self.kind}]"


class Unit:
    """`5*s`, `s*5`, `10*ms` ... build quantities. Integer values convert
    through exact rationals (as Unitful does) so `10*ms` is the double nearest 1/100."""

    __array_ufunc__ = None      # make numpy scalars defer to __rmul__ (keeps Float32 gains Float32)

    def __init__(self, kind, scale, label):
        self.kind = kind
        self.scale = scale
        self.label = label

    def __rmul__(self, v):
        if isinstance(v, (bool, np.bool_)):
            raise TypeError("bool is not a number here")
        if isinstance(v, (int, np.integer)) and isinstance(self.scale, Fraction):
            q = Fraction(int(v)) * self.scale
            val = int(q) if q.denominator == 1 else float(q)
        elif isinstance(v, np.floating) and not isinstance(v, np.float64):
            # keep narrow float types (e.g. -10f0*dB, test/runtests.jl:722)
            val = type(v)(v * type(v)(float(self.scale)))
        else:
            val = v * float(self.scale)
        return Quantity(val, self.kind, self.label)

    __mul__ = __rmul__


Hz = Unit("freq", Fraction(1), "Hz")
kHz = Unit("freq", Fraction(1000), "kHz")
s = Unit("time", Fraction(1), "s")
ms = Unit("time", Fraction(1, 1000), "ms")
frames = Unit("frames", Fraction(1), "frames")
kframes = Unit("frames", Fraction(1000), "kframes")
dB = Unit("gain", Fraction(1), "dB")
rad = Unit("angle", Fraction(1), "rad")


class _Deg(Unit):
    def __rmul__(self, v):
        # v*pi/180 evaluated as (v*pi)/180 so that 180*deg == pi exactly
        return Quantity(v * math.pi / 180, "angle", "deg")

    __mul__ = __rmul__


deg = _Deg("angle", None, "deg")


def inHz(x, typ=None):
    """SignalBase.inHz: quantity -> Hz, bare number -> itself, missing -> None."""
    if x is None:
        return None
    if isinstance(x, Quantity):
        if x.kind != "freq":
            raise ValueError(f"expected a frequency, got {x!r}")
        x = x.value
    return float(x) if typ is float else x


def maybeseconds(x):
    """src/util.jl:23-24 — a bare number used as a time means seconds."""
    if isinstance(x, Quantity):
        return x
    return Quantity(x, "time", "s")


def inseconds(x, fs=None):
    if isinstance(x, Quantity):
        if x.kind == "time":
            return x.value
        if x.kind == "frames":
            if fs is None:
                return None
            return x.value / inHz(fs)
        raise ValueError(f"expected a time, got {x!r}")
    return x


def inframes(t, fs=None):
    """`inframes(Int,t,fs)`: floor(seconds*Hz); frame quantities pass through.
    Returns None (missing) when a time is given but the rate is unknown."""
    t = maybeseconds(t)
    if t.kind == "frames":
        return int(math.floor(t.value))
    if t.kind != "time":
        raise ValueError(f"expected a time, got {t!r}")
    if fs is None:
        return None
    return int(math.floor(t.value * inHz(fs)))


def inradians(phi, omega=None):
    """SignalBase.inradians: bare = radians; angle units converted; a time
    is `2π·seconds·Hz(ω)` (pinned by test/runtests.jl:74-79)."""
    if isinstance(phi, Quantity):
        if phi.kind == "angle":
            return float(phi.value)
        if phi.kind == "time":
            if omega is None:
                raise ValueError("a phase given as a time needs a frequency")
            return float(phi.value * inHz(omega) * 2 * math.pi)
        raise ValueError(f"expected a phase, got {phi!r}")
    return float(phi)


def gain_to_amplitude(q):
    """`uconvertrp(NoUnits, x dB)` = 10^(x/20) (src/numbers.jl:48-49)."""
    v = q.value
    if isinstance(v, np.floating) and not isinstance(v, np.float64):
        return type(v)(10.0 ** (float(v) / 20))
    return 10.0 ** (v / 20)
