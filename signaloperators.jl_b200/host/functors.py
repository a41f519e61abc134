"""Generator functions the device can evaluate (the `Functor` hook of
src/functions.jl:65-66,88): Julia closures cannot cross a C ABI, so the GPU sink
recognises an enumerated set and materialises anything else on the host."""
from __future__ import annotations

import math

import numpy as np

from . import graph as G

FN_SIN, FN_COS, FN_SAW, FN_AFFINE_SIN, FN_AFFINE_COS, FN_IDENTITY = 1, 2, 3, 4, 5, 6


class Sawtooth(G.Functor):
    """README sound3: `ϕ -> ϕ/π - 1`"""

    def __call__(self, phi):
        return phi / math.pi - 1


class AffineSin(G.Functor):
    """README sound4: `ϕ -> 0.5sin(ϕ) + 0.5` generalised to a*sin(ϕ)+b"""

    def __init__(self, a=0.5, b=0.5):
        self.a, self.b = float(a), float(b)

    def __call__(self, phi):
        return self.a * math.sin(phi) + self.b


class AffineCos(G.Functor):
    def __init__(self, a=0.5, b=0.5):
        self.a, self.b = float(a), float(b)

    def __call__(self, phi):
        return self.a * math.cos(phi) + self.b


def is_sin(fn):
    """Stands in for dispatch on `typeof(sin)` (src/functions.jl:57-60)."""
    return fn is G.sin or fn is math.sin or fn is np.sin


def functor_code(fn):
    if is_sin(fn):
        return FN_SIN, 0.0, 0.0
    if fn is G.cos or fn is math.cos or fn is np.cos:
        return FN_COS, 0.0, 0.0
    if fn is G.identity:
        return FN_IDENTITY, 0.0, 0.0
    if isinstance(fn, Sawtooth):
        return FN_SAW, 0.0, 0.0
    if isinstance(fn, AffineSin):
        return FN_AFFINE_SIN, fn.a, fn.b
    if isinstance(fn, AffineCos):
        return FN_AFFINE_COS, fn.a, fn.b
    return None
