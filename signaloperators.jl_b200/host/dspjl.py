"""Host-side stand-in for the parts of DSP.jl (0.6.x) the sink path calls.

In the Julia integration these calls stay in DSP.jl (filter *design* is host
code and is not replaced: SURVEY.md §1 "L5/L4 … filter design stay").  There is
no Julia in this image, so the Python host mirror needs its own design code to
produce the SOS coefficients / FIR taps that `libsignalops_cuda.so` consumes.
It is written from the published algorithms (marked
[RECALLED] there) and is cross-checked in tests against scipy.signal and the
independent restatement in `oracle/dspjl_ref.py`.

Reference call sites that reach these functions:
  src/filters.jl:10-11   digitalfilter(design(args...,fs=fs), method)
  src/filters.jl:94      DF2TFilter(h)            (-> SecondOrderSections)
  src/reformatting.jl:92-99  resample_filter / FIRFilter / timedelay / setphase!
  src/filters.jl:187,192,248 DSP.outputlength
"""
from __future__ import annotations

import cmath
import math
from fractions import Fraction

import numpy as np

__all__ = ["Lowpass", "Highpass", "Bandpass", "Bandstop", "Butterworth",
           "Chebyshev1", "digitalfilter", "ZeroPoleGain", "Biquad",
           "SecondOrderSections", "PolynomialRatio", "to_sos",
           "resample_filter", "FIRFilter", "rationalize"]


# ----------------------------------------------------------------------------
# response types (normalised so that 1.0 == Nyquist, as DSP.jl does)

class FilterType:
    pass


def _norm(w, fs):
    w = float(w)
    f = 2.0 * w / fs
    if not (0.0 < f < 1.0):
        raise ValueError("frequencies must be positive and below the Nyquist frequency")
    return f


class Lowpass(FilterType):
    def __init__(self, w, fs=2):
        self.w = _norm(w, fs)


class Highpass(FilterType):
    def __init__(self, w, fs=2):
        self.w = _norm(w, fs)


class Bandpass(FilterType):
    def __init__(self, w1, w2, fs=2):
        if not w1 < w2:
            raise ValueError("w1 must be less than w2")
        self.w1, self.w2 = _norm(w1, fs), _norm(w2, fs)


class Bandstop(FilterType):
    def __init__(self, w1, w2, fs=2):
        if not w1 < w2:
            raise ValueError("w1 must be less than w2")
        self.w1, self.w2 = _norm(w1, fs), _norm(w2, fs)


# ----------------------------------------------------------------------------
# coefficient containers

class ZeroPoleGain:
    def __init__(self, z, p, k):
        self.z = [complex(v) for v in z]
        self.p = [complex(v) for v in p]
        self.k = float(k)


class Biquad:
    """H(z) = (b0 + b1 z^-1 + b2 z^-2) / (1 + a1 z^-1 + a2 z^-2)."""

    def __init__(self, b0, b1, b2, a1, a2):
        self.b0, self.b1, self.b2, self.a1, self.a2 = (float(v) for v in (b0, b1, b2, a1, a2))

    def astuple(self):
        return (self.b0, self.b1, self.b2, self.a1, self.a2)


class SecondOrderSections:
    def __init__(self, biquads, g):
        self.biquads = list(biquads)
        self.g = float(g)

    def coef_table(self):
        """Row-major (M,5) array [b0 b1 b2 a1 a2] — the layout the C ABI takes."""
        return np.array([b.astuple() for b in self.biquads], dtype=np.float64).reshape(-1, 5)


class PolynomialRatio:
    def __init__(self, b, a):
        b = np.atleast_1d(np.asarray(b, dtype=np.float64))
        a = np.atleast_1d(np.asarray(a, dtype=np.float64))
        self.b = b / a[0]
        self.a = a / a[0]


# ----------------------------------------------------------------------------
# analog prototypes (DSP.jl Filters/design.jl)

def _snap1(v, npoles):
    """DSP.jl snaps a product that is 1 to within np*eps to exactly 1."""
    r = complex(v).real
    return 1.0 if abs(r - 1.0) < npoles * (abs(r) * 2.0 ** -52) else r


def Butterworth(n):
    n = int(n)
    if n <= 0:
        raise ValueError("n must be positive")
    poles = []
    for i in range(1, n // 2 + 1):
        w = (2 * i - 1) / (2 * n)
        pole = complex(-math.sin(math.pi * w), math.cos(math.pi * w))
        poles += [pole, pole.conjugate()]
    if n % 2:
        poles.append(complex(-1.0, 0.0))
    proto = ZeroPoleGain([], poles, 1.0)
    proto.spec = ("butterworth", n)
    return proto


def Chebyshev1(n, ripple):
    n = int(n)
    if n <= 0:
        raise ValueError("n must be positive")
    if ripple < 0:
        raise ValueError("ripple must be non-negative")
    eps = math.sqrt(10.0 ** (ripple / 10.0) - 1.0)
    mu = math.asinh(1.0 / eps) / n
    b, c = -math.sinh(mu), math.cosh(mu)
    poles = []
    for i in range(1, n // 2 + 1):
        w = (2 * i - 1) / (2 * n)
        pole = complex(b * math.sin(math.pi * w), c * math.cos(math.pi * w))
        poles += [pole, pole.conjugate()]
    if n % 2:
        w = (2 * (n // 2) + 1) / (2 * n)
        poles.append(complex(b * math.sin(math.pi * w), 0.0))
    k = 1.0
    for i in range(1, n // 2 + 1):
        k *= abs(poles[2 * i - 1]) ** 2
    if n % 2 == 0:
        k /= math.sqrt(1.0 + eps * eps)
    else:
        k *= (-poles[-1]).real
    proto = ZeroPoleGain([], poles, k)
    proto.spec = ("chebyshev1", n, float(ripple))
    return proto


# ----------------------------------------------------------------------------
# prototype -> digital (prewarp, frequency transform, bilinear with fs=2)

def _prewarp(w):
    return 4.0 * math.tan(math.pi * w / 2.0)


def _transform(ftype, proto):
    z, p, k = proto.z, proto.p, proto.k
    if isinstance(ftype, Lowpass):
        w = _prewarp(ftype.w)
        return ZeroPoleGain([w * v for v in z], [w * v for v in p],
                            k * w ** (len(p) - len(z)))
    if isinstance(ftype, Highpass):
        w = _prewarp(ftype.w)
        n = max(len(p), len(z))
        newz, newp = [0j] * n, [0j] * n
        num = 1.0 + 0j
        for i, v in enumerate(z):
            num *= -v
            newz[i] = w / v
        den = 1.0 + 0j
        for i, v in enumerate(p):
            den *= -v
            newp[i] = w / v
        return ZeroPoleGain(newz, newp, k * _snap1(num, len(p)) / _snap1(den, len(p)))
    w1, w2 = _prewarp(ftype.w1), _prewarp(ftype.w2)
    bw = w2 - w1
    if isinstance(ftype, Bandpass):
        ncommon = min(len(z), len(p))
        newz = [0j] * (2 * len(z) + len(p) - ncommon)
        newp = [0j] * (2 * len(p) + len(z) - ncommon)
        for (src, dst) in ((p, newp), (z, newz)):
            for i, v in enumerate(src):
                b = v * (bw / 2.0)
                pm = cmath.sqrt(b * b - w2 * w1)
                dst[2 * i] = b + pm
                dst[2 * i + 1] = b - pm
        return ZeroPoleGain(newz, newp, k * bw ** (len(p) - len(z)))
    if isinstance(ftype, Bandstop):
        n = max(len(z), len(p))
        newz, newp = [0j] * (2 * n), [0j] * (2 * n)
        num = 1.0 + 0j
        for i, v in enumerate(z):
            num *= -v
            b = (bw / 2.0) / v
            pm = cmath.sqrt(b * b - w2 * w1)
            newz[2 * i], newz[2 * i + 1] = b - pm, b + pm
        den = 1.0 + 0j
        for i, v in enumerate(p):
            den *= -v
            b = (bw / 2.0) / v
            pm = cmath.sqrt(b * b - w2 * w1)
            newp[2 * i], newp[2 * i + 1] = b - pm, b + pm
        npm = cmath.sqrt(-complex(w2 * w1))
        for i in range(len(z), n):
            newz[2 * i], newz[2 * i + 1] = -npm, npm
        for i in range(len(p), n):
            newp[2 * i], newp[2 * i + 1] = -npm, npm
        return ZeroPoleGain(newz, newp, k * _snap1(num, len(p)) / _snap1(den, len(p)))
    raise TypeError(f"unknown response type {ftype!r}")


def _bilinear(f, fs=2.0):
    n = max(len(f.p), len(f.z))
    z = [complex(-1.0)] * n
    p = []
    num = 1.0 + 0j
    for i, v in enumerate(f.z):
        z[i] = (2 + v / fs) / (2 - v / fs)
        num *= (2 * fs - v)
    den = 1.0 + 0j
    for v in f.p:
        p.append((2 + v / fs) / (2 - v / fs))
        den *= (2 * fs - v)
    return ZeroPoleGain(z, p, f.k * (num / den).real)


def digitalfilter(ftype, proto):
    """DSP.digitalfilter(ftype, proto) for IIR prototypes -> ZeroPoleGain (z-plane)."""
    return _bilinear(_transform(ftype, proto), 2.0)


# ----------------------------------------------------------------------------
# ZPK -> SOS (DSP.jl Filters/coefficients.jl)

def _split_real_complex(vals):
    cplx, real = [], []
    pending = []
    for v in vals:
        if abs(v.imag) <= 1e-14 * max(1.0, abs(v)):
            real.append(v.real)
        else:
            pending.append(v)
    pos = sorted((v for v in pending if v.imag > 0), key=lambda v: (v.real, v.imag))
    neg = sorted((v for v in pending if v.imag < 0), key=lambda v: (v.real, -v.imag))
    if len(pos) != len(neg):
        raise ValueError("complex roots could not be matched to their conjugates")
    for a, b in zip(pos, neg):
        if abs(a - b.conjugate()) > 1e-8 * max(1.0, abs(a)):
            raise ValueError("complex roots could not be matched to their conjugates")
        cplx.append(a)
    return cplx, real


def _take_closest(pool, target):
    j = min(range(len(pool)), key=lambda i: abs(pool[i] - target))
    return pool.pop(j)


def _poly2(roots):
    """Real monic polynomial coefficients (c1, c2) of z^2 + c1 z + c2 from ≤2 roots
    (written in z^-1 form: 1 + c1 z^-1 + c2 z^-2)."""
    if len(roots) == 0:
        return 0.0, 0.0
    if len(roots) == 1:
        return -complex(roots[0]).real, 0.0
    r0, r1 = complex(roots[0]), complex(roots[1])
    return -(r0 + r1).real, (r0 * r1).real


def zpk_to_sos(f):
    """convert(SecondOrderSections, ::ZeroPoleGain): poles nearest the unit circle
    are paired first with their closest zeros; sections are emitted in reverse so
    the least-damped section runs last; the gain is kept aside as `g`."""
    if len(f.z) > len(f.p):
        raise ValueError("ZeroPoleGain must not have more zeros than poles")
    cz, rz = _split_real_complex(f.z)
    cp, rp = _split_real_complex(f.p)
    cp.sort(key=lambda v: abs(abs(v) - 1.0))
    rp.sort(key=lambda v: abs(abs(v) - 1.0))

    groups = []  # (zeros, poles) per section, most-resonant first
    cz_pool, rz_pool = list(cz), list(rz)
    for pole in cp:
        if cz_pool:
            zc = _take_closest(cz_pool, pole)
            zeros = [zc, zc.conjugate()]
        else:
            zeros = []
            for _ in range(2):
                if rz_pool:
                    zeros.append(complex(_take_closest(rz_pool, pole)))
        groups.append((zeros, [pole, pole.conjugate()]))
    rp_pool = list(rp)
    while len(rp_pool) >= 2:
        p0 = rp_pool.pop(0)
        p1 = rp_pool.pop(0)
        if cz_pool:
            zc = _take_closest(cz_pool, complex(p0))
            zeros = [zc, zc.conjugate()]
        else:
            zeros = []
            for tgt in (p0, p1):
                if rz_pool:
                    zeros.append(complex(_take_closest(rz_pool, complex(tgt))))
        groups.append((zeros, [complex(p0), complex(p1)]))
    last = None
    if rp_pool:
        p0 = rp_pool.pop(0)
        zeros = [complex(_take_closest(rz_pool, complex(p0)))] if rz_pool else []
        last = (zeros, [complex(p0)])
    if cz_pool or rz_pool:
        raise ValueError("could not assign every zero to a section")

    sections = []
    if last is not None:
        sections.append(last)
    sections += list(reversed(groups))
    biquads = []
    for zeros, poles in sections:
        b1, b2 = _poly2(zeros)
        a1, a2 = _poly2(poles)
        biquads.append(Biquad(1.0, b1, b2, a1, a2))
    return SecondOrderSections(biquads, f.k)


def to_sos(h):
    """What `DF2TFilter(h)` (src/filters.jl:94) runs for each coefficient type."""
    if isinstance(h, SecondOrderSections):
        return h
    if isinstance(h, ZeroPoleGain):
        return zpk_to_sos(h)
    if isinstance(h, Biquad):
        return SecondOrderSections([h], 1.0)
    if isinstance(h, PolynomialRatio):
        if len(h.a) <= 3 and len(h.b) <= 3:
            b = list(h.b) + [0.0] * (3 - len(h.b))
            a = list(h.a) + [0.0] * (3 - len(h.a))
            return SecondOrderSections([Biquad(b[0], b[1], b[2], a[1], a[2])], 1.0)
        # order n: DSP.jl runs the order-n DF2T recurrence (src/filters.jl:89-95, App. B.2).  The device runs
        # biquad cascades, so the ratio is factored (roots of numerator and denominator -> zero-pole-gain ->
        # the same pairing as `convert(SecondOrderSections, ::ZeroPoleGain)`): the same transfer function,
        # equal to the direct form up to rounding (and better conditioned for high orders).
        b = np.trim_zeros(np.asarray(h.b, dtype=np.float64), "f")
        lead = len(h.b) - len(b)                       # leading zeros of b: pure delays z^-lead
        if len(b) == 0:
            return SecondOrderSections([Biquad(0.0, 0.0, 0.0, 0.0, 0.0)], 1.0)
        # H = b[0] * prod(1 - z_i z^-1) / prod(1 - p_i z^-1) * z^-lead; roots at the origin are factors of 1, so the
        # two lists are padded with them to equal length (no delay is introduced by that)
        z = list(np.roots(b))
        pz = list(np.roots(np.asarray(h.a, dtype=np.float64)))
        n = max(len(z), len(pz))
        z += [0.0] * (n - len(z))
        pz += [0.0] * (n - len(pz))
        sos = zpk_to_sos(ZeroPoleGain(z, pz, float(b[0]))) if n else SecondOrderSections([], float(b[0]))
        delays = [Biquad(0.0, 0.0, 1.0, 0.0, 0.0)] * (lead // 2) + [Biquad(0.0, 1.0, 0.0, 0.0, 0.0)] * (lead % 2)
        return SecondOrderSections(list(sos.biquads) + delays, sos.g)
    raise TypeError(f"not a filter coefficient object: {h!r}")


# ----------------------------------------------------------------------------
# resampling filters (DSP.jl Filters/stream_filt.jl)

def rationalize(x):
    """What src/reformatting.jl:103-111 needs from Julia's `rationalize`: the ratio
    p//q with max(p,q) <= 3 if `x` is within eps(x) of one, else None."""
    tol = abs(x) * 2.0 ** -52
    for q in (1, 2, 3):
        for p in (1, 2, 3):
            if math.gcd(p, q) == 1 and abs(p / q - x) <= tol:
                return Fraction(p, q)
    return None


def _kaiserord(transitionwidth, attenuation=60.0):
    n = math.ceil((attenuation - 7.95) / (math.pi * 2.285 * transitionwidth)) + 1
    if attenuation > 50:
        beta = 0.1102 * (attenuation - 8.7)
    elif attenuation >= 21:
        beta = 0.5842 * (attenuation - 21) ** 0.4 + 0.07886 * (attenuation - 21)
    else:
        beta = 0.0
    return n, beta / math.pi


def _kaiser(n, alpha):
    """DSP.Windows.kaiser(n, α): I0(πα·sqrt(1-(2k/(n-1)-1)^2)) / I0(πα)."""
    k = np.arange(n, dtype=np.float64)
    t = 2.0 * k / (n - 1) - 1.0
    return np.i0(math.pi * alpha * np.sqrt(np.maximum(0.0, 1.0 - t * t))) / np.i0(math.pi * alpha)


def _fir_lowpass_window(cutoff, window):
    """digitalfilter(Lowpass(cutoff), FIRWindow(window)) — windowed sinc scaled to
    unit DC gain."""
    n = len(window)
    k = np.arange(n, dtype=np.float64) - (n - 1) / 2.0
    h = cutoff * np.sinc(cutoff * k) * window
    return h / h.sum()


def resample_filter(rate, Nphases=32, rel_bw=1.0, attenuation=60.0):
    if isinstance(rate, Fraction):
        Nphases = rate.numerator
        decimation = rate.denominator
        f_nyq = min(1.0 / Nphases, 1.0 / decimation)
    else:
        f_nyq = 1.0 / Nphases if rate >= 1.0 else rate / Nphases
    cutoff = f_nyq * rel_bw
    tw = cutoff * 0.2
    hlen, alpha = _kaiserord(tw, attenuation)
    hlen = Nphases * math.ceil(hlen / Nphases)
    if hlen % 2 == 0:
        hlen += 1
    h = _fir_lowpass_window(cutoff, _kaiser(hlen, alpha))
    return h * Nphases


def _taps2pfb(h, nphases):
    hlen = len(h)
    tapsper = -(-hlen // nphases)
    padded = np.zeros(tapsper * nphases, dtype=np.float64)
    padded[:hlen] = h
    # pfb[row, col]: column φ holds taps h[φ], h[φ+N], ... stored reversed (row
    # tapsPerφ-1 is the first tap) so a dot with the input window is a convolution.
    pfb = padded.reshape(tapsper, nphases)[::-1, :].copy()
    return pfb


class FIRFilter:
    """FIRFilter(h, ratio): picks the DSP.jl kernel from the type/value of `ratio`.

    kind: 'standard' | 'interpolator' | 'decimator' | 'rational' | 'arbitrary'
    State set by `setphase(timedelay())` exactly once after construction
    (src/reformatting.jl:92-99).  0-based `phase0` / `deficit0` describe where
    the first output sits; `phi_acc0` is the Float64 accumulator for 'arbitrary'.
    """

    def __init__(self, h, ratio=1):
        self.h = np.asarray(h, dtype=np.float64)
        self.hlen = len(self.h)
        self.input_deficit = 1      # DSP.jl kernels start with inputDeficit = 1
        if isinstance(ratio, (float, np.floating)):
            self.kind = "arbitrary"
            self.rate = float(ratio)
            self.nphases = 32
            self.pfb = _taps2pfb(self.h, self.nphases)
            dh = np.append(np.diff(self.h), 0.0)
            self.dpfb = _taps2pfb(dh, self.nphases)
            self.tapsper = self.pfb.shape[0]
            self.delta = self.nphases / self.rate
            self.phi_acc = 1.0
            self.ratio = self.rate
        else:
            r = Fraction(ratio)
            self.ratio = r
            self.rate = float(r)
            p, q = r.numerator, r.denominator
            if p == 1 and q == 1:
                self.kind = "standard"
                self.nphases, self.tapsper = 1, self.hlen
            elif p == 1:
                self.kind = "decimator"
                self.decimation = q
                self.nphases, self.tapsper = 1, self.hlen
            elif q == 1:
                self.kind = "interpolator"
                self.nphases = p
                self.decimation = 1
                self.phi_idx = 1
                self.pfb = _taps2pfb(self.h, p)
                self.tapsper = self.pfb.shape[0]
            else:
                self.kind = "rational"
                self.nphases = p
                self.decimation = q
                self.pfb = _taps2pfb(self.h, p)
                self.tapsper = self.pfb.shape[0]
                self.phi_idx = 1
        if self.kind in ("standard", "decimator"):
            self.hrev = self.h[::-1].copy()

    def timedelay(self):
        if self.kind in ("standard", "decimator"):
            return (self.hlen - 1) / 2.0
        return (self.hlen - 1) / (2.0 * self.nphases)

    def setphase(self, tau):
        if self.kind in ("standard", "decimator"):
            self.input_deficit += int(round(tau))
            return self
        frac, whole = math.modf(tau)
        self.input_deficit += int(round(whole))
        if self.kind == "arbitrary":
            self.phi_acc = frac * self.nphases + 1.0
        elif self.kind in ("rational", "interpolator"):
            self.phi_idx = int(round(frac * self.nphases + 1.0))
        return self

    def outputlength(self, inputlength):
        n = inputlength - self.input_deficit + 1
        if self.kind == "standard":
            return max(0, n)
        if self.kind == "interpolator":
            return max(0, n * self.nphases - (self.phi_idx - 1)) if n > 0 else 0
        if self.kind == "decimator":
            return max(0, -(-n // self.decimation)) if n > 0 else 0
        if self.kind == "rational":
            if n <= 0:
                return 0
            p, q = self.nphases, self.decimation
            return max(0, -(-(n * p - (self.phi_idx - 1)) // q))
        return max(0, int(math.ceil(n * self.rate)))
