"""Independent restatement of the DSP.jl pieces behind the hot path (oracle side).

Deliberately does NOT reuse signalops.dspjl: IIR designs go through
scipy.signal's zpk designers (same prewarp + bilinear maths as DSP.jl,
SURVEY.md App. B.1), the resampling prototype through scipy.signal.firwin with
a Kaiser window (App. B.3), so a mistake in the host-side design code shows up
as a parity failure.  Filtering itself is oracle/cpu_ref.c.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from fractions import Fraction

import numpy as np
from scipy import signal as sps

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libcpuref.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE, "libcpuref.so"])
        L = C.CDLL(path)
        dp, i64 = C.POINTER(C.c_double), C.c_int64
        L.oracle_sos_filt.argtypes = [dp, dp, i64, dp, C.c_int, C.c_double, dp]
        L.oracle_sos_filt.restype = None
        L.oracle_fir_filt.argtypes = [dp, i64, C.POINTER(FirState), dp, i64]
        L.oracle_fir_filt.restype = i64
        L.oracle_iir_amplify_batch.argtypes = [dp, dp, i64, i64, C.c_int, dp, C.c_int, C.c_double,
                                               C.c_double, i64, C.c_int]
        L.oracle_iir_amplify_batch.restype = None
        L.oracle_resample_batch.argtypes = [dp, dp, i64, i64, i64, C.c_int, C.POINTER(FirState), i64, C.c_int]
        L.oracle_resample_batch.restype = None
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


# ---- IIR design -------------------------------------------------------------------

def design_zpk(kind, bounds_hz, fs, spec):
    """digitalfilter(kind(bounds...; fs), method) via scipy (src/filters.jl:10-11)."""
    btype = {"Lowpass": "lowpass", "Highpass": "highpass", "Bandpass": "bandpass",
             "Bandstop": "bandstop"}[kind]
    wn = bounds_hz[0] if len(bounds_hz) == 1 else list(bounds_hz)
    if spec[0] == "butterworth":
        return sps.butter(spec[1], wn, btype=btype, fs=fs, output="zpk")
    if spec[0] == "chebyshev1":
        return sps.cheby1(spec[1], spec[2], wn, btype=btype, fs=fs, output="zpk")
    raise ValueError(f"unknown prototype {spec!r}")


def zpk2sos_dspjl(z, p, k):
    """convert(SecondOrderSections, ZeroPoleGain) the way DSP.jl orders it
    (SURVEY.md App. B.2): poles sorted by distance to the unit circle, each pair
    grouped with its closest zeros, sections emitted in reverse, gain kept apart.
    Returns (coef[M,5] rows b0 b1 b2 a1 a2, g)."""
    z, p = np.asarray(z, dtype=complex), np.asarray(p, dtype=complex)

    def split(v):
        tol = 1e-13
        real = sorted(float(x.real) for x in v if abs(x.imag) <= tol * max(1.0, abs(x)))
        cpos = [x for x in v if x.imag > tol * max(1.0, abs(x))]
        return cpos, real

    cz, rz = split(z)
    cp, rp = split(p)
    cp.sort(key=lambda x: abs(abs(x) - 1))
    rp.sort(key=lambda x: abs(abs(x) - 1))
    secs = []
    for pole in cp:
        if cz:
            j = int(np.argmin([abs(q - pole) for q in cz]))
            q = cz.pop(j)
            zs = [q, np.conj(q)]
        else:
            zs = []
            for _ in range(2):
                if rz:
                    j = int(np.argmin([abs(q - pole) for q in rz]))
                    zs.append(rz.pop(j))
        secs.append((zs, [pole, np.conj(pole)]))
    while len(rp) >= 2:
        p0, p1 = rp.pop(0), rp.pop(0)
        if cz:
            j = int(np.argmin([abs(q - p0) for q in cz]))
            q = cz.pop(j)
            zs = [q, np.conj(q)]
        else:
            zs = []
            for t in (p0, p1):
                if rz:
                    j = int(np.argmin([abs(q - t) for q in rz]))
                    zs.append(rz.pop(j))
        secs.append((zs, [p0, p1]))
    first = None
    if rp:
        p0 = rp.pop(0)
        zs = []
        if rz:
            j = int(np.argmin([abs(q - p0) for q in rz]))
            zs.append(rz.pop(j))
        first = (zs, [p0])
    ordered = ([first] if first else []) + secs[::-1]
    rows = []
    for zs, ps in ordered:
        b = np.real(np.poly(zs)) if zs else np.array([1.0])
        a = np.real(np.poly(ps))
        b = np.concatenate([b, np.zeros(3 - len(b))])
        a = np.concatenate([a, np.zeros(3 - len(a))])
        rows.append([b[0], b[1], b[2], a[1], a[2]])
    return np.array(rows, dtype=np.float64).reshape(-1, 5), float(k)


def sos_filt(x, coef, g, state):
    """Stream one channel through the DF2T cascade; `state` (M,2) is updated."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    lib().oracle_sos_filt(_dp(out), _dp(x), len(x), _dp(coef), coef.shape[0], float(g), _dp(state))
    return out


# ---- resampling filters ---------------------------------------------------------------

class FirState(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_phi", C.c_int32), ("taps_per_phi", C.c_int64),
                ("h_len", C.c_int64), ("interpolation", C.c_int32), ("decimation", C.c_int32),
                ("phi_step", C.c_int32), ("phi_idx", C.c_int32), ("input_deficit", C.c_int64),
                ("x_idx", C.c_int64), ("rate", C.c_double), ("delta", C.c_double),
                ("phi_acc", C.c_double), ("alpha", C.c_double), ("pfb", C.POINTER(C.c_double)),
                ("dpfb", C.POINTER(C.c_double)), ("history", C.POINTER(C.c_double)), ("simd", C.c_int32)]


def resample_filter(rate):
    """DSP.resample_filter (App. B.3) with scipy's Kaiser-windowed sinc."""
    if isinstance(rate, Fraction):
        nphi = rate.numerator
        fnyq = min(1.0 / nphi, 1.0 / rate.denominator)
    else:
        nphi = 32
        fnyq = 1.0 / nphi if rate >= 1.0 else rate / nphi
    cutoff = fnyq
    tw = 0.2 * cutoff
    n = math.ceil((60 - 7.95) / (math.pi * 2.285 * tw)) + 1
    beta = 0.1102 * (60 - 8.7)
    hlen = nphi * math.ceil(n / nphi)
    if hlen % 2 == 0:
        hlen += 1
    h = sps.firwin(hlen, cutoff, window=("kaiser", beta), scale=True)
    return h * nphi, nphi


def taps2pfb(h, nphi):
    """DSP.taps2pfb: Julia-layout (column-major) matrix, rows filled bottom-up."""
    tapsper = -(-len(h) // nphi)
    pfb = np.zeros((tapsper, nphi), dtype=np.float64, order="F")
    idx = 0
    for row in range(tapsper - 1, -1, -1):
        for col in range(nphi):
            pfb[row, col] = h[idx] if idx < len(h) else 0.0
            idx += 1
    return pfb


class StandardFIR:
    """FIRFilter(h) with ratio 1 (DSP.jl FIRStandard, App. B.4): y[n] = sum_k h[k] x[n-k], history carried
    between calls, no phase shift."""

    def __init__(self, h):
        h = np.asarray(h, dtype=np.float64)
        st = FirState()
        st.kind = 0
        st.h_len = len(h)
        st.n_phi, st.taps_per_phi = 1, len(h)
        st.interpolation = st.decimation = 1
        st.input_deficit = 1
        st.phi_idx = 1
        st.x_idx = 1
        st.rate = 1.0
        self._pfb = np.asfortranarray(h[::-1].reshape(-1, 1))
        self._hist = np.zeros(max(1, len(h) - 1), dtype=np.float64)
        st.pfb = _dp(self._pfb)
        st.dpfb = _dp(self._pfb)
        st.history = _dp(self._hist)
        self.st = st

    rate = 1.0

    def outputlength(self, n):
        return max(0, n - self.st.input_deficit + 1)

    def filt(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        buf = np.empty(len(x) + 64, dtype=np.float64)
        n = lib().oracle_fir_filt(_dp(buf), len(buf), C.byref(self.st), _dp(x), len(x))
        return buf[:n]


class Resampler:
    """FIRFilter(resample_filter(ratio), ratio) + setphase!(timedelay) — the object
    `ResamplerFn(fs)` returns at src/reformatting.jl:92-99 — as streaming state."""

    def __init__(self, ratio):
        h, nphi = resample_filter(ratio)
        self.h = h
        st = FirState()
        st.h_len = len(h)
        st.input_deficit = 1
        st.phi_idx = 1
        st.alpha = 0.0
        st.x_idx = 1
        if isinstance(ratio, Fraction):
            p, q = ratio.numerator, ratio.denominator
            st.interpolation, st.decimation = p, q
            st.rate = p / q
            if p == 1:
                st.kind = 2                              # decimator: reversed taps
                self._pfb = np.asfortranarray(h[::-1].reshape(-1, 1))
                st.n_phi, st.taps_per_phi = 1, len(h)
                tau = (len(h) - 1) / 2
                st.input_deficit += int(round(tau))
            else:
                st.kind = 1 if q == 1 else 3
                self._pfb = taps2pfb(h, p)
                st.n_phi, st.taps_per_phi = p, self._pfb.shape[0]
                st.phi_step = q % p
                tau = (len(h) - 1) / (2.0 * p)
                frac, whole = math.modf(tau)
                st.input_deficit += int(round(whole))
                st.phi_idx = int(round(frac * p + 1.0))
            self._dpfb = self._pfb
        else:
            st.kind = 4
            st.rate = float(ratio)
            self._pfb = taps2pfb(h, nphi)
            self._dpfb = taps2pfb(np.append(np.diff(h), 0.0), nphi)
            st.n_phi, st.taps_per_phi = nphi, self._pfb.shape[0]
            st.delta = nphi / st.rate
            tau = (len(h) - 1) / (2.0 * nphi)
            frac, whole = math.modf(tau)
            st.input_deficit += int(round(whole))
            st.phi_acc = frac * nphi + 1.0
            st.phi_idx = int(math.floor(st.phi_acc))
            st.alpha = st.phi_acc - st.phi_idx
        self._hist = np.zeros(max(1, st.taps_per_phi - 1), dtype=np.float64)
        st.pfb = _dp(self._pfb)
        st.dpfb = _dp(self._dpfb)
        st.history = _dp(self._hist)
        self.st = st

    def outputlength(self, n_in):
        """DSP.outputlength (an upper estimate for the arbitrary kernel)."""
        st = self.st
        n = n_in - st.input_deficit + 1
        if st.kind == 4:
            return max(0, int(math.ceil(n * st.rate)))
        if st.kind == 2:
            return max(0, -(-n // st.decimation)) if n > 0 else 0
        return max(0, -(-(n * st.interpolation - (st.phi_idx - 1)) // st.decimation)) if n > 0 else 0

    def filt(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        cap = int(len(x) * max(self.st.rate, 1.0)) + 64
        buf = np.empty(cap, dtype=np.float64)
        n = lib().oracle_fir_filt(_dp(buf), cap, C.byref(self.st), _dp(x), len(x))
        if n < 0:
            raise RuntimeError("oracle FIR output buffer too small")
        return buf[:n]
