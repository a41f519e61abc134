"""Seeded differential test: random batches of filter / resample chains, GPU vs oracle.

Sizes are chosen to land on every kernel-selection branch of csrc/runtime.cu: tensor-map and
per-lane TMA IIR, WARM and MAIN+CARRY+FIX chunkings, Float32 and Float64, ragged row groups,
frame counts that are / are not multiples of 16, inputs shorter than the output, tensor-core
and scalar FIR.  Reference behaviour: src/filters.jl:204-262, src/reformatting.jl:92-122."""
import os

import numpy as np
import pytest

import oracle
from signalops import (Amplify, Bandpass, Bandstop, Filt, Highpass, Hz, Lowpass, Normpower, Pad, Signal, ToFramerate, Until,
                       dB, frames, kHz, sink_batch, zero)

pytestmark = pytest.mark.gpu


def rms(a):
    return float(np.sqrt(np.mean(np.asarray(a, dtype=np.float64) ** 2)))


def make_case(seed):
    r = np.random.default_rng(seed)
    f32 = bool(r.integers(0, 4) == 0)
    nch = int(r.choice([1, 2, 3]))
    ninst = int(r.choice([1, 3, 16, 32, 33, 64, 130]))
    n = int(r.choice([257, 1000, 4096, 16000, 16016, 44100, 48000]))
    fs = float(r.choice([8000.0, 44100.0, 48000.0]))
    kind = int(r.integers(0, 6))
    order = int(r.integers(1, 9))
    lo, hi = 0.02 * fs, 0.2 * fs
    gain = float(r.uniform(-30, 6))
    gq = (np.float32(gain) if f32 else gain) * dB
    if kind == 0:
        tail = lambda x: x >> Filt(Lowpass, hi * Hz, order=order) >> Amplify(gq)                          # noqa: E731
    elif kind == 1:
        tail = lambda x: x >> Filt(Highpass, lo * Hz, order=order) >> Normpower >> Amplify(gq)            # noqa: E731
    elif kind == 2:
        tail = lambda x: x >> Filt(Bandpass, lo * Hz, hi * Hz, order=min(order, 4))                       # noqa: E731
    elif kind == 3:
        tail = lambda x: x >> Filt(Bandstop, lo * Hz, hi * Hz, order=min(order, 4)) >> Amplify(gq)        # noqa: E731
    elif kind == 4:
        tail = lambda x: x >> Pad(zero) >> Until((n + 777) * frames) >> Filt(Lowpass, hi * Hz, order=order)   # noqa: E731
    else:
        fs_out = float(r.choice([fs * 1.5, fs / 2, 48000.0 if fs != 48000.0 else 44100.0]))
        tail = lambda x: ToFramerate(x, fs_out * Hz)                                                      # noqa: E731
    dt = np.float32 if f32 else np.float64
    xs = [r.standard_normal((n, nch)).astype(dt) for _ in range(ninst)]
    return xs, (lambda x: tail(Signal(x, fs * Hz))), f32


@pytest.mark.parametrize("seed", range(int(os.environ.get("SIGOPS_FUZZ_SEEDS", "60"))))
def test_random_batch_matches_oracle(gpu, seed):
    xs, chain, f32 = make_case(seed)
    # every other case runs the batch as ONE wave, so that the many-row kernels see it whole
    saved = os.environ.pop("SIGOPS_HOST_WAVES", None)
    try:
        if seed % 2:
            os.environ["SIGOPS_HOST_WAVES"] = "1"
        got = sink_batch([chain(x) for x in xs], gpu)
    finally:
        os.environ.pop("SIGOPS_HOST_WAVES", None)
        if saved is not None:
            os.environ["SIGOPS_HOST_WAVES"] = saved
    tol = 1e-5 if f32 else 1e-9
    for k in sorted({0, len(xs) // 2, len(xs) - 1}):
        want, fs = oracle.sink(chain(xs[k]))
        y, fs_got = got[k]
        assert fs_got == fs and y.shape == want.shape and y.dtype == want.dtype
        assert np.max(np.abs(y.astype(np.float64) - want.astype(np.float64))) <= tol * max(rms(want), 1e-300), (seed, k)


def make_map_case(seed):
    from signalops import AddChannel, After, Append, FadeTo, Mix, Ramp, ToChannels, cycle, mirror, ms, s, sin
    r = np.random.default_rng(1000 + seed)
    nch = int(r.choice([1, 2]))
    ninst = int(r.choice([2, 40, 150]))
    n = int(r.choice([3000, 20000, 44100]))
    fs = float(r.choice([8000.0, 44100.0]))
    dur = n / fs
    kind = int(r.integers(0, 5))
    w = float(r.uniform(50, 900))
    if kind == 0:      # appended pieces with different programs, joint Normpower, gain
        chain = lambda x: (Append(Signal(sin, ω=w * Hz) >> ToChannels(nch) >> Until(0.3 * s) >> Ramp(5 * ms), Signal(x, fs * Hz))   # noqa: E731
                           >> Normpower >> Amplify(-12 * dB))
    elif kind == 1:    # pads of every kind, cut in the padding
        pad = [zero, cycle, mirror][int(r.integers(0, 3))]
        chain = lambda x: Signal(x, fs * Hz) >> Pad(pad) >> Until((dur * 1.7) * s) >> Ramp(10 * ms) >> Normpower   # noqa: E731
    elif kind == 2:    # mix with a generated tone, skip the head
        chain = lambda x: Mix(Signal(x, fs * Hz), Signal(sin, ω=w * Hz, ϕ=0.25)) >> After(0.05 * s) >> Until((dur / 2) * s) >> Amplify(3 * dB)   # noqa: E731
    elif kind == 3:    # cross-fade two data signals
        chain = lambda x: FadeTo(Signal(x, fs * Hz), Signal(x[::-1].copy(), fs * Hz), 20 * ms)   # noqa: E731
    else:              # channel plumbing + normalisation
        chain = lambda x: Signal(x, fs * Hz) >> AddChannel(Signal(sin, ω=w * Hz) >> Until(dur * s)) >> Normpower >> Amplify(-20 * dB)   # noqa: E731
    xs = [r.standard_normal((n, nch)) for _ in range(ninst)]
    return xs, chain


@pytest.mark.parametrize("seed", range(int(os.environ.get("SIGOPS_FUZZ_MAP_SEEDS", "20"))))
def test_random_map_batch_matches_oracle(gpu, seed):
    xs, chain = make_map_case(seed)
    saved = os.environ.pop("SIGOPS_HOST_WAVES", None)
    try:
        os.environ["SIGOPS_HOST_WAVES"] = "1"      # one wave: large launches take the multi-pass map kernel
        got = sink_batch([chain(x) for x in xs], gpu)
    finally:
        os.environ.pop("SIGOPS_HOST_WAVES", None)
        if saved is not None:
            os.environ["SIGOPS_HOST_WAVES"] = saved
    for k in sorted({0, len(xs) - 1}):
        want, fs = oracle.sink(chain(xs[k]))
        y, fs_got = got[k]
        assert fs_got == fs and y.shape == want.shape
        assert np.max(np.abs(y - want)) <= 1e-9 * max(rms(want), 1e-300), (seed, k)
