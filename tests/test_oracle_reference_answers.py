"""The oracle against every exact / known-answer assertion the reference's own tests hold
for the sink path (test/runtests.jl, listed in SURVEY.md §4), plus independent scipy
cross-checks for the DSP.jl-backed parts (which the reference pins only relatively)."""
import numpy as np
import pytest
from scipy import signal as sps

import oracle
from signalops import (AddChannel, After, Amplify, Append, Bandpass, Bandstop, Chebyshev1, Extend, FadeTo, Filt,
                       Highpass, Lowpass, Mix, Normpower, OperateOn, Pad, Ramp, RampOff, RampOn, SelectChannel,
                       Signal, SignalError, ToChannels, ToFramerate, Until, Window, cycle, dB, deg, digitalfilter,
                       duration, frames, Hz, identity, inflen, kHz, lastframe, mirror, ms, nchannels, nframes,
                       one, rad, randn, s, sin, zero)
from signalops import cos as cos_


def arr(x):
    out = oracle.sink(x)
    return out[0] if isinstance(out, tuple) else out


def test_array_tuple_output():                                   # runtests.jl:66-70
    x = np.random.default_rng(0).random((10, 2))
    data, fs = oracle.sink(Mix(Signal(x, 10 * Hz), 1))
    assert fs == 10 and np.array_equal(data, x + 1)


def test_function_signals_phase_units():                         # runtests.jl:73-86
    base = arr(Signal(sin, ω=5 * Hz, ϕ=np.pi) >> Until(1 * s) >> ToFramerate(20 * Hz))
    assert np.array_equal(base, arr(Signal(sin, ω=5 * Hz, ϕ=np.pi * rad) >> Until(1 * s) >> ToFramerate(20 * Hz)))
    assert np.array_equal(base, arr(Signal(sin, ω=5 * Hz, ϕ=100 * ms) >> Until(1 * s) >> ToFramerate(20 * Hz)))
    assert np.array_equal(base, arr(Signal(sin, ω=5 * Hz, ϕ=180 * deg) >> Until(1 * s) >> ToFramerate(20 * Hz)))
    a = arr(Signal(sin, ϕ=1 * s) >> Until(1 * s) >> ToFramerate(20 * Hz))
    b = arr(Signal(sin, ω=1 * Hz, ϕ=0) >> Until(1 * s) >> ToFramerate(20 * Hz))
    assert np.allclose(a, b)
    with pytest.raises(SignalError):
        oracle.sink(Signal(sin, ϕ=2 * np.pi * rad) >> Until(1 * s) >> ToFramerate(20 * Hz))
    assert duration(oracle_signal(Signal(identity, 10 * Hz, ω=2 * Hz) >> Until(10 * frames))) == 1.0


def oracle_signal(x):
    data, fs = oracle.sink(x)
    return Signal(data, fs * Hz)


def test_first_sample_is_one_frame_in():                         # runtests.jl:90-91, SURVEY C-7
    tone = arr(Signal(sin, 44.1 * kHz, ω=100 * Hz) >> Until(5 * s))
    assert tone[0] < tone[109]
    assert tone[0, 0] == pytest.approx(np.sin(2 * np.pi * 100 / 44100), abs=1e-15)


@pytest.mark.parametrize("nch", [1, 2])
def test_cutting(nch):                                           # runtests.jl:118-170
    tone = Signal(sin, 44.1 * kHz, ω=100 * Hz) >> ToChannels(nch) >> Until(5 * s)
    assert nframes(tone) == 44100 * 5
    rng = np.random.default_rng(1)
    assert arr(After(rng.random((10, nch)), 0 * frames)).shape[0] == 10
    assert np.array_equal(arr(Until(np.arange(1, 11), 5 * frames)).ravel(), np.arange(1, 6))
    assert arr(Until(np.arange(1, 11), -5 * frames)).size == 0
    x = rng.random((12, nch))
    assert nframes(Signal(x, 6 * Hz) >> After(0.5 * s) >> Until(1 * s)) == 6
    assert nframes(Signal(x, 6 * Hz) >> Until(1 * s) >> After(0.5 * s)) == 3
    assert np.array_equal(arr(Signal(x, 6 * Hz) >> Until(1 * s) >> Until(0.5 * s)), arr(Signal(x, 6 * Hz) >> Until(0.5 * s)))
    with pytest.raises(SignalError):
        oracle.sink(Signal(np.arange(1, 11), 5 * Hz) >> After(3 * s))
    x = rng.random((20, nch))
    assert np.array_equal(arr(Window(x, from_=0 * frames, to=5 * frames)), x[:5])
    assert np.array_equal(arr(Window(x, from_=15 * frames, to=25 * frames)), x[15:20])
    xs = Signal(rng.random((12, nch)), 6 * Hz)
    assert nframes(Append(Until(xs, 1 * s), After(xs, 1 * s))) == 12
    assert nframes(tone >> After(2 * s)) == 44100 * 3


@pytest.mark.parametrize("nch", [1, 2, 3])
def test_padding(nch):                                           # runtests.jl:173-231
    tone = arr(Signal(sin, 22 * Hz, ω=10 * Hz) >> ToChannels(nch) >> Until(5 * s) >> Pad(zero) >> Until(7 * s))
    assert np.mean(np.abs(tone[:110])) > 0 and np.mean(np.abs(tone[110:154])) == 0
    rng = np.random.default_rng(2)
    assert np.array_equal(arr(Signal(rng.random((10, nch)), 10 * Hz) >> Pad(zero) >> After(15 * frames) >> Until(10 * frames)),
                          np.zeros((10, nch)))
    x = rng.random((10, nch))
    assert np.array_equal(arr(Pad(Signal(x, 10 * Hz), cycle) >> Until(30 * frames)), np.vstack([x, x, x]))
    assert np.array_equal(arr(Pad(Signal(x, 10 * Hz), mirror) >> Until(30 * frames)), np.vstack([x, x[::-1], x]))
    r = arr(Pad(Signal(x, 10 * Hz), lastframe) >> Until(15 * frames))
    assert np.all(r[10:] == r[9:10])
    gen = Signal(sin, 10 * Hz) >> ToChannels(nch) >> Until(1 * s)
    with pytest.raises(SignalError):
        oracle.sink(Pad(gen, cycle) >> Until(15 * frames))
    r = arr(Pad(gen, lastframe) >> Until(15 * frames))
    assert np.all(r[10:] == r[9:10])
    padv = rng.random(nch)
    r = arr(Pad(gen, padv) >> Until(15 * frames))
    assert np.all(r[10:] == padv)
    y = rng.random((15, nch))
    assert nframes(Extend(x, one)) is inflen
    assert nframes(Mix(Extend(x, one), y)) == 15 and nframes(Mix(y, Extend(x, one))) == 15
    assert nframes(Mix(Pad(x, one), y)) is inflen
    assert nframes(Mix(1, rng.random((10, 2)))) == 10 and nframes(Mix(rng.random((10, 2)), 1)) == 10
    assert nframes(Mix(sin, 1, rng.random((10, 2)))) is inflen


@pytest.mark.parametrize("nch", [1, 2])
def test_appending(nch):                                         # runtests.jl:234-257
    a = Signal(sin, 22 * Hz, ω=10 * Hz) >> ToChannels(nch) >> Until(5 * s)
    b = Signal(sin, 22 * Hz, ω=5 * Hz) >> ToChannels(nch) >> Until(5 * s)
    tones = a >> Append(b)
    assert duration(tones) == 10
    assert np.array_equal(arr(tones), np.vstack([arr(a), arr(b)])) and arr(tones).shape[0] == 220
    q = Signal(2, 3) >> ToChannels(nch) >> Until(2 * s) >> Append(Signal(3, 3)) >> Until(4 * s)
    assert arr(q).shape[0] == 12
    with pytest.raises(SignalError):
        Append(sin, np.arange(10))


def test_mixing_and_padded_maps():                               # runtests.jl:260-311, 875-878
    for nch in (1, 2):
        a = Signal(2, 3 * Hz) >> ToChannels(nch) >> Until(2 * s) >> Append(Signal(3, 3 * Hz)) >> Until(4 * s)
        b = Signal(3, 3 * Hz) >> ToChannels(nch) >> Until(3 * s)
        assert np.array_equal(arr(Mix(a, b))[:, 0], [5] * 6 + [6] * 3 + [3] * 3)
        assert np.array_equal(arr(Amplify(a, b))[:, 0], [6] * 6 + [9] * 3 + [3] * 3)
    rng = np.random.default_rng(3)
    x, y = rng.random((20, 65)), rng.random((20, 65))
    assert np.array_equal(arr(Mix(x, y) >> ToFramerate(20 * Hz)), x + y)
    x = rng.random((20, 2))
    r = arr(OperateOn(lambda v: (v[1], v[0]), x, bychannel=False) >> ToFramerate(20 * Hz))
    assert np.array_equal(r, x[:, ::-1])
    x, y, z = rng.random((10, 2)), rng.random((5, 2)), np.ones((10, 4))
    oracle.sink_into(z, Signal(x, 10 * Hz) >> AddChannel(y))
    assert np.all(z[5:, 2:] == 0) and np.array_equal(z[:5, 2:], y)
    r = arr(Mix(Append(Until(1, 1 * s), Until(2, 2 * s)), Append(Until(3, 2 * s), Until(4, 1 * s))) >> ToFramerate(10 * Hz))
    assert np.array_equal(r[:, 0], [4] * 10 + [5] * 10 + [6] * 10)


def test_numbers_and_db_are_exact():                             # runtests.jl:503-526
    for nch in (1, 2):
        assert np.all(arr(Signal(1, 10 * Hz) >> ToChannels(nch) >> Until(1 * s) >> Amplify(20 * dB)) == 10)
        assert np.all(arr(Signal(1, 10 * Hz) >> ToChannels(nch) >> Until(1 * s) >> Amplify(40 * dB)) == 100)
        x = arr(Signal(1, 5 * Hz) >> ToChannels(nch) >> Until(5 * s))
        assert x.dtype.kind == "i"
        tone = arr(Signal(sin, 200 * Hz, ω=10 * Hz) >> ToChannels(nch) >> Mix(1.5) >> Until(5 * s))
        assert np.all(tone >= 0.5)


@pytest.mark.parametrize("nch", [1, 2])
def test_filtering(nch):                                         # runtests.jl:314-371
    a = Signal(sin, 100 * Hz, ω=10 * Hz) >> ToChannels(nch) >> Until(5 * s)
    b = Signal(sin, 100 * Hz, ω=5 * Hz) >> ToChannels(nch) >> Until(5 * s)
    cm = Mix(a, b)
    high = arr(cm >> Filt(Highpass, 8 * Hz, method=Chebyshev1(5, 1)))
    low = arr(cm >> Filt(Lowpass, 6 * Hz))
    highlow = arr(Signal(low, 100 * Hz) >> Filt(Highpass, 8 * Hz, method=Chebyshev1(5, 1)))
    bp1 = arr(cm >> Filt(Bandpass, 20 * Hz, 30 * Hz, method=Chebyshev1(5, 1)))
    bp2 = arr(cm >> Filt(Bandpass, 2 * Hz, 12 * Hz, method=Chebyshev1(5, 1)))
    bs1 = arr(cm >> Filt(Bandstop, 20 * Hz, 30 * Hz, method=Chebyshev1(5, 1)))
    bs2 = arr(cm >> Filt(Bandstop, 2 * Hz, 12 * Hz, method=Chebyshev1(5, 1)))
    for bad in (lambda: Filt(a, Highpass, 75 * Hz), lambda: Filt(a, Lowpass, 75 * Hz),
                lambda: Filt(a, Bandpass, 75 * Hz, 80 * Hz), lambda: Filt(a, Bandstop, 75 * Hz, 80 * Hz)):
        with pytest.raises(SignalError):
            bad()
    assert high.shape[0] == low.shape[0] == highlow.shape[0] == 500
    assert np.mean(high) < 0.01 and np.mean(low) < 0.02
    assert 10 * np.mean(np.abs(highlow)) < np.mean(np.abs(low))
    assert 10 * np.mean(np.abs(highlow)) < np.mean(np.abs(high))
    assert 10 * np.mean(np.abs(bp1)) < np.mean(np.abs(bp2))
    assert 10 * np.mean(np.abs(bs2)) < np.mean(np.abs(bs1))
    # blocksize invariance and state across After (runtests.jl:353-362, 797-806)
    high2 = arr(cm >> Filt(Highpass, 8 * Hz, method=Chebyshev1(5, 1), blocksize=100))
    assert np.array_equal(high2, high)
    high3 = arr(cm >> Filt(Highpass, 8 * Hz, method=Chebyshev1(5, 1), blocksize=64) >> After(1 * s))
    assert np.array_equal(high3, high[100:])
    # custom filter object == designed filter (runtests.jl:364-368), to rounding of the two designs
    h = digitalfilter(Highpass(8, fs=100), Chebyshev1(5, 1))
    high4 = arr(cm >> Filt(h))
    assert np.max(np.abs(high4 - high)) < 1e-12 * np.sqrt(np.mean(high ** 2)) + 1e-15
    # independent check: scipy's own design and recurrence
    sos = sps.cheby1(5, 1, 8, btype="highpass", fs=100, output="sos")
    ref = sps.sosfilt(sos, arr(cm)[:, 0])
    assert np.max(np.abs(high[:, 0] - ref)) < 1e-10 * np.sqrt(np.mean(ref ** 2))


def test_short_block_operators():                                # runtests.jl:793-812
    x, y, z = Signal(np.ones((25, 2)), 10 * Hz), Signal(np.ones((10, 2)), 10 * Hz), Signal(np.ones((15, 2)), 10 * Hz)
    for mk in (lambda bs: x >> Append(y) >> Append(z) >> Filt(Lowpass, 3 * Hz, blocksize=bs),
               lambda bs: x >> RampOn(7 * frames) >> Filt(Lowpass, 3 * Hz, blocksize=bs),
               lambda bs: x >> Ramp(3 * frames) >> Filt(Lowpass, 3 * Hz, blocksize=bs)):
        assert np.array_equal(arr(mk(5)), arr(mk(4096)))
    assert arr(ToFramerate(y, 40 * Hz)).shape == (40, 2)
    assert arr(ToFramerate(y, 5 * Hz)).shape == (5, 2)
    with pytest.raises(SignalError):
        oracle.sink(ToFramerate(y, 40 * Hz, blocksize=5))


@pytest.mark.parametrize("nch", [1, 2])
def test_ramps(nch):                                             # runtests.jl:374-403
    tone = Signal(sin, 50 * Hz, ω=10 * Hz) >> ToChannels(nch) >> Until(5 * s)
    ramped = arr(tone >> Ramp(500 * ms))
    sq = ramped ** 2
    assert sq[:25].mean() < sq[25:50].mean() and sq[225:].mean() < sq[200:225].mean()
    x = Signal(sin, 22 * Hz, ω=10 * Hz) >> ToChannels(nch) >> Until(2 * s)
    y = Signal(sin, 22 * Hz, ω=5 * Hz) >> ToChannels(nch) >> Until(2 * s)
    fading = FadeTo(x, y, 500 * ms)
    r = arr(fading)
    assert nframes(fading) == int(np.ceil((2 + 2 - 0.5) * 22)) == r.shape[0]
    assert np.array_equal(r[:33], arr(x)[:33]) and np.array_equal(r[43:], arr(y)[10:])
    r2 = arr(Signal(sin, 500 * Hz, ω=20 * Hz, ϕ=np.pi / 2) >> ToChannels(nch) >> Until(100 * ms) >> Ramp(identity))
    assert np.mean(np.abs(r2[:5])) < np.mean(np.abs(r2[5:10]))
    # envelope formulas: starts at fn(0)=0, last frame is fn(0)=0
    ones_ = arr(Signal(np.ones(100), 1 * kHz) >> Ramp(10 * ms))
    assert ones_[0, 0] == 0 and ones_[-1, 0] == 0 and ones_[50, 0] == 1
    assert ones_[1, 0] == pytest.approx(np.sin(np.pi * 0.5 * 0.1))


@pytest.mark.parametrize("nch", [1, 2])
def test_resampling(nch):                                        # runtests.jl:406-458
    tone = Signal(sin, 20 * Hz, ω=5 * Hz) >> ToChannels(nch) >> Until(5 * s)
    assert nframes(ToFramerate(tone, 40 * Hz)) == 200
    down = ToFramerate(tone, 15 * Hz)
    assert nframes(down) == 75 and arr(down).shape[0] == 75
    toned = oracle_signal(tone)
    r1 = arr(ToFramerate(toned, 40 * Hz))
    r2 = arr(ToFramerate(toned, 40 * Hz, blocksize=64))
    assert r1.shape[0] == 200 and np.allclose(r1, r2)
    assert np.array_equal(r2, arr(ToFramerate(toned, 40 * Hz, blocksize=64)))    # state reset
    padded = tone >> Pad(one) >> Until(7 * s)
    assert arr(ToFramerate(padded, 40 * Hz)).shape == (280, nch)
    assert ToFramerate(tone, 20 * Hz) is tone
    twice = ToFramerate(ToFramerate(toned, 15 * Hz), 50 * Hz)
    assert twice.signal is toned                                  # resampler of resampler collapses
    a = Signal(sin, 48 * Hz, ω=10 * Hz) >> ToChannels(nch) >> Until(3 * s)
    high = Mix(a, a) >> Filt(Highpass, 8 * Hz, method=Chebyshev1(5, 1))
    assert arr(ToFramerate(high, 24 * Hz)).shape == (72, nch)
    # exact-rational kernels against scipy's polyphase filter with the same taps
    from oracle import dspjl_ref as D
    from fractions import Fraction
    x = np.random.default_rng(5).standard_normal(400)
    for p, q in ((2, 1), (3, 2), (1, 2), (2, 3)):
        got = arr(ToFramerate(Signal(x, 1000 * Hz), 1000 * p / q * Hz))[:, 0]
        R = D.Resampler(Fraction(p, q))
        # zero-stuffed, filtered, full-rate signal; the kernel's first output sits at the
        # group-delay-compensated position u0 and then takes every q-th sample
        full = sps.upfirdn(R.h, x, up=p, down=1)
        u0 = (R.st.input_deficit - 1) * p + (R.st.phi_idx - 1)
        assert u0 == (len(R.h) - 1) // 2
        n = min(len(got), (len(full) - u0 + q - 1) // q)
        assert n > 100
        assert np.max(np.abs(got[:n] - full[u0:u0 + n * q:q])) < 1e-12, (p, q)


def test_normpower():                                            # runtests.jl:490-500
    for nch in (1, 2):
        tone = Signal(sin, 10 * Hz, ω=2 * Hz) >> ToChannels(nch) >> Until(2 * s) >> Ramp() >> Normpower
        assert np.allclose(np.sqrt(np.mean(arr(tone) ** 2, axis=0)), 1)
        assert np.allclose(np.sqrt(np.mean(arr(tone >> ToFramerate(20 * Hz)) ** 2, axis=0)), 1)
    with pytest.raises(SignalError):
        oracle.sink(Signal(sin, 200 * Hz) >> Normpower >> Until(1 * s))


def test_stress_combinations():                                  # runtests.jl:815-889
    a, b = Until(sin, 2 * s), Until(cos_, 2 * s)
    x = Append(a, b) >> After(3 * s)
    assert np.array_equal(arr(x >> ToFramerate(20 * Hz)), arr(b >> After(1 * s) >> ToFramerate(20 * Hz)))
    noise = oracle_signal(Signal(randn, 20 * Hz, rng=np.random.default_rng(6)) >> Until(6 * s))
    x4 = arr(noise >> Filt(Lowpass, 7 * Hz) >> Until(4 * s))
    after = arr(noise >> Filt(Lowpass, 7 * Hz) >> Until(4 * s) >> After(2 * s))
    assert np.allclose(x4[40:], after)
    x = (oracle_signal(Signal(sin, 20 * Hz, ω=10 * Hz) >> Until(4 * s)) >> ToFramerate(30 * Hz) >> Filt(Lowpass, 10 * Hz)
         >> FadeTo(Signal(sin, ω=5 * Hz) >> Until(4 * s), 500 * ms) >> ToFramerate(22 * Hz))
    assert x.framerate == 22 and duration(x) == 7.5 and arr(x).shape[0] == 165
    xa = arr(noise >> Filt(Lowpass, 9 * Hz) >> Mix(Signal(sin, ω=12 * Hz) >> Until(6 * s))
             >> Filt(Highpass, 4 * Hz, method=Chebyshev1(5, 1)))
    ya = arr(noise >> Filt(Lowpass, 9 * Hz, blocksize=11) >> Mix(Signal(sin, ω=12 * Hz) >> Until(6 * s))
             >> Filt(Highpass, 4 * Hz, method=Chebyshev1(5, 1), blocksize=9))
    assert np.allclose(xa, ya)
    x = (Signal(sin, ω=5 * Hz) >> After(2 * s) >> Until(20 * s) >> After(2 * s) >> Until(15 * s) >> After(2 * s)
         >> After(2 * s) >> Until(5 * s) >> Until(2 * s) >> ToFramerate(12 * Hz))
    assert duration(oracle_signal(x)) == 2
    x = (Signal(randn, rng=np.random.default_rng(7)) >> Until(4 * s) >> After(50 * ms) >> Filt(Lowpass, 5 * Hz)
         >> Mix(Signal(sin, ω=7 * Hz)) >> Until(3.5 * s) >> Filt(Highpass, 2 * Hz)
         >> Append(np.random.default_rng(8).random((10, 2))) >> Append(np.random.default_rng(9).random((5, 2)))
         >> ToFramerate(20 * Hz))
    assert duration(oracle_signal(x)) == 4.25


def test_float32_is_preserved():                                 # runtests.jl:707-729
    x = Signal(np.random.default_rng(10).random((100, 2)).astype(np.float32), 10 * Hz)
    y = Signal(np.random.default_rng(11).random((50, 2)).astype(np.float32), 10 * Hz)
    for g in (x >> Until(5 * s), x >> Append(y) >> After(2 * s), x >> Pad(zero) >> Until(15 * s),
              x >> Filt(Lowpass, 3 * Hz), x >> Normpower >> Amplify(np.float32(-10) * dB), x >> Mix(y),
              x >> AddChannel(y), x >> SelectChannel(1), x >> Ramp(), x >> FadeTo(y)):
        assert arr(g).dtype == np.float32


def test_unknown_frame_rates():                                  # runtests.jl:758-790
    x, y = np.random.default_rng(12).random((100, 2)), np.random.default_rng(13).random((50, 2))
    assert Signal(x).framerate is None and nchannels(x) == 2 and nframes(x) == 100
    assert oracle.sink(ToFramerate(x, 10 * Hz))[1] == 10
    for g, n in ((Until(x, 3 * s), 30), (After(x, 3 * s), 70), (Append(x, y), 150),
                 (Append(x, y) >> After(2 * s), 130), (Append(x, y) >> Until(13 * s), 130),
                 (Pad(x, zero) >> Until(15 * s), 150), (Filt(x, Lowpass, 3 * Hz), 100),
                 (Normpower(x) >> Amplify(-10 * dB), 100), (Mix(x, y), 100), (AddChannel(x, y), 100),
                 (SelectChannel(x, 1), 100), (Ramp(x), 100)):
        assert arr(g >> ToFramerate(10 * Hz)).shape[0] == n
    with pytest.raises(SignalError):
        oracle.sink(FadeTo(x, y) >> ToFramerate(10 * Hz))


def test_change_channel_count():                                  # runtests.jl:103-114
    tone = Signal(sin, 22 * Hz, ω=10 * Hz) >> Until(5 * s)
    assert nchannels(tone >> ToChannels(2)) == 2 and nchannels(tone >> ToChannels(1)) == 1
    data = arr(tone >> ToChannels(2))
    assert data.shape[1] == 2
    data2 = arr(Signal(data, 22 * Hz) >> ToChannels(1))
    assert data2.shape[1] == 1 and np.array_equal(data2, data.sum(axis=1, keepdims=True))     # sum, not mean
    with pytest.raises(SignalError):
        tone >> ToChannels(2) >> ToChannels(3)


def test_automatic_reformatting():                                # runtests.jl:461-470
    a = Signal(sin, 200 * Hz, ω=10 * Hz) >> ToChannels(2) >> Until(5 * s)
    b = Signal(sin, 100 * Hz, ω=5 * Hz) >> Until(3 * s)
    mixed = Mix(a, b)
    assert nchannels(mixed) == 2 and mixed.framerate == 200
    assert arr(mixed).shape[0] == 1000
    assert arr(Mix(a, b, 1)).shape[0] == 1000


@pytest.mark.parametrize("nch", [1, 2])
def test_empty_and_infinite_signals(nch):                         # runtests.jl:481-488, 550-575
    tone = Signal(sin, 200 * Hz, ω=10 * Hz) >> ToChannels(nch) >> Until(10 * frames) >> Until(0 * frames)
    assert nframes(tone) == 0 and arr(tone).shape == (0, nch)
    assert nframes(OperateOn(np.negative, tone)) == 0
    tone = Signal(sin, 200 * Hz, ω=10 * Hz) >> ToChannels(nch) >> Until(10 * frames) >> After(5 * frames) >> After(2 * frames)
    assert nframes(tone) == 3 and arr(tone).shape == (3, nch)
    for mk in (lambda: Signal(sin, 200 * Hz, ω=10 * Hz) >> ToChannels(nch) >> After(5 * frames) >> Until(5 * frames),
               lambda: Signal(sin, 200 * Hz, ω=10 * Hz) >> ToChannels(nch) >> Until(10 * frames) >> After(5 * frames)):
        tone = mk()
        assert nframes(tone) == 5
        got = arr(tone)
        assert got.shape == (5, nch) and got[0, 0] > 0.9           # frame 6 of a 10 Hz tone at 200 Hz
    with pytest.raises(SignalError):
        oracle.sink(Signal(sin, 200 * Hz) >> Normpower >> Until(1 * s))
    with pytest.raises(SignalError):
        oracle.sink(Signal(sin, 200 * Hz) >> ToChannels(nch))


def test_frame_units():                                           # runtests.jl:602-614
    rng = np.random.default_rng(3)
    x = Signal(rng.random((100, 2)), 10 * Hz)
    y = Signal(rng.random((50, 2)), 10 * Hz)
    assert arr(x >> Until(30 * frames)).shape[0] == 30
    assert arr(x >> After(30 * frames)).shape[0] == 70
    assert arr(x >> Append(y) >> After(20 * frames)).shape[0] == 130
    assert arr(x >> Append(y) >> Until(130 * frames)).shape[0] == 130
    assert arr(x >> Pad(zero) >> Until(150 * frames)).shape[0] == 150
    assert arr(x >> Ramp(10 * frames)).shape[0] == 100
    assert arr(x >> FadeTo(y, 10 * frames)).shape[0] > 100


@pytest.mark.parametrize("nch", [1, 2])
def test_arrays_as_signals(nch):                                  # runtests.jl:527-535
    ramp10 = 10.0 * np.arange(1, 11)
    tone = arr(Signal(sin, 200 * Hz, ω=10 * Hz) >> ToChannels(nch) >> Until(10 * frames) >> Mix(ramp10))
    assert np.all(tone[:10, :] >= ramp10[:, None] - 1.0) and tone.shape == (10, nch)
    x = arr(Signal(ramp10, 5 * Hz) >> ToChannels(nch) >> Until(1 * s))
    assert x.dtype == np.float64 and x.shape == (5, nch)
    with pytest.raises(SignalError):
        Signal(np.zeros((2, 2, 2)))                                # poorly shaped arrays, runtests.jl:546


def test_readme_sound1_at_4khz():                                 # runtests.jl:896-900
    sound1 = Signal(sin, ω=1 * kHz) >> Until(5 * s) >> Ramp() >> Normpower >> Amplify(-20 * dB)
    data, fs = oracle.sink(sound1 >> ToFramerate(4 * kHz))
    assert fs == 4000 and data.shape[0] == 4000 * 5 and np.mean(np.abs(data)) > 0
    # Normpower then -20 dB: rms 0.1 (the ramp is inside the normalisation)
    assert np.sqrt(np.mean(data ** 2)) == pytest.approx(0.1, rel=1e-12)
