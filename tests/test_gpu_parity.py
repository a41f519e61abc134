"""GPU parity: sink(chain, GPUSink()) against the CPU oracle on the same inputs.

Tolerances are BASELINE.json's: lengths / frame counts bit-exact; Float64 samples
within 1e-9 of the signal RMS; Float32 within 1e-5; integer-valued chains exact.
Every call goes through the C ABI (ctypes -> libsignalops_cuda.so).
"""
import numpy as np
import pytest

import oracle
from signalops import (AddChannel, AffineSin, After, Amplify, Append, Bandpass, Bandstop, Chebyshev1,
                       Extend, FadeTo, Filt, Highpass, Lowpass, Mix, Normpower, Operate, OperateOn, Pad,
                       Ramp, RampOff, RampOn, Sawtooth, SelectChannel, Signal, ToChannels, ToEltype,
                       ToFramerate, Until, Window, cos, cycle, dB, frames, Hz, identity, kHz, lastframe,
                       mirror, ms, one, randn, s, sin, sink, sink_batch, sink_into, zero)

pytestmark = pytest.mark.gpu

F64_TOL = 1e-9
F32_TOL = 1e-5


def rms(a):
    a = np.asarray(a, dtype=np.float64)
    return float(np.sqrt(np.mean(a ** 2))) if a.size else 0.0


def check(gpu, make, tol=F64_TOL, exact=False):
    """`make()` builds a fresh graph (fresh RNG state) each time it is called."""
    want = oracle.sink(make())
    got = sink(make(), gpu)
    if isinstance(want, tuple):
        assert isinstance(got, tuple) and got[1] == want[1]
        want, got = want[0], got[0]
    assert got.shape == want.shape
    assert got.dtype == want.dtype
    if exact:
        assert np.array_equal(got, want)
    elif want.size:
        scale = max(rms(want), 1e-300)
        err = float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64)))) / scale
        assert err <= tol, f"max err / rms = {err:g}"
    return got


def rng(seed=1983):
    return np.random.default_rng(seed)


# ---- K1: generators, cuts, pads, appends, maps, ramps -------------------------------------

def test_tone(gpu):
    check(gpu, lambda: Signal(sin, 44.1 * kHz, ω=100 * Hz) >> Until(5 * s), tol=1e-10)


def test_tone_phase_and_no_omega(gpu):
    check(gpu, lambda: Signal(sin, ω=5 * Hz, ϕ=np.pi) >> Until(1 * s) >> ToFramerate(20 * Hz), tol=1e-10)
    check(gpu, lambda: Signal(sin, ϕ=1 * s) >> Until(1 * s) >> ToFramerate(20 * Hz), tol=1e-10)
    check(gpu, lambda: Signal(cos, 100 * Hz, ω=7 * Hz) >> Until(3 * s), tol=1e-10)
    check(gpu, lambda: Signal(Sawtooth(), 8 * kHz, ω=1 * kHz) >> Until(2 * s), tol=1e-10)
    check(gpu, lambda: Signal(AffineSin(0.5, 0.5), 8 * kHz, ω=5 * Hz) >> Until(2 * s), tol=1e-10)


def test_arbitrary_callable_is_host_materialised(gpu):
    check(gpu, lambda: Signal(lambda t: t * t - 1.0, 50 * Hz) >> Until(2 * s) >> Amplify(2.0), tol=1e-14)


def test_numbers_exact(gpu):
    check(gpu, lambda: Signal(1, 10 * Hz) >> ToChannels(2) >> Until(1 * s) >> Amplify(20 * dB), exact=True)
    check(gpu, lambda: Signal(1, 5 * Hz) >> ToChannels(2) >> Until(5 * s), exact=True)
    x = rng().random((10, 2))
    got = check(gpu, lambda: Mix(Signal(x, 10 * Hz), 1), exact=True)
    assert np.array_equal(got, x + 1)


@pytest.mark.parametrize("nch", [1, 2])
def test_cutting(gpu, nch):
    x = rng().random((12, nch))
    check(gpu, lambda: Signal(x, 6 * Hz) >> After(0.5 * s) >> Until(1 * s), exact=True)
    check(gpu, lambda: Signal(x, 6 * Hz) >> Until(1 * s) >> After(0.5 * s), exact=True)
    x20 = rng(2).random((20, nch))
    got = check(gpu, lambda: Window(x20, from_=15 * frames, to=25 * frames), exact=True)
    assert np.array_equal(got, x20[15:20])
    check(gpu, lambda: Signal(sin, 200 * Hz, ω=10 * Hz) >> ToChannels(nch) >> Until(10 * frames)
          >> After(5 * frames) >> After(2 * frames), tol=1e-10)
    check(gpu, lambda: Signal(sin, 200 * Hz, ω=10 * Hz) >> ToChannels(nch) >> Until(10 * frames)
          >> Until(0 * frames), exact=True)


@pytest.mark.parametrize("nch", [1, 3])
def test_padding(gpu, nch):
    x = rng().random((10, nch))
    check(gpu, lambda: Signal(x, 10 * Hz) >> Pad(zero) >> After(15 * frames) >> Until(10 * frames), exact=True)
    got = check(gpu, lambda: Pad(Signal(x, 10 * Hz), cycle) >> Until(30 * frames), exact=True)
    assert np.array_equal(got, np.vstack([x, x, x]))
    got = check(gpu, lambda: Pad(Signal(x, 10 * Hz), mirror) >> Until(30 * frames), exact=True)
    assert np.array_equal(got, np.vstack([x, x[::-1], x]))
    check(gpu, lambda: Pad(Signal(x, 10 * Hz), lastframe) >> Until(15 * frames), exact=True)
    padv = rng(3).random(nch)
    check(gpu, lambda: Pad(Signal(sin, 10 * Hz) >> ToChannels(nch) >> Until(1 * s), padv)
          >> Until(15 * frames), tol=1e-10)
    check(gpu, lambda: Pad(Signal(sin, 10 * Hz) >> ToChannels(nch) >> Until(1 * s), lastframe)
          >> Until(15 * frames), tol=1e-10)
    check(gpu, lambda: Signal(sin, 22 * Hz, ω=10 * Hz) >> ToChannels(nch) >> Until(5 * s) >> Pad(one)
          >> Until(7 * s), tol=1e-10)


@pytest.mark.parametrize("nch", [1, 2])
def test_appending_and_padded_maps(gpu, nch):
    check(gpu, lambda: (Signal(sin, 22 * Hz, ω=10 * Hz) >> ToChannels(nch) >> Until(5 * s))
          >> Append(Signal(sin, 22 * Hz, ω=5 * Hz) >> ToChannels(nch) >> Until(5 * s)), tol=1e-10)

    def ab():
        a = Signal(2, 3 * Hz) >> ToChannels(nch) >> Until(2 * s) >> Append(Signal(3, 3 * Hz)) >> Until(4 * s)
        b = Signal(3, 3 * Hz) >> ToChannels(nch) >> Until(3 * s)
        return a, b
    got = check(gpu, lambda: Mix(*ab()), exact=True)
    assert np.array_equal(got[:, 0], [5] * 6 + [6] * 3 + [3] * 3)
    got = check(gpu, lambda: Amplify(*ab()), exact=True)
    assert np.array_equal(got[:, 0], [6] * 6 + [9] * 3 + [3] * 3)
    got = check(gpu, lambda: Mix(Append(Until(1, 1 * s), Until(2, 2 * s)),
                                 Append(Until(3, 2 * s), Until(4, 1 * s))) >> ToFramerate(10 * Hz), exact=True)
    assert np.array_equal(got[:, 0], [4] * 10 + [5] * 10 + [6] * 10)


def test_channel_ops(gpu):
    x, y = rng().random((10, 2)), rng(5).random((5, 2))
    z = np.ones((10, 4))
    zo = np.ones((10, 4))
    sink_into(z, Signal(x, 10 * Hz) >> AddChannel(y), gpu)
    oracle.sink_into(zo, Signal(x, 10 * Hz) >> AddChannel(y))
    assert np.array_equal(z, zo) and np.all(z[5:, 2:] == 0)
    check(gpu, lambda: Signal(x, 10 * Hz) >> SelectChannel(2), exact=True)
    check(gpu, lambda: Signal(x, 10 * Hz) >> ToChannels(1), tol=1e-15)
    big = rng(6).random((20, 65))
    big2 = rng(7).random((20, 65))
    got = check(gpu, lambda: Mix(big, big2) >> ToFramerate(20 * Hz), exact=True)
    assert np.array_equal(got, big + big2)
    check(gpu, lambda: Signal(x, 10 * Hz) >> Operate(lambda *a: None) if False else
          OperateOn("-", Signal(x, 10 * Hz), y), exact=True)


@pytest.mark.parametrize("nch", [1, 2])
def test_ramps(gpu, nch):
    check(gpu, lambda: Signal(sin, 50 * Hz, ω=10 * Hz) >> ToChannels(nch) >> Until(5 * s) >> Ramp(500 * ms), tol=1e-10)
    check(gpu, lambda: Signal(sin, 500 * Hz, ω=20 * Hz, ϕ=np.pi / 2) >> ToChannels(nch) >> Until(100 * ms)
          >> Ramp(identity), tol=1e-10)
    check(gpu, lambda: Signal(sin, 500 * Hz, ω=20 * Hz) >> ToChannels(nch) >> Until(100 * ms)
          >> RampOn(20 * ms, lambda v: v * v), tol=1e-10)
    check(gpu, lambda: Signal(sin, 500 * Hz, ω=20 * Hz) >> ToChannels(nch) >> Until(100 * ms)
          >> RampOff(20 * ms, lambda v: v ** 3), tol=1e-10)

    def fade():
        a = Signal(sin, 22 * Hz, ω=10 * Hz) >> ToChannels(nch) >> Until(2 * s)
        b = Signal(sin, 22 * Hz, ω=5 * Hz) >> ToChannels(nch) >> Until(2 * s)
        return FadeTo(a, b, 500 * ms)
    got = check(gpu, fade, tol=1e-10)
    assert got.shape[0] == int(np.ceil((2 + 2 - 0.5) * 22))


# ---- K2: Normpower ---------------------------------------------------------------------------

@pytest.mark.parametrize("nch", [1, 2])
def test_normpower(gpu, nch):
    got = check(gpu, lambda: Signal(sin, 10 * Hz, ω=2 * Hz) >> ToChannels(nch) >> Until(2 * s) >> Ramp() >> Normpower,
                tol=1e-10)
    assert abs(rms(got) - 1) < 1e-12
    x = rng().standard_normal((5000, nch))
    got = check(gpu, lambda: Signal(x, 1 * kHz) >> Normpower >> Amplify(-10 * dB), tol=1e-10)
    check(gpu, lambda: Mix(Signal(x, 1 * kHz) >> Normpower, Signal(sin, 1 * kHz, ω=50 * Hz) >> Until(2 * s)
                           >> Normpower >> Amplify(-6 * dB)), tol=1e-10)
    check(gpu, lambda: Signal(x, 1 * kHz) >> Normpower >> After(1 * s) >> Until(2 * s), tol=1e-10)


# ---- K3: IIR ------------------------------------------------------------------------------------

@pytest.mark.parametrize("nch", [1, 2])
def test_filters_small(gpu, nch):
    def cmplx():
        a = Signal(sin, 100 * Hz, ω=10 * Hz) >> ToChannels(nch) >> Until(5 * s)
        b = Signal(sin, 100 * Hz, ω=5 * Hz) >> ToChannels(nch) >> Until(5 * s)
        return Mix(a, b)
    for mk in (lambda: cmplx() >> Filt(Highpass, 8 * Hz, method=Chebyshev1(5, 1)),
               lambda: cmplx() >> Filt(Lowpass, 6 * Hz),
               lambda: cmplx() >> Filt(Bandpass, 20 * Hz, 30 * Hz, method=Chebyshev1(5, 1)),
               lambda: cmplx() >> Filt(Bandstop, 2 * Hz, 12 * Hz, method=Chebyshev1(5, 1)),
               lambda: cmplx() >> Filt(Highpass, 8 * Hz, method=Chebyshev1(5, 1), blocksize=64) >> After(1 * s),
               lambda: cmplx() >> Filt(Lowpass, 6 * Hz) >> Filt(Highpass, 8 * Hz, method=Chebyshev1(5, 1))):
        check(gpu, mk)


@pytest.mark.parametrize("n,nch,order,fc", [(44100, 1, 5, 2000.0), (480000, 2, 8, 4000.0),
                                            (100000, 3, 3, 300.0), (70001, 2, 16, 6000.0),
                                            (31, 1, 8, 4000.0), (4097, 2, 2, 20.0)])
def test_iir_lowpass_sizes(gpu, n, nch, order, fc):
    x = rng(n).standard_normal((n, nch))
    check(gpu, lambda: Signal(x, 48 * kHz) >> Filt(Lowpass, fc * Hz, order=order) >> Amplify(-20 * dB))


def test_iir_slow_decay_uses_the_carry_matrix(gpu):
    # 2 Hz high-pass at 48 kHz: the zero-input response outlives any chunk
    x = rng(11).standard_normal((300000, 2)) + 0.5
    check(gpu, lambda: Signal(x, 48 * kHz) >> Filt(Highpass, 2 * Hz, order=5))
    check(gpu, lambda: Signal(x, 48 * kHz) >> Filt(Bandpass, 500 * Hz, 4 * kHz) >> Ramp() >> Normpower)


def test_readme_scene(gpu):
    noise = rng().standard_normal(44100)

    def scene():
        x = Signal(sin, ω=1 * kHz) >> Until(1 * s) >> Ramp() >> Normpower >> Amplify(-20 * dB + 5 * dB)
        y = Signal(noise) >> Until(1 * s) >> Filt(Bandstop, 0.5 * kHz, 2 * kHz) >> Normpower >> Amplify(-20 * dB)
        return Mix(x, y) >> ToFramerate(44.1 * kHz)
    got = check(gpu, scene)
    assert got.shape == (44100, 1)


def test_randn_leaf_draws_match(gpu):
    def chain(seed):
        return (Signal(randn, rng=rng(seed)) >> Until(4 * s) >> After(50 * ms) >> Filt(Lowpass, 5 * Hz)
                >> Mix(Signal(sin, ω=7 * Hz)) >> Until(3.5 * s) >> Filt(Highpass, 2 * Hz)
                >> Append(rng(3).random((10, 2))) >> Append(rng(4).random((5, 2))) >> ToFramerate(20 * Hz))
    want, fs = oracle.sink(chain(7))
    got, fs2 = sink(chain(7), gpu)
    assert fs == fs2 == 20.0 and got.shape == want.shape == (85, 2)
    assert np.max(np.abs(got - want)) <= F64_TOL * rms(want)


def test_fused_am_noise_bandpass_ramp_mix(gpu):
    # BASELINE config 5 shape, small: prologue load*modulator, epilogue *ramp + tone
    x = rng(5).standard_normal((96000, 4))

    def chain():
        am = Amplify(Signal(x, 96 * kHz), Signal(AffineSin(0.5, 0.5), ω=5 * Hz)) >> Until(1 * s)
        return am >> Filt(Bandpass, 500 * Hz, 4 * kHz) >> Ramp(10 * ms) >> Mix(Signal(sin, ω=1 * kHz) >> Until(1 * s))
    check(gpu, chain)


# ---- K4/K5: resampling ----------------------------------------------------------------------------------

@pytest.mark.parametrize("fs_in,fs_out,n,nch", [(44100, 48000, 44100, 2), (48000, 44100, 48000, 1),
                                                (20, 40, 100, 2), (20, 15, 100, 2), (1000, 500, 3000, 3),
                                                (1000, 3000, 1000, 1), (3000, 1000, 3001, 2),
                                                (1000, 1500, 999, 2), (1500, 1000, 1200, 1),
                                                (1000, 3141.592653589793, 2000, 2), (10, 2000, 10, 2)])
def test_resample(gpu, fs_in, fs_out, n, nch):
    x = rng(n).standard_normal((n, nch))
    got = check(gpu, lambda: ToFramerate(Signal(x, fs_in * Hz), fs_out * Hz))
    assert got.shape[0] == int(np.ceil(n * fs_out / fs_in))


def test_resample_then_more(gpu):
    x = rng(8).standard_normal((8000, 2))
    check(gpu, lambda: ToFramerate(Signal(x, 8 * kHz), 11.025 * kHz) >> Normpower >> Amplify(-20 * dB))
    check(gpu, lambda: Signal(sin, 20 * Hz, ω=5 * Hz) >> ToChannels(2) >> Until(5 * s) >> Pad(one) >> Until(7 * s)
          >> ToFramerate(40 * Hz))
    check(gpu, lambda: Mix(Signal(x, 8 * kHz), Signal(sin, 16 * kHz, ω=440 * Hz) >> Until(1 * s)))


# ---- Float32 ------------------------------------------------------------------------------------------------

def test_float32_stays_float32(gpu):
    x = rng().random((100, 2)).astype(np.float32)
    y = rng(2).random((50, 2)).astype(np.float32)
    X, Y = (lambda: Signal(x, 10 * Hz)), (lambda: Signal(y, 10 * Hz))
    for mk in (lambda: X() >> Until(5 * s), lambda: X() >> Append(Y()) >> After(2 * s),
               lambda: X() >> Pad(zero) >> Until(15 * s), lambda: X() >> Filt(Lowpass, 3 * Hz),
               lambda: X() >> Normpower >> Amplify(np.float32(-10) * dB), lambda: X() >> Mix(Y()),
               lambda: X() >> AddChannel(Y()), lambda: X() >> SelectChannel(1), lambda: X() >> Ramp(),
               lambda: X() >> FadeTo(Y())):
        got = check(gpu, mk, tol=F32_TOL)
        assert got.dtype == np.float32
    got = check(gpu, lambda: Signal(rng().random((10, 2)), 10 * Hz) >> ToEltype(np.float32), tol=F32_TOL)
    assert got.dtype == np.float32


def test_more_reference_sets(gpu):
    """Chains of runtests.jl:103-114 (channel counts), 461-470 (automatic reformatting),
    481-488 / 550-575 (empty and cut infinite signals), 602-614 (frame units)."""
    tone = lambda: Signal(sin, 22 * Hz, ω=10 * Hz) >> Until(5 * s)   # noqa: E731
    data = check(gpu, lambda: tone() >> ToChannels(2), tol=1e-10)
    check(gpu, lambda: Signal(data, 22 * Hz) >> ToChannels(1), exact=True)
    a = lambda: Signal(sin, 200 * Hz, ω=10 * Hz) >> ToChannels(2) >> Until(5 * s)   # noqa: E731
    b = lambda: Signal(sin, 100 * Hz, ω=5 * Hz) >> Until(3 * s)                     # noqa: E731
    assert check(gpu, lambda: Mix(a(), b()), tol=1e-10).shape == (1000, 2)
    assert check(gpu, lambda: Mix(a(), b(), 1), tol=1e-10).shape == (1000, 2)
    for nch in (1, 2):
        t0 = lambda: Signal(sin, 200 * Hz, ω=10 * Hz) >> ToChannels(nch)   # noqa: E731
        assert check(gpu, lambda: t0() >> Until(10 * frames) >> Until(0 * frames)).shape == (0, nch)
        assert check(gpu, lambda: t0() >> Until(10 * frames) >> After(5 * frames) >> After(2 * frames), tol=1e-10).shape == (3, nch)
        assert check(gpu, lambda: t0() >> After(5 * frames) >> Until(5 * frames), tol=1e-10)[0, 0] > 0.9
    x, y = rng(3).random((100, 2)), rng(4).random((50, 2))
    X, Y = (lambda: Signal(x, 10 * Hz)), (lambda: Signal(y, 10 * Hz))
    for mk, n in ((lambda: X() >> Until(30 * frames), 30), (lambda: X() >> After(30 * frames), 70),
                  (lambda: X() >> Append(Y()) >> After(20 * frames), 130), (lambda: X() >> Append(Y()) >> Until(130 * frames), 130),
                  (lambda: X() >> Pad(zero) >> Until(150 * frames), 150), (lambda: X() >> Ramp(10 * frames), 100)):
        assert check(gpu, mk).shape[0] == n
    assert check(gpu, lambda: X() >> FadeTo(Y(), 10 * frames)).shape[0] > 100


def test_float32_resample(gpu):
    """Float32 data through ToFramerate: widened for the FIR stage, rounded after it, Float32 out."""
    x = rng(5).standard_normal((3000, 2)).astype(np.float32)
    for mk in (lambda: ToFramerate(Signal(x, 8 * kHz), 11.025 * kHz),
               lambda: ToFramerate(Signal(x, 8 * kHz), 4 * kHz) >> Amplify(np.float32(-6) * dB),
               lambda: ToFramerate(Signal(x, 8 * kHz), 12 * kHz) >> Normpower):
        got = check(gpu, mk, tol=F32_TOL)
        assert got.dtype == np.float32


# ---- batches -----------------------------------------------------------------------------------------------------

def test_batch_matches_single(gpu):
    xs = [rng(k).standard_normal((20000, 2)) for k in range(9)]

    def chain(x):
        return Signal(x, 48 * kHz) >> Filt(Lowpass, 4 * kHz, order=8) >> Amplify(-20 * dB)
    outs = sink_batch([chain(x) for x in xs], gpu)
    for x, (o, fs) in zip(xs, outs):
        want, _ = oracle.sink(chain(x))
        assert fs == 48000.0 and np.max(np.abs(o - want)) <= F64_TOL * rms(want)


def test_errors(gpu):
    from signalops import SignalError
    with pytest.raises(SignalError):
        sink(Signal(sin, 200 * Hz), gpu)                             # infinite
    with pytest.raises(SignalError):
        sink(Signal(np.arange(10.0), 5 * Hz) >> After(3 * s), gpu)   # too short to skip
    with pytest.raises(SignalError):
        sink(Signal(sin, 200 * Hz) >> Normpower >> Until(1 * s), gpu)
    with pytest.raises(SignalError):
        sink_into(np.ones((10, 2)), np.ones((5, 2)), gpu)


# ---- host arrays that are not dense (strides must not be ignored) -----------------

def test_strided_views_are_read_through_their_strides(gpu):
    x = rng(21).standard_normal(4000)
    st = rng(22).standard_normal((2000, 2))                 # C-ordered: st[:, 0] has a stride of 2 samples
    check(gpu, lambda: Signal(x[::2], 1 * kHz) >> Filt(Lowpass, 100 * Hz))
    check(gpu, lambda: Signal(x[::-1], 1 * kHz) >> Amplify(-6 * dB), tol=1e-12)
    check(gpu, lambda: Signal(st[:, 0], 1 * kHz) >> Mix(Signal(st[:, 1], 1 * kHz)), tol=1e-12)
    check(gpu, lambda: Signal(st[::3], 1 * kHz) >> Filt(Highpass, 50 * Hz))
    got = sink(Signal(x[::2], 1 * kHz), gpu)[0]
    assert np.array_equal(got[:, 0], x[::2])


def test_sink_into_strided_and_narrow_results(gpu):
    sig = Signal(np.arange(20.0), 10 * Hz)
    parent = np.full(40, -1.0)
    view = parent[::2]
    sink_into(view, sig, gpu)
    assert np.array_equal(view, np.arange(20.0)) and np.all(parent[1::2] == -1.0)      # neighbours untouched
    for dt in (np.int32, np.int16, np.uint8):
        guard = np.full(30, 7, dtype=dt)
        res = guard[5:25]
        sink_into(res, Signal(np.arange(20), 10 * Hz), gpu)
        assert np.array_equal(res, np.arange(20).astype(dt))
        assert np.all(guard[:5] == 7) and np.all(guard[25:] == 7)                      # no write past the result
    r32 = np.zeros(20, dtype=np.int32)
    sink_into(r32, sig, gpu)                                                           # integral floats convert
    assert np.array_equal(r32, np.arange(20))
    from signalops import SignalError
    with pytest.raises(SignalError):                                                   # Julia: InexactError
        sink_into(np.zeros(20, dtype=np.int32), Signal(np.arange(20.0) + 0.5, 10 * Hz), gpu)
    m = np.zeros((20, 2), dtype=np.float32)[:, ::-1]                                   # negative channel stride
    sink_into(m, Signal(np.arange(40.0).reshape(20, 2), 10 * Hz), gpu)
    assert np.array_equal(m, np.arange(40.0).reshape(20, 2).astype(np.float32))


# ---- raw filter objects: Filt(h) with DSP.jl coefficient types (src/filters.jl:65,89-95) --------

def test_raw_filter_objects(gpu):
    from signalops import Biquad, Butterworth, PolynomialRatio, digitalfilter
    x = rng(31).standard_normal((3000, 2))
    h = digitalfilter(Highpass(8, fs=100), Chebyshev1(5, 1))
    a = check(gpu, lambda: Signal(x, 100 * Hz) >> Filt(h))
    b = check(gpu, lambda: Signal(x, 100 * Hz) >> Filt(Highpass, 8 * Hz, method=Chebyshev1(5, 1)))
    assert np.max(np.abs(a - b)) <= 1e-12 * rms(b)                                    # runtests.jl:364-368
    check(gpu, lambda: Signal(x, 100 * Hz) >> Filt(digitalfilter(Lowpass(20, fs=100), Butterworth(4))))
    check(gpu, lambda: Signal(x, 100 * Hz) >> Filt(Biquad(0.2, 0.3, 0.1, -0.5, 0.25)))
    check(gpu, lambda: Signal(x, 100 * Hz) >> Filt(PolynomialRatio([0.3, 0.2], [1.0, -0.4, 0.1])))


# ---- operator leftovers: order-n PolynomialRatio, single-rate FIR, whole-frame `reverse` ----

def test_polynomial_ratio_of_any_order(gpu):
    """src/filters.jl:65,89-95: `Filt(x, PolynomialRatio(b, a))` runs DSP.jl's order-n DF2T recurrence; the GPU
    runs the same transfer function factored into biquads (oracle: scipy.signal.lfilter, the direct form)."""
    from scipy import signal as sps
    from signalops import PolynomialRatio
    x = rng(41).standard_normal((6000, 2))
    for b, a in (sps.butter(5, 0.3), sps.cheby1(7, 1, 0.2), ([0.0, 0.5, 0.2, 0.1], [1, -0.3, 0.2, 0.05, 0.01]),
                 ([1, 0.4, 0.3, 0.2, 0.1, 0.05], [1.0, -0.2])):
        got = check(gpu, lambda: Signal(x, 100 * Hz) >> Filt(PolynomialRatio(b, a)), tol=1e-9)
        ref = np.stack([sps.lfilter(b, a, x[:, c]) for c in range(2)], axis=1)
        assert np.max(np.abs(got - ref)) <= 1e-9 * rms(ref)


def test_single_rate_fir_filter(gpu):
    """`Filt(x, FIRFilter(h))` (DSP.jl FIRStandard through src/filters.jl:240-262): every FIR kernel of the library."""
    from scipy import signal as sps
    from signalops import dspjl
    h = sps.firwin(31, 0.3)
    x = rng(42).standard_normal((7001, 2))
    got = check(gpu, lambda: Signal(x, 100 * Hz) >> Filt(dspjl.FIRFilter(h, 1)) >> Amplify(-6 * dB))
    ref = np.stack([sps.lfilter(h, 1.0, x[:, c]) for c in range(2)], axis=1) * 10 ** (-6 / 20)
    assert got.shape == (7001, 2) and np.max(np.abs(got - ref)) <= 1e-12 * rms(ref)
    xs = [rng(43 + k).standard_normal((5000, 2)) for k in range(40)]          # 80 rows: the tensor-map kernel
    outs = sink_batch([Signal(v, 100 * Hz) >> Filt(dspjl.FIRFilter(h, 1)) for v in xs], gpu)
    for k in (0, 39):
        ref = np.stack([sps.lfilter(h, 1.0, xs[k][:, c]) for c in range(2)], axis=1)
        assert np.max(np.abs(outs[k][0] - ref)) <= 1e-12 * rms(ref)


def test_whole_frame_reverse(gpu):
    """test/runtests.jl:273-276: `OperateOn(reverse, x, bychannel=false)` swaps the channels."""
    from signalops import reverse
    x = rng(44).random((20, 2))
    got = sink(OperateOn(reverse, x, bychannel=False) >> ToFramerate(20 * Hz), gpu)
    assert np.array_equal(got[0], np.stack([x[:, 1], x[:, 0]], axis=1))
    y = rng(45).random((50, 5))
    check(gpu, lambda: OperateOn(reverse, Signal(y, 10 * Hz), bychannel=False) >> Amplify(2), exact=True)


def test_sink_into_c_ordered_result_in_place(gpu):
    """`sink!` into numpy's own (nframes, nchannels) layout (src/sink.jl:158-168): written frame-interleaved by the device
    (the WAV layout path, Float64 and Float32), no temporary on the host."""
    x = rng(61).standard_normal((4000, 3))
    for dt in (np.float64, np.float32):
        sig = Signal(x.astype(dt), 48 * kHz) >> Filt(Lowpass, 4 * kHz) >> Amplify(-6 * dB)
        res = np.full((3500, 3), np.nan, dtype=dt)
        assert sink_into(res, sig, gpu) is res
        want = np.empty((3500, 3), dtype=dt)
        oracle.sink_into(want, sig)
        tol = F64_TOL if dt == np.float64 else F32_TOL
        assert not np.isnan(res).any()
        assert np.max(np.abs(res.astype(np.float64) - want)) <= tol * rms(want)
