"""Physical sanity of the resampler, independent of any restated DSP.jl detail: a tone in must be
the same tone out at the new rate with ZERO net delay (what `setphase!(timedelay)` at
src/reformatting.jl:92-99 is for), the DC gain must be 1, and the error must sit at the
60 dB design spec of `resample_filter`.  A shared misreading of the
`timedelay` convention ((hLen-1)/(2 Nphi) vs hLen/(2 Nphi)), of `setphase!` or of the tap layout would
shift or scale the output and fail here, for the oracle and for the GPU alike.

The CPU half runs on the oracle (C restatement); the `gpu` half runs the same checks through
sink(x, GPUSink())."""
import numpy as np
import pytest

from oracle import dspjl_ref as D
from signalops import Hz, Signal, ToFramerate, sink

RATES = [(44100.0, 48000.0), (48000.0, 44100.0)]
TONES = [1000.0, 15000.0]
SPEC = 10 ** (-60 / 20) * 1.2      # 60 dB stop band / pass-band ripple of the Kaiser design, 20 % margin


def tone_error(y, f, fs_out, delay=0.0, skip=2000):
    m = np.arange(len(y))
    ref = np.sin(2 * np.pi * f * (m + delay) / fs_out)
    sl = slice(skip, len(y) - skip)
    return float(np.max(np.abs(y[sl] - ref[sl])))


def best_delay(y, f, fs_out, skip=2000):
    """Delay (in output samples) that minimises the error against the ideal tone, by a fine scan."""
    ds = np.linspace(-0.05, 0.05, 101)
    errs = [tone_error(y, f, fs_out, d, skip) for d in ds]
    return float(ds[int(np.argmin(errs))])


def oracle_resample(x, fi, fo):
    r = D.Resampler(fo / fi)
    n_out = int(np.ceil(len(x) * fo / fi))
    return r.filt(np.concatenate([x, np.zeros(4096)]))[:n_out]


def gpu_resample(gpu, x, fi, fo):
    y, fs = sink(ToFramerate(Signal(x, fi * Hz), fo * Hz), gpu)
    assert fs == fo
    return y[:, 0]


def check_tone(resample, fi, fo, f):
    n = int(fi)                                    # one second
    x = np.sin(2 * np.pi * f * np.arange(n) / fi)
    y = resample(x, fi, fo)
    assert len(y) == int(np.ceil(n * fo / fi))     # reformatting.jl / filters.jl:159-167 length rule
    assert tone_error(y, f, fo) < SPEC, "tone does not come out at the same frequency/phase within the 60 dB spec"
    # zero net delay: the best-fitting delay is 0 to within 1/500 of an output sample; the alternative
    # `timedelay` convention would sit at 1/(2*32) of an input sample = 0.0156
    assert abs(best_delay(y, f, fo)) <= 0.002


def check_dc(resample, fi, fo):
    y = resample(np.ones(20000), fi, fo)
    mid = y[2000:-2000]
    assert abs(float(np.mean(mid)) - 1.0) < 1e-4 and float(np.max(np.abs(mid - 1.0))) < SPEC


@pytest.mark.parametrize("fi,fo", RATES)
@pytest.mark.parametrize("f", TONES)
def test_oracle_tone_in_tone_out(fi, fo, f):
    check_tone(oracle_resample, fi, fo, f)


@pytest.mark.parametrize("fi,fo", RATES)
def test_oracle_dc_gain(fi, fo):
    check_dc(oracle_resample, fi, fo)


def test_oracle_rational_ratios_tone():
    """Exact-rational kernels (1/2, 2, 3/2): same property, coarser spec near Nyquist."""
    for fi, fo in [(1000.0, 500.0), (1000.0, 2000.0), (1000.0, 1500.0)]:
        f = 50.0
        x = np.sin(2 * np.pi * f * np.arange(4000) / fi)
        from fractions import Fraction
        r = D.Resampler(Fraction(int(fo), int(fi)))
        y = r.filt(np.concatenate([x, np.zeros(2048)]))[: int(np.ceil(4000 * fo / fi))]
        assert tone_error(y, f, fo, skip=400) < 2e-3
        assert abs(best_delay(y, f, fo, skip=400)) <= 0.01


@pytest.mark.gpu
@pytest.mark.parametrize("fi,fo", RATES)
@pytest.mark.parametrize("f", TONES)
def test_gpu_tone_in_tone_out(gpu, fi, fo, f):
    check_tone(lambda x, a, b: gpu_resample(gpu, x, a, b), fi, fo, f)


@pytest.mark.gpu
@pytest.mark.parametrize("fi,fo", RATES)
def test_gpu_dc_gain(gpu, fi, fo):
    check_dc(lambda x, a, b: gpu_resample(gpu, x, a, b), fi, fo)
