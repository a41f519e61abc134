"""K4/K5 on the FP64 tensor cores (csrc/k_fir_mma.cuh): batches large enough to take the
128- and 64-row blocks, ragged row groups, every FIR kernel kind, signal ends inside a tile,
and agreement with the scalar-FMA kernel (csrc/k_fir.cuh) it stands in for.

Reference behaviour: ToFramerate on data signals, src/reformatting.jl:92-122 +
DSP.jl FIRFilter kernels."""
import os

import numpy as np
import pytest

import oracle
from signalops import Amplify, Hz, Normpower, Signal, ToFramerate, dB, kHz, sink_batch

pytestmark = pytest.mark.gpu
F64_TOL = 1e-9


def rms(a):
    return float(np.sqrt(np.mean(np.asarray(a, dtype=np.float64) ** 2)))


def run_batch(gpu, xs, chain, no_mma=False):
    old = os.environ.pop("SIGOPS_NO_FIR_MMA", None)
    try:
        if no_mma:
            os.environ["SIGOPS_NO_FIR_MMA"] = "1"
        return sink_batch([chain(x) for x in xs], gpu)
    finally:
        os.environ.pop("SIGOPS_NO_FIR_MMA", None)
        if old is not None:
            os.environ["SIGOPS_NO_FIR_MMA"] = old


@pytest.mark.parametrize("fs_in,fs_out,n,nch,ninst", [
    (44100, 48000, 4411, 2, 70),      # arbitrary ratio; 140 rows = one full 128-row block + a ragged one
    (48000, 44100, 5003, 2, 20),      # arbitrary ratio, down; 40 rows -> 64-row blocks
    (1000, 1500, 2999, 1, 130),       # rational 3/2; single channel
    (1000, 500, 6001, 2, 33),         # decimator; 66 rows
    (1000, 2000, 1777, 3, 11),        # interpolator; 33 rows
    (1000, 3141.592653589793, 900, 2, 5),   # irrational ratio, 10 rows -> 32-row blocks
])
def test_batched_resample_matches_oracle_and_scalar_kernel(gpu, fs_in, fs_out, n, nch, ninst):
    rng = np.random.default_rng(n)
    xs = [rng.standard_normal((n, nch)) for _ in range(ninst)]
    chain = lambda x: ToFramerate(Signal(x, fs_in * Hz), fs_out * Hz)   # noqa: E731
    got = run_batch(gpu, xs, chain)
    ref = run_batch(gpu, xs, chain, no_mma=True)
    nout = int(np.ceil(n * fs_out / fs_in))
    for k, ((y, fs), (yr, _)) in enumerate(zip(got, ref)):
        assert y.shape == (nout, nch) and fs == float(fs_out)
        # both kernels do the same products in a different summation order
        assert np.max(np.abs(y - yr)) <= 1e-12 * max(rms(yr), 1e-300)
        if k in (0, ninst // 2, ninst - 1):          # the Python oracle is slow: spot-check instances
            want, _ = oracle.sink(chain(xs[k]))
            assert np.max(np.abs(y - want)) <= F64_TOL * rms(want), k


def test_batched_resample_feeds_normpower(gpu):
    """The tensor-core kernel also accumulates the sum of squares the next stage divides by."""
    rng = np.random.default_rng(7)
    xs = [rng.standard_normal((3000, 2)) * (1 + k) for k in range(40)]
    chain = lambda x: ToFramerate(Signal(x, 8 * kHz), 11.025 * kHz) >> Normpower >> Amplify(-20 * dB)   # noqa: E731
    got = run_batch(gpu, xs, chain)
    for k in (0, 17, 39):
        want, _ = oracle.sink(chain(xs[k]))
        assert np.max(np.abs(got[k][0] - want)) <= F64_TOL * rms(want)
        assert abs(rms(got[k][0]) - 0.1) < 1e-12


def test_long_signal_many_segments(gpu):
    """One long stereo pair: 32-row block, the time axis cut into one segment per SM."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal((200000, 2))
    chain = lambda v: ToFramerate(Signal(v, 44.1 * kHz), 48 * kHz)   # noqa: E731
    (y, fs), = run_batch(gpu, [x], chain)
    (yr, _), = run_batch(gpu, [x], chain, no_mma=True)
    assert fs == 48000.0 and y.shape == yr.shape
    assert np.max(np.abs(y - yr)) <= 1e-12 * rms(yr)
    want, _ = oracle.sink(chain(x[:30000]))
    # the FIR is causal: the head of the long signal equals the resampled head (minus the filter tail)
    assert np.max(np.abs(y[:30000] - want[:30000])) <= F64_TOL * rms(want)
