"""BASELINE config 3 at its real length: one-minute stereo signals (2 646 000 frames at 44.1 kHz ->
2 880 000 at 48 kHz, and back), every output sample compared with the C oracle
(oracle/cpu_ref.c `oracle_resample_batch`: the reference's FilteredSignal block loop, blocksize 4096,
over DSP.jl's FIRArbitrary kernel — src/filters.jl:240-262, src/reformatting.jl:92-122).  The
Float64 phase accumulator is stepped 2.88 M times; every FIR kernel of the library is exercised:
tensor-map (k_fir_tmap, batches of > 64 rows), per-row TMA (k_fir_mma) and scalar (k_fir)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import dspjl_ref as D
from signalops import Amplify, Hz, Signal, ToFramerate, dB, sink_batch

pytestmark = pytest.mark.gpu
F64_TOL = 1e-9


def oracle_batch(x, fi, fo, n_out, threads=8):
    """x: (nsig, nch, n_in) -> (nsig, nch, n_out)"""
    r = D.Resampler(fo / fi)
    x = np.ascontiguousarray(x)
    y = np.empty((x.shape[0], x.shape[1], n_out))
    dp = C.POINTER(C.c_double)
    D.lib().oracle_resample_batch(x.ctypes.data_as(dp), y.ctypes.data_as(dp), x.shape[0], x.shape[2], n_out, x.shape[1],
                                  C.byref(r.st), 4096, threads)
    return y


def with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    try:
        os.environ.update(env)
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("fi,fo,n_in", [(44100.0, 48000.0, 2646000), (48000.0, 44100.0, 2880000)])
def test_one_minute_stereo_batch_every_sample(gpu, fi, fo, n_in):
    """34 stereo signals = 68 rows: takes the tensor-map kernel; three whole instances are compared."""
    ninst = 34
    rng = np.random.default_rng(int(fi))
    base = rng.standard_normal((3, 2, n_in))
    n_out = int(np.ceil(n_in * fo / fi))
    # distinct data in the instances that are checked, cheap copies elsewhere
    xs = [np.asfortranarray(base[k % 3].T * (1.0 + k)) for k in range(ninst)]
    got = sink_batch([ToFramerate(Signal(x, fi * Hz), fo * Hz) for x in xs], gpu)
    assert gpu.last_stats["launches"] >= 1
    check = [0, 16, 33]
    want = oracle_batch(np.stack([xs[k].T for k in check]), fi, fo, n_out, threads=3)
    for j, k in enumerate(check):
        y, fs = got[k]
        assert fs == fo and y.shape == (n_out, 2)
        err = float(np.max(np.abs(y.T - want[j])) / np.sqrt(np.mean(want[j] ** 2)))
        assert err <= F64_TOL, (k, err)


@pytest.mark.parametrize("env", [{"SIGOPS_NO_FIR_TMAP": "1"}, {"SIGOPS_NO_FIR_TMAP": "1", "SIGOPS_NO_FIR_MMA": "1"}],
                         ids=["per-row-TMA kernel", "scalar kernel"])
def test_one_minute_stereo_other_kernels(gpu, env):
    fi, fo, n_in = 44100.0, 48000.0, 2646000
    n_out = 2880000
    x = np.asfortranarray(np.random.default_rng(5).standard_normal((n_in, 2)))
    (y, fs), = with_env(env, lambda: sink_batch([ToFramerate(Signal(x, fi * Hz), fo * Hz)], gpu))
    want = oracle_batch(x.T[None], fi, fo, n_out, threads=1)[0]
    assert y.shape == (n_out, 2) and fs == fo
    assert float(np.max(np.abs(y.T - want)) / np.sqrt(np.mean(want ** 2))) <= F64_TOL


def test_resample_then_gain_is_one_launch(gpu):
    """`ToFramerate |> Amplify(c)`: the constant gain is folded into the taps of the tensor-core kernels."""
    rng = np.random.default_rng(11)
    xs = [np.asfortranarray(rng.standard_normal((30000, 2))) for _ in range(40)]      # 80 rows: tensor-map kernel
    chain = lambda x: ToFramerate(Signal(x, 44100.0 * Hz), 48000.0 * Hz) >> Amplify(-6 * dB)   # noqa: E731
    got = sink_batch([chain(x) for x in xs], gpu)
    assert gpu.last_stats["launches"] == 1 * 16 or gpu.last_stats["launches"] <= 16   # one kernel per wave
    n_out = int(np.ceil(30000 * 48000 / 44100))
    want = oracle_batch(np.stack([xs[k].T for k in (0, 39)]), 44100.0, 48000.0, n_out, threads=2) * 10 ** (-6 / 20)
    for j, k in enumerate((0, 39)):
        assert float(np.max(np.abs(got[k][0].T - want[j])) / np.sqrt(np.mean(want[j] ** 2))) <= F64_TOL
    # small batch (per-row kernel) takes the same fused form
    (y, _), = sink_batch([chain(xs[0])], gpu)
    assert gpu.last_stats["launches"] == 1
    assert float(np.max(np.abs(y.T - want[0])) / np.sqrt(np.mean(want[0] ** 2))) <= F64_TOL


def test_period_table_and_rebuilt_bands_agree(gpu):
    """A periodic resampler (44.1 -> 48 kHz: index pattern repeats every 160 outputs) reads its merged tap bands from a
    host-built period table; SIGOPS_NO_FIR_BANDS=1 makes the helper warps rebuild them per tile from the polyphase banks.
    Both must give the reference's result — they differ only where the drifting Float64 phase accumulator sits on either
    side of an integer phase (continuous in the taps: ~1e-13)."""
    rng = np.random.default_rng(21)
    fi, fo, n_in = 44100.0, 48000.0, 200000
    xs = [np.asfortranarray(rng.standard_normal((n_in, 2))) for _ in range(40)]       # 80 rows: tensor-map kernel
    chain = lambda x: ToFramerate(Signal(x, fi * Hz), fo * Hz)                         # noqa: E731
    a = sink_batch([chain(x) for x in xs], gpu)
    b = with_env({"SIGOPS_NO_FIR_BANDS": "1"}, lambda: sink_batch([chain(x) for x in xs], gpu))
    n_out = a[0][0].shape[0]
    want = oracle_batch(np.stack([xs[k].T for k in (0, 39)]), fi, fo, n_out, threads=2)
    for j, k in enumerate((0, 39)):
        r = np.sqrt(np.mean(want[j] ** 2))
        assert float(np.max(np.abs(a[k][0].T - want[j])) / r) <= F64_TOL
        assert float(np.max(np.abs(b[k][0].T - want[j])) / r) <= F64_TOL
        assert float(np.max(np.abs(a[k][0] - b[k][0])) / r) <= 1e-11


def test_ratio_without_a_short_period_rebuilds_its_bands(gpu):
    """44.1 kHz -> 47 999 Hz: no period within the table limit, so the general path (helper warps) serves the
    tensor-map kernel; every sample against the oracle."""
    rng = np.random.default_rng(22)
    fi, fo, n_in = 44100.0, 47999.0, 150000
    xs = [np.asfortranarray(rng.standard_normal((n_in, 2))) for _ in range(36)]       # 72 rows
    got = sink_batch([ToFramerate(Signal(x, fi * Hz), fo * Hz) for x in xs], gpu)
    n_out = got[0][0].shape[0]
    want = oracle_batch(np.stack([xs[k].T for k in (0, 35)]), fi, fo, n_out, threads=2)
    for j, k in enumerate((0, 35)):
        assert got[k][1] == fo
        assert float(np.max(np.abs(got[k][0].T - want[j])) / np.sqrt(np.mean(want[j] ** 2))) <= F64_TOL
