"""How caller arrays reach the C ABI (host/lowering.py `add_input`, host/gpusink.py `_colmajor_array`): Julia's
column-major matrices as they are, numpy's C-ordered (nframes, nchannels) matrices as frame-interleaved buffers
(SIGOPS_INTERLEAVED, transposed on the device — no host copy), strided views as dense copies."""
import numpy as np

import oracle.cpu_sink as oracle
from plan_emulator import Emulator
from signalops import Amplify, Filt, Lowpass, Signal, dB, kHz
from signalops.lowering import lower
from signalops.wav import WavRaw


def chain(x):
    return Signal(x, 48 * kHz) >> Filt(Lowpass, 4 * kHz, order=4) >> Amplify(-6 * dB)


def test_c_ordered_matrix_is_passed_interleaved_without_a_copy():
    rng = np.random.default_rng(3)
    for dt in (np.float64, np.float32):
        x = rng.standard_normal((5000, 3)).astype(dt)                    # numpy's default layout
        plan = lower(chain(x))
        (a,) = plan.input_arrays
        assert isinstance(a, WavRaw) and np.shares_memory(a.raw, x) and a.raw.dtype == dt
        assert (plan.inputs[0].nframes, plan.inputs[0].nchannels) == (5000, 3)
        got = Emulator(plan.tobytes()).run(plan.input_arrays)[0]
        want, _ = oracle.sink(chain(x))
        tol = 1e-12 if dt == np.float64 else 1e-5
        assert np.max(np.abs(got - want)) <= tol * np.sqrt(np.mean(want.astype(np.float64) ** 2))


def test_column_major_and_mono_arrays_are_passed_as_they_are():
    rng = np.random.default_rng(4)
    xf = np.asfortranarray(rng.standard_normal((5000, 2)))
    (a,) = lower(chain(xf)).input_arrays
    assert isinstance(a, np.ndarray) and np.shares_memory(a, xf)
    mono = rng.standard_normal(5000)
    (a,) = lower(chain(mono)).input_arrays
    assert isinstance(a, np.ndarray) and np.shares_memory(a, mono)
    col = rng.standard_normal((5000, 1))
    (a,) = lower(chain(col)).input_arrays
    assert isinstance(a, np.ndarray) and np.shares_memory(a, col)


def test_strided_views_become_dense_copies():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((10000, 2))
    view = x[::2]                                                        # neither layout
    (a,) = lower(chain(view)).input_arrays
    assert isinstance(a, np.ndarray) and not np.shares_memory(a, x) and a.flags.f_contiguous
    ints = (rng.standard_normal((100, 2)) * 100).astype(np.int32)        # integers are widened to Int64 first
    (a,) = lower(Signal(ints, 48 * kHz)).input_arrays
    assert isinstance(a, np.ndarray) and a.dtype == np.int64


def test_c_ordered_result_is_written_in_place():
    """`sink!` into numpy's own (nframes, nchannels) layout: the device writes the frame-interleaved result directly."""
    from signalops import sink_into
    from test_lowering_emulated import EmulatedSink
    rng = np.random.default_rng(6)
    x = rng.standard_normal((3000, 2))
    for dt in (np.float64, np.float32):
        res = np.full((2500, 2), np.nan, dtype=dt)                       # C order
        out = sink_into(res, chain(x.astype(dt)), EmulatedSink())
        assert out is res and not np.isnan(res).any()
        want = np.empty((2500, 2), dtype=dt)
        oracle.sink_into(want, chain(x.astype(dt)))
        tol = 1e-12 if dt == np.float64 else 1e-5
        assert np.max(np.abs(res.astype(np.float64) - want)) <= tol * np.sqrt(np.mean(want.astype(np.float64) ** 2))
