"""K3 through 2-D/3-D tensor-map TMA (csrc/k_iir_tmap.cuh; SIGOPS_NO_TMAP=1 switches it off): same results
as the default per-lane TMA kernel and as the oracle, including ragged row groups, inputs shorter
than the output (zero padding comes from the tensor bounds) and a following Normpower.

Reference behaviour: Filt on data signals, src/filters.jl:204-262 + DSP.jl DF2T SOS filt!."""
import os

import numpy as np
import pytest

import oracle
from signalops import Amplify, Filt, Highpass, Hz, Lowpass, Normpower, Pad, Signal, Until, dB, kHz, s, sink_batch, zero

pytestmark = pytest.mark.gpu
F64_TOL = 1e-9


def rms(a):
    return float(np.sqrt(np.mean(np.asarray(a, dtype=np.float64) ** 2)))


def run(gpu, chains, tmap):
    """One wave per batch, so that the whole batch is one [rows][frames] matrix for the tensor map."""
    saved = {k: os.environ.pop(k, None) for k in ("SIGOPS_NO_TMAP", "SIGOPS_HOST_WAVES")}
    try:
        os.environ["SIGOPS_HOST_WAVES"] = "1"
        if not tmap:
            os.environ["SIGOPS_NO_TMAP"] = "1"
        return sink_batch(chains, gpu)
    finally:
        for k, v in saved.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v


@pytest.mark.parametrize("ninst,nch,n", [(16, 2, 48000), (160, 2, 20000), (300, 1, 16016)])
def test_lowpass_gain_batch(gpu, ninst, nch, n):
    rng = np.random.default_rng(ninst)
    xs = [rng.standard_normal((n, nch)) for _ in range(ninst)]
    chain = lambda x: Signal(x, 48 * kHz) >> Filt(Lowpass, 4 * kHz, order=8) >> Amplify(-20 * dB)   # noqa: E731
    a = run(gpu, [chain(x) for x in xs], tmap=True)
    b = run(gpu, [chain(x) for x in xs], tmap=False)
    for k in range(ninst):
        assert a[k][0].shape == (n, nch)
        assert np.max(np.abs(a[k][0] - b[k][0])) <= 1e-12 * rms(b[k][0])
    for k in (0, ninst - 1):
        want, _ = oracle.sink(chain(xs[k]))
        assert np.max(np.abs(a[k][0] - want)) <= F64_TOL * rms(want)


def test_padded_input_and_normpower(gpu):
    """Input shorter than the output (Pad(zero) |> Until) and the sum of squares for Normpower."""
    rng = np.random.default_rng(5)
    xs = [rng.standard_normal((30000, 2)) for _ in range(32)]
    chain = lambda x: (Signal(x, 48 * kHz) >> Pad(zero) >> Until(1 * s) >> Filt(Highpass, 1 * kHz, order=4)   # noqa: E731
                       >> Normpower >> Amplify(-10 * dB))
    a = run(gpu, [chain(x) for x in xs], tmap=True)
    for k in (0, 13, 31):
        want, _ = oracle.sink(chain(xs[k]))
        assert a[k][0].shape == want.shape
        assert np.max(np.abs(a[k][0] - want)) <= F64_TOL * rms(want)


def test_slow_decay_batch_takes_short_chunks_and_the_carry_pass(gpu):
    """A band-stop whose response needs ~3300 frames to die out, on a batch: the planner picks
    MAIN + CARRY + FIX with chunks far shorter than the decay (csrc/runtime.cu cost model)."""
    from signalops import Bandstop
    rng = np.random.default_rng(11)
    xs = [rng.standard_normal((44100, 2)) for _ in range(40)]
    chain = lambda x: Signal(x, 44.1 * kHz) >> Filt(Bandstop, 0.5 * kHz, 2 * kHz) >> Normpower >> Amplify(-20 * dB)   # noqa: E731
    got = run(gpu, [chain(x) for x in xs], tmap=True)
    for k in (0, 21, 39):
        want, _ = oracle.sink(chain(xs[k]))
        assert got[k][0].shape == want.shape
        assert np.max(np.abs(got[k][0] - want)) <= F64_TOL * rms(want)


@pytest.mark.parametrize("ninst,nch,n", [(16, 2, 48000), (40, 2, 16032)])
def test_float32_batch(gpu, ninst, nch, n):
    """Float32 signals stay Float32 (runtests.jl:707-729); state and arithmetic are Float64."""
    rng = np.random.default_rng(ninst + 1)
    xs = [rng.standard_normal((n, nch)).astype(np.float32) for _ in range(ninst)]
    chain = lambda x: Signal(x, 48 * kHz) >> Filt(Lowpass, 4 * kHz, order=8) >> Amplify(np.float32(-20) * dB)   # noqa: E731
    a = run(gpu, [chain(x) for x in xs], tmap=True)
    b = run(gpu, [chain(x) for x in xs], tmap=False)
    for k in range(ninst):
        assert a[k][0].dtype == np.float32 and a[k][0].shape == (n, nch)
        assert np.max(np.abs(a[k][0].astype(np.float64) - b[k][0])) <= 2e-7 * rms(b[k][0]) + 1e-12
    for k in (0, ninst - 1):
        want, _ = oracle.sink(chain(xs[k]))
        assert want.dtype == np.float32
        assert np.max(np.abs(a[k][0].astype(np.float64) - want)) <= 1e-5 * rms(want)


def test_padded_input_through_the_tensor_map_kernel(gpu):
    """Pad(zero) |> Until |> Filt on a batch: the padded signal is materialised by an elementwise stage
    and the tensor-map kernel filters that temporary (frame counts multiples of 16)."""
    from signalops import frames
    rng = np.random.default_rng(21)
    xs = [rng.standard_normal((32000, 2)) for _ in range(16)]
    chain = lambda x: Signal(x, 48 * kHz) >> Pad(zero) >> Until(48000 * frames) >> Filt(Lowpass, 4 * kHz, order=8) >> Amplify(-3 * dB)   # noqa: E731
    a = run(gpu, [chain(x) for x in xs], tmap=True)
    b = run(gpu, [chain(x) for x in xs], tmap=False)
    for k in range(16):
        assert a[k][0].shape == (48000, 2)
        assert np.max(np.abs(a[k][0] - b[k][0])) <= 1e-12 * rms(b[k][0])
    want, _ = oracle.sink(chain(xs[7]))
    assert np.max(np.abs(a[7][0] - want)) <= F64_TOL * rms(want)


def test_fused_row_invariant_programs(gpu):
    """BASELINE config 5's shape on the tensor-map kernel: AM noise |> Filt(Bandpass) |> Ramp |> Mix(tone) with the
    modulator, the ramps and the tone (all functions of the frame only) evaluated once per warp and stage —
    src/functions.jl:53-60, src/ramps.jl:56-72, src/filters.jl:221-262.  One launch, one HBM round trip."""
    import os
    from signalops import AffineSin, Bandpass, Mix, Ramp, Until, ms, s, sin, sink_batch
    rng = np.random.default_rng(55)
    xs = [np.asfortranarray(rng.standard_normal((96000, 32))) for _ in range(3)]   # 96 rows, 1 s at 96 kHz (column-major: no decode launch)

    def chain(x):
        am = Amplify(Signal(x, 96 * kHz), Signal(AffineSin(0.5, 0.5), ω=5 * Hz)) >> Until(1 * s)
        return am >> Filt(Bandpass, 500 * Hz, 4 * kHz) >> Ramp(10 * ms) >> Mix(Signal(sin, ω=1 * kHz) >> Until(1 * s))

    got = sink_batch([chain(x) for x in xs], gpu)
    assert gpu.last_stats["launches"] == 1
    old = os.environ.get("SIGOPS_NO_TMAP_LEAVES")
    os.environ["SIGOPS_NO_TMAP_LEAVES"] = "1"
    try:
        ref = sink_batch([chain(x) for x in xs], gpu)
    finally:
        os.environ.pop("SIGOPS_NO_TMAP_LEAVES")
        if old is not None:
            os.environ["SIGOPS_NO_TMAP_LEAVES"] = old
    for k in range(3):
        assert got[k][0].shape == (96000, 32) and got[k][1] == 96000.0
        assert np.max(np.abs(got[k][0] - ref[k][0])) <= 1e-11 * rms(ref[k][0])
    want, _ = oracle.sink(chain(xs[1]))
    assert np.max(np.abs(got[1][0] - want)) <= F64_TOL * rms(want)


def test_fused_programs_ragged_row_groups(gpu):
    """The warps of a block share one chunk of 8 consecutive row groups (the leaf vectors are evaluated once per block):
    270 rows = 8 full groups + one group of 14 rows, so the second block runs with one partly filled warp and seven
    idle ones that still take part in the block's barriers."""
    from signalops import AffineSin, Bandpass, Mix, Ramp, Until, ms, s, sin, sink_batch
    rng = np.random.default_rng(56)
    xs = [np.asfortranarray(rng.standard_normal((48000, 30))) for _ in range(9)]   # 270 rows, 0.5 s at 96 kHz

    def chain(x):
        am = Amplify(Signal(x, 96 * kHz), Signal(AffineSin(0.5, 0.5), ω=5 * Hz)) >> Until(0.5 * s)
        return am >> Filt(Bandpass, 500 * Hz, 4 * kHz) >> Ramp(10 * ms) >> Mix(Signal(sin, ω=1 * kHz) >> Until(0.5 * s))

    got = sink_batch([chain(x) for x in xs], gpu)
    assert gpu.last_stats["launches"] == 1
    for k in (0, 4, 8):
        want, _ = oracle.sink(chain(xs[k]))
        assert got[k][0].shape == want.shape
        assert np.max(np.abs(got[k][0] - want)) <= F64_TOL * rms(want)
