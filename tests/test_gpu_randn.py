"""`Signal(randn; rng = PhiloxRNG(seed))` generated on the device (LEAF_RANDN, csrc/interp.cuh `randn_value`) against
the numpy definition of the same generator (host/philox.py, pinned by the Random123 known-answer vectors in
tests/test_philox.py) and against the CPU sink.  Reference: src/functions.jl:98-114 — one `randn(rng)` per frame."""
import os

import numpy as np
import pytest

import oracle
from signalops import (After, Amplify, Bandstop, Filt, GPUSink, Lowpass, Mix, Normpower, PhiloxRNG, Ramp, Signal,
                       ToFramerate, Until, dB, kHz, randn, s, sin, sink, sink_batch)

pytestmark = pytest.mark.gpu
TOL = 1e-9


def rms(a):
    return float(np.sqrt(np.mean(np.asarray(a, dtype=np.float64) ** 2)))


def test_raw_noise_matches_the_numpy_definition(gpu):
    """The bare leaf: 100 001 frames of one stream, every sample against host/philox.py (libm vs CUDA log / sincospi:
    a few ulp)."""
    rng = PhiloxRNG(2 ** 63 + 12345, stream=7)                      # a seed with the top bit set
    y, fs = sink(Signal(randn, 44.1 * kHz, rng=rng) >> Until(100001 * (1 / 44100) * s), gpu)
    want = rng.frames(1, y.shape[0] + 1)
    assert y.shape == (100001, 1) and fs == 44100.0
    assert np.max(np.abs(y[:, 0] - want)) <= 1e-13
    assert gpu.last_stats["h2d_bytes"] == 0                          # nothing crosses the link


def scene(rng, fs=44.1 * kHz):
    """README scene with the noise drawn on the device (runtests.jl:896-918 shape)."""
    x = Signal(sin, ω=1 * kHz) >> Until(0.5 * s) >> Ramp() >> Normpower >> Amplify(-15 * dB)
    y = (Signal(randn, fs, rng=rng) >> After(0.01 * s) >> Until(0.5 * s) >> Filt(Bandstop, 0.5 * kHz, 2 * kHz) >> Normpower
         >> Amplify(-20 * dB))
    return Mix(x, y) >> ToFramerate(fs)


def test_scene_with_device_noise_matches_the_cpu_sink(gpu):
    x = scene(PhiloxRNG(1983))
    got, fs = sink(x, gpu)
    want, wfs = oracle.sink(x)
    assert got.shape == want.shape and fs == wfs
    assert np.max(np.abs(got - want)) <= TOL * rms(want)
    again, _ = sink(x, gpu)
    assert np.max(np.abs(got - again)) <= 1e-12 * rms(got)


def test_every_instance_of_a_batch_gets_its_own_stream(gpu):
    """70 graphs in 16 waves (and over every device of the context): instance k must see stream 5 + k whatever wave or
    device it lands on — the launches carry the index of their first instance."""
    xs = [scene(PhiloxRNG(77, stream=5 + k)) for k in range(70)]
    outs = sink_batch(xs, gpu)
    for k in (0, 1, 4, 5, 33, 69):
        want, _ = oracle.sink(xs[k])
        assert np.max(np.abs(outs[k][0] - want)) <= TOL * rms(want), k
    assert not np.allclose(outs[0][0], outs[1][0])
    old = os.environ.get("SIGOPS_HOST_WAVES")
    os.environ["SIGOPS_HOST_WAVES"] = "1"
    try:
        one = sink_batch(xs, gpu)
    finally:
        os.environ.pop("SIGOPS_HOST_WAVES")
        if old is not None:
            os.environ["SIGOPS_HOST_WAVES"] = old
    # (not bit for bit: the Normpower sums are accumulated with atomics, in a different order per launch shape)
    assert all(np.max(np.abs(a[0] - b[0])) <= 1e-12 * rms(a[0]) for a, b in zip(outs, one))
    import torch
    if torch.cuda.device_count() >= 2:
        multi = GPUSink(list(range(torch.cuda.device_count())))
        try:
            spread = sink_batch(xs, multi)
            assert all(np.max(np.abs(a[0] - b[0])) <= 1e-12 * rms(a[0]) for a, b in zip(outs, spread))
        finally:
            multi.close()


def test_noise_feeding_a_resampler_and_a_lowpass(gpu):
    """Noise as the input of IIR and FIR stages (materialised by the stage's own input program or a map stage)."""
    rng = PhiloxRNG(5)
    x = Signal(randn, 8 * kHz, rng=rng) >> Until(1 * s) >> Filt(Lowpass, 1 * kHz) >> ToFramerate(11.025 * kHz)
    got, fs = sink(x, gpu)
    want, _ = oracle.sink(x)
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) <= TOL * rms(want)
