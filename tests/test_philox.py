"""Device noise for `Signal(randn; rng)` (src/functions.jl:98-114): the counter-based generator behind LEAF_RANDN.
CPU side: known-answer vectors of Philox4x32-10, the statistics of the Box-Muller output, and the lowering /
plan-interpreter / oracle agreement.  The CUDA implementation is compared with the same numpy function in
tests/test_gpu_randn.py."""
import numpy as np
import pytest

import oracle.cpu_sink as oracle
from plan_emulator import Emulator
from signalops import (After, Amplify, Bandstop, Filt, Hz, Mix, Normpower, PhiloxRNG, Ramp, Signal, Until, dB, kHz,
                       randn, s, sin, sink_batch)
from signalops.lowering import Lowerer, lower
from signalops.philox import philox4x32_10


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32 10 rounds."""
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        got = philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert [int(v) for v in got] == want
    # vectorised over a leading axis
    got = philox4x32_10(np.array([k[0] for k in kat], dtype=np.uint32), np.array([k[1] for k in kat], dtype=np.uint32))
    assert got.tolist() == [k[2] for k in kat]


def test_noise_is_standard_normal_and_streams_are_independent():
    a = PhiloxRNG(1983).frames(1, 400001)
    b = PhiloxRNG(1983, stream=1).frames(1, 400001)
    c = PhiloxRNG(1984).frames(1, 400001)
    for x in (a, b, c):
        assert abs(x.mean()) < 5e-3 and abs(x.var() - 1) < 1e-2
        assert abs(np.mean(x ** 3)) < 2e-2 and abs(np.mean(x ** 4) - 3) < 5e-2
        assert abs(np.corrcoef(x[:-1], x[1:])[0, 1]) < 5e-3          # cos / sin halves of a pair included
    assert abs(np.corrcoef(a, b)[0, 1]) < 5e-3 and abs(np.corrcoef(a, c)[0, 1]) < 5e-3
    assert np.all(np.isfinite(a)) and np.abs(a).max() < 9


def test_frames_are_a_function_of_the_frame_index():
    r = PhiloxRNG(7, stream=5)
    x = r.frames(1, 1001)
    assert np.array_equal(r.frames(338, 700), x[337:699])
    assert np.array_equal(r.standard_normal(100), x[:100])
    assert r.frames(10, 10).size == 0
    big = PhiloxRNG(2 ** 64 - 3, stream=2 ** 40 + 1).frames(2 ** 33 + 1, 2 ** 33 + 9)     # 64-bit seed / stream / counter
    assert np.all(np.isfinite(big)) and len(set(big.tolist())) == 8


def chain(rng, fs=8 * kHz):
    noise = Signal(randn, fs, rng=rng) >> After(0.01 * s) >> Until(0.25 * s) >> Filt(Bandstop, 0.5 * kHz, 2 * kHz)
    return Mix(noise >> Normpower >> Amplify(-20 * dB), Signal(sin, ω=1 * kHz) >> Until(0.25 * s) >> Ramp() >> Normpower)


def test_device_noise_lowers_to_a_leaf_and_matches_the_oracle():
    x = chain(PhiloxRNG(42, stream=3))
    plan = lower(x)
    assert plan.input_arrays == []                       # nothing crosses the link: the noise is a leaf
    got = Emulator(plan.tobytes()).run([], inst=0)[0]
    want, fs = oracle.sink(x)
    assert got.shape == want.shape and fs == 8000.0
    assert np.max(np.abs(got - want)) <= 1e-12 * np.sqrt(np.mean(want ** 2))
    # a numpy generator still takes the host-materialised path
    y = Signal(randn, 8 * kHz, rng=np.random.default_rng(1)) >> Until(0.1 * s)
    assert len(lower(y).input_arrays) == 1


def test_batch_streams_lower_to_one_plan():
    xs = [chain(PhiloxRNG(42, stream=10 + k)) for k in range(4)]
    blobs = [Lowerer(instance_index=k).build(x).tobytes() for k, x in enumerate(xs)]
    assert all(b == blobs[0] for b in blobs)
    for k in (0, 3):                                     # instance k of the plan = stream 10 + k
        got = Emulator(blobs[0]).run([], inst=k)[0]
        want, _ = oracle.sink(xs[k])
        assert np.max(np.abs(got - want)) <= 1e-12 * np.sqrt(np.mean(want ** 2))
    from test_lowering_emulated import EmulatedSink
    with pytest.raises(Exception, match="same plan"):
        sink_batch([chain(PhiloxRNG(42, stream=0)), chain(PhiloxRNG(42, stream=5))], EmulatedSink())
    outs = sink_batch(xs, EmulatedSink())
    assert not np.allclose(outs[0][0], outs[1][0])
