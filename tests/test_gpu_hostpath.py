"""The host-buffer path of the C ABI (`sigops_plan_run`): page-locked caller arrays are read and written by
the DMA engines directly, pageable ones (what a Julia `Array` or a numpy array is) are staged through the
library's pinned ring by copy threads; three staging buffers keep H2D, kernels and D2H in flight at once.
Also: replay of a prepared wave as a CUDA graph, calls on several streams sharing one workspace, and the
in-process sharding of a batch over all the devices of one context (`sigops_ctx_create(devices, n)`).

Reference behaviour being replaced: the `sink!` block loop, src/sink.jl:158-168,225-267."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from signalops import (Amplify, Bandstop, Filt, GPUSink, Hz, Lowpass, Mix, Normpower, Ramp, Signal, ToFramerate, Until,
                       cabi, dB, kHz, s, sin, sink, sink_batch)
from signalops.lowering import lower

pytestmark = pytest.mark.gpu
TOL = 1e-9


def rms(a):
    return float(np.sqrt(np.mean(np.asarray(a, dtype=np.float64) ** 2)))


def chain(x):
    return Signal(x, 48 * kHz) >> Filt(Lowpass, 4 * kHz, order=8) >> Amplify(-20 * dB)


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


def test_pageable_and_pinned_callers_agree(gpu):
    rng = np.random.default_rng(1)
    n, nch, ninst = 48000, 2, 40
    xs = [np.asfortranarray(rng.standard_normal((n, nch))) for _ in range(ninst)]
    plan = lower(chain(xs[0]))
    cp = gpu.compiled(plan.tobytes())
    # pageable in, pageable out
    ys_pg = [np.zeros((n, nch), order="F") for _ in range(ninst)]
    st = cp.run_host(ninst, xs, ys_pg)
    assert st["h2d_bytes"] == ninst * n * nch * 8 and st["d2h_bytes"] == ninst * n * nch * 8
    # pinned in, pinned out
    xp = [cabi.pinned_empty((n, nch), np.float64) for _ in range(ninst)]
    for a, b in zip(xp, xs):
        a[...] = b
    yp = [cabi.pinned_empty((n, nch), np.float64) for _ in range(ninst)]
    cp.run_host(ninst, xp, yp)
    for k in range(ninst):
        assert np.array_equal(ys_pg[k], yp[k])
    for k in (0, 17, ninst - 1):
        want, _ = oracle.sink(chain(xs[k]))
        assert np.max(np.abs(yp[k] - want)) <= TOL * rms(want)
    # mixed: pinned in, pageable out — and every wave through the staging ring
    ys2 = [np.zeros((n, nch), order="F") for _ in range(ninst)]
    with_env({"SIGOPS_FORCE_STAGING": "1", "SIGOPS_HOST_WAVES": "7"}, lambda: cp.run_host(ninst, xp, ys2))
    for k in range(ninst):
        assert np.array_equal(ys2[k], yp[k])


def test_strided_channels_through_the_staging_ring(gpu):
    """ld > nframes on the host side (a view of a taller matrix): rows are gathered channel by channel."""
    rng = np.random.default_rng(2)
    big = np.asfortranarray(rng.standard_normal((30000, 2)))
    x = big[:24000]                                   # column-major view: ld = 30000
    outbig = np.full((26000, 2), 7.0, order="F")
    y = outbig[:24000]
    cp = gpu.compiled(lower(chain(np.zeros((24000, 2)))).tobytes())
    cp.run_host(1, [x], [y])
    want, _ = oracle.sink(chain(np.ascontiguousarray(x)))
    assert np.max(np.abs(y - want)) <= TOL * rms(want)
    assert np.all(outbig[24000:] == 7.0)


def test_sink_can_return_page_locked_results():
    """GPUSink(pin_results=True): results live in sigops_host_alloc memory, recycled through a pool."""
    pinned = GPUSink([0], pin_results=True)
    try:
        x = np.random.default_rng(3).standard_normal((300000, 2))
        y, fs = sink(chain(x), pinned)
        assert y.flags.f_contiguous and y.shape == (300000, 2)
        want, _ = oracle.sink(chain(x))
        assert np.max(np.abs(y - want)) <= TOL * rms(want)
        addr = y.ctypes.data
        del y
        y2, _ = sink(chain(x), pinned)                  # the block comes back from the pool
        assert y2.ctypes.data == addr and np.max(np.abs(y2 - want)) <= TOL * rms(want)
    finally:
        pinned.close()


def test_replay_as_cuda_graph_matches_eager(gpu):
    """Same plan on the same device buffers four times: from the third run on the wave is one cudaGraphLaunch."""
    import torch
    noise = np.random.default_rng(4).standard_normal(44100)

    def scene(v):
        a = Signal(sin, ω=1 * kHz) >> Until(1 * s) >> Ramp() >> Normpower >> Amplify(-15 * dB)
        b = Signal(v, 44.1 * kHz) >> Until(1 * s) >> Filt(Bandstop, 0.5 * kHz, 2 * kHz) >> Normpower >> Amplify(-20 * dB)
        return Mix(a, b)

    plan = lower(scene(noise))
    cp = gpu.compiled(plan.tobytes())
    x = torch.from_numpy(noise).cuda().reshape(1, 1, -1).contiguous()
    y = torch.zeros((1, 1, 44100), dtype=torch.float64, device="cuda")
    ins = (cabi.Buffer * 1)(cabi.Buffer(x.data_ptr(), 44100, 1, cabi.F64, 44100))
    outs = (cabi.Buffer * 1)(cabi.Buffer(y.data_ptr(), 44100, 1, cabi.F64, 44100))
    stream = torch.cuda.Stream()
    res = []
    for _ in range(5):
        y.zero_()
        torch.cuda.synchronize()
        cp.run_device(1, ins, outs, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        res.append(y.cpu().numpy().copy())
    want, _ = oracle.sink(scene(noise))
    # (runs agree to rounding, not bit for bit: the Normpower sums of squares are accumulated with atomics)
    for r in res:
        assert np.max(np.abs(r[0, 0] - want[:, 0])) <= TOL * rms(want)
        assert np.max(np.abs(r - res[0])) <= 1e-12 * rms(want)
    # ... and eagerly, with graphs switched off
    y.zero_()
    with_env({"SIGOPS_NO_GRAPH": "1"}, lambda: cp.run_device(1, ins, outs, stream=stream.cuda_stream))
    torch.cuda.synchronize()
    assert np.max(np.abs(y.cpu().numpy() - res[0])) <= 1e-12 * rms(want)


def test_calls_on_two_streams_share_the_workspace_safely(gpu):
    """Asynchronous run_device calls on different streams: the library orders them."""
    import torch
    rng = np.random.default_rng(5)
    n, nch, ninst = 96000, 2, 16
    plan = lower(chain(np.zeros((n, nch))))
    cp = gpu.compiled(plan.tobytes())
    xa = torch.from_numpy(rng.standard_normal((ninst, nch, n))).cuda()
    xb = torch.from_numpy(rng.standard_normal((ninst, nch, n))).cuda()
    ya, yb = torch.empty_like(xa), torch.empty_like(xb)

    def bufs(t):
        arr = (cabi.Buffer * ninst)()
        for i in range(ninst):
            arr[i] = cabi.Buffer(t[i].data_ptr(), n, nch, cabi.F64, n)
        return arr
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(3):
        cp.run_device(ninst, bufs(xa), bufs(ya), stream=s1.cuda_stream)
        cp.run_device(ninst, bufs(xb), bufs(yb), stream=s2.cuda_stream)
    torch.cuda.synchronize()
    for x, y in ((xa, ya), (xb, yb)):
        xi = x[3].cpu().numpy().T
        want, _ = oracle.sink(chain(xi))
        assert np.max(np.abs(y[3].cpu().numpy().T - want)) <= TOL * rms(want)


def test_batch_sharded_over_every_device_of_one_context():
    """north_star: "independent signals are sharded across the 8 B200s of one box" — ONE process, ONE context."""
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs at least two GPUs (run under gpurun --gpus N)")
    multi = GPUSink(list(range(ndev)))
    single = GPUSink([0])
    try:
        rng = np.random.default_rng(6)
        ninst = 8 * ndev + 3                            # ragged split
        xs = [rng.standard_normal((44100, 2)) for _ in range(ninst)]
        mk = lambda x: ToFramerate(Signal(x, 44.1 * kHz), 48 * kHz) >> Amplify(-6 * dB)   # noqa: E731
        got = sink_batch([mk(x) for x in xs], multi)
        st = dict(multi.last_stats)
        ref = sink_batch([mk(x) for x in xs], single)
        for (a, fa), (b, fb) in zip(got, ref):
            assert fa == fb == 48000.0 and np.array_equal(a, b)
        want, _ = oracle.sink(mk(xs[-1]))
        assert np.max(np.abs(got[-1][0] - want)) <= TOL * rms(want)
        assert st["out_samples"] == ninst * 48000 * 2
        # IIR path too, with Normpower (joint over the channels of one signal: channels never split across devices)
        mk2 = lambda x: Signal(x, 44.1 * kHz) >> Filt(Lowpass, 3 * kHz) >> Normpower   # noqa: E731
        got2 = sink_batch([mk2(x) for x in xs], multi)
        for k in (0, ninst // 2, ninst - 1):
            w2, _ = oracle.sink(mk2(xs[k]))
            assert np.max(np.abs(got2[k][0] - w2)) <= TOL * rms(w2)
    finally:
        multi.close()
        single.close()


@pytest.mark.parametrize("env", [{}, {"SIGOPS_NO_FIR_TMAP": "1"}, {"SIGOPS_NO_FIR_TMAP": "1", "SIGOPS_NO_FIR_MMA": "1"}],
                         ids=["tensor-map", "per-row TMA", "scalar"])
def test_every_output_sample_is_written(gpu, env):
    """Caller arrays poisoned with NaN before each call, many waves, several repetitions: a write that is skipped or
    overtaken somewhere in the H2D | kernels | D2H pipeline cannot hide behind a stale but plausible value."""
    rng = np.random.default_rng(2999)
    ninst, n = 130, 2999
    xs = [rng.standard_normal((n, 1)) for _ in range(ninst)]
    mk = lambda x: ToFramerate(Signal(x, 1000 * Hz), 1500 * Hz)   # noqa: E731
    plan = lower(mk(xs[0]))
    nout = plan.outputs[0].nframes
    cp = gpu.compiled(plan.tobytes())
    first = None
    for rep in range(4):
        ys = [np.full((nout, 1), np.nan, order="F") for _ in range(ninst)]
        with_env(env, lambda: cp.run_host(ninst, xs, ys))
        assert not any(np.isnan(y).any() for y in ys), f"repetition {rep}: unwritten output samples"
        if first is None:
            first = ys
        else:
            assert all(np.array_equal(a, b) for a, b in zip(first, ys))
    want, _ = oracle.sink(mk(xs[77]))
    assert np.max(np.abs(first[77] - want)) <= TOL * rms(want)
