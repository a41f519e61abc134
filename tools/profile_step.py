"""Run a few device-resident steps of one BASELINE workload (for ncu / quick timing).
usage: python tools/profile_step.py [cfg1|cfg2|cfg2f|cfg3|cfg3f|cfg3a|cfg4|cfg5] [steps] [ninst]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from signalops import (AffineSin, Amplify, Append, Bandpass, Bandstop, Filt, Lowpass, Mix, Normpower, Ramp, Sawtooth,  # noqa: E402
                       Signal, ToFramerate, Until, cabi, dB, Hz, kHz, ms, s, sin)
from signalops.lowering import lower  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
Z = np.zeros


def scene(fs, noise):
    x = Signal(sin, ω=1 * kHz) >> Until(1 * s) >> Ramp() >> Normpower >> Amplify(-20 * dB + 5 * dB)
    y = Signal(noise, fs) >> Until(1 * s) >> Filt(Bandstop, 0.5 * kHz, 2 * kHz) >> Normpower >> Amplify(-20 * dB)
    return Mix(x, y)


if cfg == "cfg1":
    ninst = 1
    g = scene(44.1 * kHz, Z(44100)) >> ToFramerate(44.1 * kHz)
elif cfg == "cfg2":
    ninst = 256
    g = Signal(Z((480000, 2)), 48 * kHz) >> Filt(Lowpass, 4 * kHz, order=8) >> Amplify(-20 * dB)
elif cfg == "cfg2f":      # config 2 on Float32 samples (half the bytes; state and arithmetic stay Float64)
    ninst = 256
    g = (Signal(np.zeros((480000, 2), dtype=np.float32), 48 * kHz) >> Filt(Lowpass, 4 * kHz, order=8)
         >> Amplify(np.float32(-20) * dB))
elif cfg == "cfg3":
    ninst = 64
    g = ToFramerate(Signal(Z((2646000, 2)), 44.1 * kHz), 48 * kHz)
elif cfg == "cfg3f":      # config 3 on Float32 samples: widen, tensor-core FIR, round
    ninst = 64
    g = ToFramerate(Signal(np.zeros((2646000, 2), dtype=np.float32), 44.1 * kHz), 48 * kHz)
elif cfg == "cfg3a":      # config 3 followed by a gain: FIR kernel + one elementwise pass
    ninst = 64
    g = ToFramerate(Signal(Z((2646000, 2)), 44.1 * kHz), 48 * kHz) >> Amplify(-6 * dB)
elif cfg == "cfg4":
    ninst = 512
    fs = 44.1 * kHz
    s1 = Signal(sin, ω=1 * kHz) >> Until(5 * s) >> Ramp() >> Normpower >> Amplify(-20 * dB)
    s2 = Signal(Z(88200), fs) >> Normpower >> Amplify(-20 * dB)
    s3 = Signal(Sawtooth(), ω=1 * kHz) >> Until(2 * s) >> Ramp() >> Normpower >> Amplify(-20 * dB)
    s4 = (Signal(Z(220500), fs) >> Amplify(Signal(AffineSin(0.5, 0.5), ω=5 * Hz)) >> Until(5 * s) >> Normpower
          >> Amplify(-20 * dB))
    g = Append(s1, s2, s3, s4, scene(fs, Z(44100))) >> Normpower >> Amplify(-20 * dB) >> ToFramerate(fs)
elif cfg == "cfg5":
    ninst = 8
    am = Amplify(Signal(Z((576000, 64)), 96 * kHz), Signal(AffineSin(0.5, 0.5), ω=5 * Hz)) >> Until(6 * s)
    g = am >> Filt(Bandpass, 500 * Hz, 4 * kHz) >> Ramp(10 * ms) >> Mix(Signal(sin, ω=1 * kHz) >> Until(6 * s))
elif cfg == "cfg5full":     # BASELINE config 5 at its real length: one-minute 64-channel signals, one wave of 16
    ninst = 16
    am = Amplify(Signal(Z((5760000, 64)), 96 * kHz), Signal(AffineSin(0.5, 0.5), ω=5 * Hz)) >> Until(60 * s)
    g = am >> Filt(Bandpass, 500 * Hz, 4 * kHz) >> Ramp(10 * ms) >> Mix(Signal(sin, ω=1 * kHz) >> Until(60 * s))
else:
    raise SystemExit("unknown cfg")
if len(sys.argv) > 3:
    ninst = int(sys.argv[3])
plan = lower(g)
ctx = cabi.Context([0])
cp = cabi.CompiledPlan(ctx, plan.tobytes())
gen = torch.Generator(device="cuda")
gen.manual_seed(1983)
tdt = lambda d: torch.float32 if d.dtype == cabi.F32 else torch.float64   # noqa: E731
xs = [torch.randn((ninst, d.nchannels, d.nframes), dtype=tdt(d), device="cuda", generator=gen) for d in plan.inputs]
ys = [torch.empty((ninst, d.nchannels, d.nframes), dtype=tdt(d), device="cuda") for d in plan.outputs]


def bufs(ts):
    arr = (cabi.Buffer * (ninst * len(ts)))()
    for i in range(ninst):
        for k, t in enumerate(ts):
            arr[i * len(ts) + k] = cabi.Buffer(t[i].data_ptr(), t.shape[2], t.shape[1], cabi.F32 if t.dtype == torch.float32 else cabi.F64, t.shape[2])
    return arr


ins, outs = bufs(xs), bufs(ys)
stream = torch.cuda.Stream()
torch.cuda.synchronize()
torch.cuda.set_stream(stream)
ctx.set_profiling(True)
for _ in range(2):
    cp.run_device(ninst, ins, outs, stream=stream.cuda_stream)
torch.cuda.synchronize()
ctx.profile_collect(0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
h0 = time.perf_counter()
for _ in range(steps):
    cp.run_device(ninst, ins, outs, stream=stream.cuda_stream)
host_ms = (time.perf_counter() - h0) * 1e3 / steps
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / steps
prof = ctx.profile_collect(0)
samples = ninst * sum(d.nchannels * d.nframes for d in plan.outputs)
alg = cp.algorithmic_bytes() * ninst
# the same steps without per-launch events: prepared waves replay as one CUDA graph
ctx.set_profiling(False)
for _ in range(3):
    cp.run_device(ninst, ins, outs, stream=stream.cuda_stream)
torch.cuda.synchronize()
e0.record()
for _ in range(steps):
    cp.run_device(ninst, ins, outs, stream=stream.cuda_stream)
e1.record()
torch.cuda.synchronize()
tg = e0.elapsed_time(e1) / steps
print(f"{cfg}: {tg:.3f} ms/step without per-launch events (graph replay), {samples / tg / 1e3:.0f} Msamples/s, {100 * alg / tg / 1e6 / 6552:.1f}% of 6552")
print(f"{cfg}: ninst={ninst} stages={len(plan.stages)} host enqueue {host_ms:.3f} ms/step; {t:.3f} ms/step, {samples / t / 1e3:.0f} Msamples/s, "
      f"alg bytes {alg / 1e9:.3f} GB -> {alg / t / 1e6:.0f} GB/s ({100 * alg / t / 1e6 / 6552:.1f}% of 6552); kernels/step "
      + ", ".join(f"{k}={v[0] / steps:.3f}ms/{v[1] // steps}" for k, v in prof.items()))
