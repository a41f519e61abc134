"""Run a few device-resident steps of one bench workload (for ncu / quick timing).
usage: python tools/profile_step.py [cfg2|cfg3|cfg5] [steps] [ninst]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from signalops import (AffineSin, Amplify, Bandpass, Filt, Lowpass, Mix, Ramp, Signal, ToFramerate, Until, cabi,  # noqa: E402
                       dB, Hz, kHz, ms, s, sin)
from signalops.lowering import lower  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if cfg == "cfg2":
    nin, nch, ninst = 480000, 2, 256
    g = Signal(np.zeros((nin, nch)), 48 * kHz) >> Filt(Lowpass, 4 * kHz, order=8) >> Amplify(-20 * dB)
elif cfg == "cfg3":
    nin, nch, ninst = 2646000, 2, 64
    g = ToFramerate(Signal(np.zeros((nin, nch)), 44.1 * kHz), 48 * kHz)
elif cfg == "cfg5":
    nin, nch, ninst = 576000, 64, 8
    am = Amplify(Signal(np.zeros((nin, nch)), 96 * kHz), Signal(AffineSin(0.5, 0.5), ω=5 * Hz)) >> Until(6 * s)
    g = am >> Filt(Bandpass, 500 * Hz, 4 * kHz) >> Ramp(10 * ms) >> Mix(Signal(sin, ω=1 * kHz) >> Until(6 * s))
else:
    raise SystemExit("unknown cfg")
if len(sys.argv) > 3:
    ninst = int(sys.argv[3])
plan = lower(g)
nout = plan.outputs[0].nframes
ctx = cabi.Context([0])
cp = cabi.CompiledPlan(ctx, plan.tobytes())
x = torch.randn((ninst, nch, nin), dtype=torch.float64, device="cuda")
y = torch.empty((ninst, nch, nout), dtype=torch.float64, device="cuda")
ins = (cabi.Buffer * ninst)(*[cabi.Buffer(x[i].data_ptr(), nin, nch, cabi.F64, nin) for i in range(ninst)])
outs = (cabi.Buffer * ninst)(*[cabi.Buffer(y[i].data_ptr(), nout, nch, cabi.F64, nout) for i in range(ninst)])
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
samples = ninst * nch * nout
print(f"{cfg}: host enqueue {host_ms:.3f} ms/step; {t:.3f} ms/step, {samples / t / 1e3:.0f} Msamples/s, alg bytes {cp.algorithmic_bytes() * ninst / 1e9:.3f} GB "
      f"-> {cp.algorithmic_bytes() * ninst / t / 1e6:.0f} GB/s; kernels/step "
      + ", ".join(f"{k}={v[0] / steps:.3f}ms" for k, v in prof.items()))
