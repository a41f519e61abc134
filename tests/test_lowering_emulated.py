"""CPU check of the host-side lowering: every chain of test_gpu_parity.py is lowered
to plan bytes, the bytes are evaluated by tests/plan_emulator.py (numpy), and the
result must match the oracle exactly as the CUDA path must.  This exercises the
graph layer, the planner and the byte format without a GPU; the CUDA kernels
themselves are covered by `-m gpu`."""
import numpy as np
import pytest

import test_gpu_parity as T
from plan_emulator import Emulator
from signalops import GPUSink


class _EmuPlan:
    def __init__(self, blob):
        self.blob = blob

    def run_host(self, ninst, ins, outs):
        n_in = len(ins) // ninst
        n_out = len(outs) // ninst
        for i in range(ninst):
            res = Emulator(self.blob).run(ins[i * n_in:(i + 1) * n_in], inst=i)
            for o, r in zip(outs[i * n_out:(i + 1) * n_out], res):
                o = o.raw if hasattr(o, "raw") else o          # (WavRaw: a C-ordered result written in place)
                o.reshape(r.shape)[...] = r
        return {"launches": 0}


class EmulatedSink(GPUSink):
    def compiled(self, plan_bytes):
        return _EmuPlan(plan_bytes)

    def alloc_result(self, shape, dtype):          # no CUDA driver here: plain pageable memory
        return np.empty(shape, dtype=dtype, order="F")


@pytest.fixture(scope="module")
def gpu():
    return EmulatedSink()


for _name in dir(T):
    if _name.startswith("test_"):
        globals()[_name] = getattr(T, _name)
