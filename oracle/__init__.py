"""CPU oracle — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A restatement of the reference's CPU `sink` (block-pull interpreter, src/sink.jl)
and of the DSP.jl kernels it calls, used to check the CUDA path.  Only tests/,
`__graft_entry__.smoke()` and bench.py's CPU baseline may import this package;
nothing under `signaloperators.jl_b200/` does.

PARITY UNPINNED for the absolute output of `Filt` / `ToFramerate`: those kernels
live in DSP.jl 0.6.10 (not vendored, no Julia here) and the reference's own tests
hold no golden vectors for them.  Everything outside DSP.jl is
pinned by the exact assertions of test/runtests.jl ported in tests/.
"""
from .cpu_sink import sink, sink_into  # noqa: F401
