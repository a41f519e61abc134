"""Diagnostic: the scalar FIR kernel under compute-sanitizer racecheck (timing perturbation) — are results still right?"""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from signalops import GPUSink, Hz, Signal, ToFramerate, sink_batch
gpu = GPUSink([0])
rng = np.random.default_rng(2999)
xs = [rng.standard_normal((2999, 1)) for _ in range(130)]
chain = lambda x: ToFramerate(Signal(x, 1000 * Hz), 1500 * Hz)
ref = None
for mode in sys.argv[1:] or ["tmap", "scalar", "scalar1"]:
    for k in ("SIGOPS_NO_FIR_TMAP", "SIGOPS_NO_FIR_MMA", "SIGOPS_HOST_WAVES"):
        os.environ.pop(k, None)
    if mode != "tmap": os.environ["SIGOPS_NO_FIR_TMAP"] = "1"
    if mode.startswith("scalar"): os.environ["SIGOPS_NO_FIR_MMA"] = "1"
    if mode.endswith("1"): os.environ["SIGOPS_HOST_WAVES"] = "1"
    out = sink_batch([chain(x) for x in xs], gpu)
    if ref is None: ref = [o[0].copy() for o in out]
    bad = [k for k, o in enumerate(out) if not np.allclose(o[0], ref[k], rtol=0, atol=1e-9)]
    print(mode, "instances differing from the first mode:", bad[:12], "of", len(bad))
