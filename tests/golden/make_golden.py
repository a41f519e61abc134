"""Regenerates tests/golden/*.npz from the CPU oracle (the reference ships no golden
vectors for Filt/ToFramerate and cannot run here: SURVEY.md §8c).  Inputs are seeded;
run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from golden_cases import CASES  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))
for name, make in CASES.items():
    out = oracle.sink(make())
    data, fs = out if isinstance(out, tuple) else (out, np.nan)
    np.savez_compressed(os.path.join(here, name + ".npz"), data=data, fs=np.float64(fs if fs is not None else np.nan))
    print(name, data.shape, data.dtype, fs)
