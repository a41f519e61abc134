"""The C-ABI library: it loads, exports every symbol include/signalops.h declares, its
struct layouts match the Python packers, and it fails loudly (no CPU fallback) when no
GPU is present.  No compute calls here — those are the `-m gpu` tests."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest

from signalops import cabi
from signalops.lowering import Instr, lower
from signalops import Signal, Filt, Lowpass, Amplify, dB, Hz, kHz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "signalops.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sigops_[a-z_]+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(cabi.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.dirname(cabi.LIB_PATH).startswith(ROOT)


def test_exports_every_declared_symbol():
    lib = ctypes.CDLL(cabi.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/signalops.h but not exported"
    assert sorted(cabi.SYMBOLS) == names


def test_abi_version_and_struct_sizes():
    lib = cabi.load()
    assert lib.sigops_abi_version() == 1
    assert ctypes.sizeof(cabi.Buffer) == 32
    assert ctypes.sizeof(cabi.Stats) == 64
    assert len(Instr().pack()) == 80
    assert struct.calcsize("<2q4i") == 32 and struct.calcsize("<10i2q2id8iq2d") == 128
    assert struct.calcsize("<10IQ") == 48


def test_sm100a_code_is_embedded():
    """The shared library carries sm_100a SASS for the hand-written kernels."""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_fails_loudly_without_a_gpu_or_with_garbage():
    lib = cabi.load()
    n = ctypes.c_int(-1)
    assert lib.sigops_device_count(ctypes.byref(n)) == 0
    if n.value == 0:
        with pytest.raises(cabi.SigopsError) as e:
            cabi.Context([0])
        assert "no CPU fallback" in str(e.value)
    # NULL arguments are errors, never crashes
    assert lib.sigops_plan_create(None, None, 0, None) != 0
    assert lib.sigops_plan_run(None, 0, None, None, None) != 0
    assert lib.sigops_last_error(None) is not None


def test_plan_bytes_are_deterministic():
    x = np.zeros((1000, 2))
    mk = lambda: lower(Signal(x, 48 * kHz) >> Filt(Lowpass, 4 * kHz, order=8) >> Amplify(-20 * dB)).tobytes()
    a, b = mk(), mk()
    assert a == b and a[:4] == b"SGOP"
