"""The plan bytes committed for the Julia glue's byte-for-byte test (tests/golden/plans/*.bin,
julia/test_plans.jl) must be what host/lowering.py emits today, and must parse as the layout of
include/signalops.h (so that a format change cannot silently strand the Julia side)."""
import os
import struct
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden", "plans"))
import make_plans  # noqa: E402


@pytest.mark.parametrize("name", sorted(make_plans.CASES))
def test_committed_plan_bytes_are_current(name):
    with open(os.path.join(HERE, "golden", "plans", name + ".bin"), "rb") as f:
        want = f.read()
    got = make_plans.plan_bytes(name)
    assert got == want, (f"{name}: lowering.py no longer emits the committed plan; rerun tests/golden/plans/make_plans.py "
                         "and update julia/test_plans.jl if the format changed")


@pytest.mark.parametrize("name", sorted(make_plans.CASES))
def test_plan_sections_add_up(name):
    b = make_plans.plan_bytes(name)
    magic, ver, n_in, n_tmp, n_out, n_scal, n_tab, n_instr, n_piece, n_stage, n_dbl = struct.unpack_from("<10IQ", b, 0)
    assert magic == 0x504F4753 and ver == 1
    assert len(b) == 48 + 16 * (n_in + n_tmp + n_out) + 16 * n_tab + 80 * n_instr + 32 * n_piece + 128 * n_stage + 8 * n_dbl
