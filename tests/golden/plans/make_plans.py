"""Plan-byte fixtures for the Julia glue (signaloperators.jl_b200/julia/GPUSink.jl).

The image has no Julia, so the general lowering of GPUSink.jl cannot be run here.  What can be pinned is its
OUTPUT: for each graph below, host/lowering.py (the executable specification GPUSink.jl transcribes) emits the
plan bytes committed as <name>.bin; julia/test_plans.jl builds the same graph with the reference's own API,
lowers it with GPUSinks.lower and compares byte for byte (coefficient tables and the three Float64 stage fields
that come out of DSP.jl are compared to 1e-12).  tests/test_plan_fixtures.py keeps the Python side honest: the
committed bytes must equal what lowering.py emits today.

usage: python tests/golden/plans/make_plans.py   (rewrites tests/golden/plans/*.bin)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(HERE))))

from signalops import (AddChannel, AffineSin, After, Amplify, Append, Bandpass, Bandstop, Filt, Hz, Lowpass, Mix,  # noqa: E402
                       Normpower, Pad, PhiloxRNG, Ramp, RampOn, Sawtooth, SelectChannel, Signal, ToChannels, ToFramerate,
                       Until, cycle, dB, frames, kHz, ms, randn, s, sin, zero)
from signalops.lowering import lower  # noqa: E402

Z = np.zeros


def cfg1():
    x = Signal(sin, ω=1 * kHz) >> Until(1 * s) >> Ramp() >> Normpower >> Amplify(-20 * dB + 5 * dB)
    y = Signal(Z(44100), 44.1 * kHz) >> Until(1 * s) >> Filt(Bandstop, 0.5 * kHz, 2 * kHz) >> Normpower >> Amplify(-20 * dB)
    return Mix(x, y) >> ToFramerate(44.1 * kHz)


def cfg2():
    return Signal(Z((480000, 2)), 48 * kHz) >> Filt(Lowpass, 4 * kHz, order=8) >> Amplify(-20 * dB)


def cfg3():
    return ToFramerate(Signal(Z((2646000, 2)), 44.1 * kHz), 48 * kHz)


def cfg3_gain():
    return ToFramerate(Signal(Z((2646000, 2)), 44.1 * kHz), 48 * kHz) >> Amplify(-6 * dB)


def cfg4():
    fs = 44.1 * kHz
    s1 = Signal(sin, ω=1 * kHz) >> Until(5 * s) >> Ramp() >> Normpower >> Amplify(-20 * dB)
    s2 = Signal(Z(88200), fs) >> Normpower >> Amplify(-20 * dB)
    s3 = Signal(Sawtooth(), ω=1 * kHz) >> Until(2 * s) >> Ramp() >> Normpower >> Amplify(-20 * dB)
    s4 = Signal(Z(220500), fs) >> Amplify(Signal(AffineSin(0.5, 0.5), ω=5 * Hz)) >> Until(5 * s) >> Normpower >> Amplify(-20 * dB)
    x = Signal(sin, ω=1 * kHz) >> Until(1 * s) >> Ramp() >> Normpower >> Amplify(-20 * dB + 5 * dB)
    y = Signal(Z(44100), fs) >> Until(1 * s) >> Filt(Bandstop, 0.5 * kHz, 2 * kHz) >> Normpower >> Amplify(-20 * dB)
    return Append(s1, s2, s3, s4, Mix(x, y)) >> Normpower >> Amplify(-20 * dB) >> ToFramerate(fs)


def cfg5():
    am = Amplify(Signal(Z((576000, 4)), 96 * kHz), Signal(AffineSin(0.5, 0.5), ω=5 * Hz)) >> Until(6 * s)
    return am >> Filt(Bandpass, 500 * Hz, 4 * kHz) >> Ramp(10 * ms) >> Mix(Signal(sin, ω=1 * kHz) >> Until(6 * s))


def plumbing():
    a = Signal(Z((100, 2)), 10 * Hz)
    b = Signal(Z((40, 2)), 10 * Hz)
    x = a >> After(2 * s) >> Append(b >> Pad(zero) >> Until(60 * frames)) >> RampOn(5 * frames)
    return Mix(x, Signal(Z((30, 2)), 10 * Hz) >> Pad(cycle) >> Until(140 * frames)) >> Amplify(0.5)


def channels():
    a = Signal(Z((50, 3)), 10 * Hz)
    return AddChannel(a >> SelectChannel(2), a >> ToChannels(1)) >> ToChannels(2) >> Amplify(2)


def noise():
    """README scene with the noise drawn on the device (LEAF_RANDN): seed with the top bit set, stream 3, skipped frames."""
    x = Signal(sin, ω=1 * kHz) >> Until(1 * s) >> Ramp() >> Normpower >> Amplify(-20 * dB + 5 * dB)
    y = (Signal(randn, 44.1 * kHz, rng=PhiloxRNG(2 ** 63 + 1983, 3)) >> After(0.5 * s) >> Until(1 * s)
         >> Filt(Bandstop, 0.5 * kHz, 2 * kHz) >> Normpower >> Amplify(-20 * dB))
    return Mix(x, y)


CASES = {"noise": noise, "cfg1": cfg1, "cfg2": cfg2, "cfg3": cfg3, "cfg3_gain": cfg3_gain, "cfg4": cfg4, "cfg5": cfg5,
         "plumbing": plumbing, "channels": channels}


def plan_bytes(name):
    return lower(CASES[name]()).tobytes()


if __name__ == "__main__":
    for name in CASES:
        with open(os.path.join(HERE, name + ".bin"), "wb") as f:
            f.write(plan_bytes(name))
        print(name, os.path.getsize(os.path.join(HERE, name + ".bin")), "bytes")
