"""Seeded chains shared by the golden-vector generator and the tests that read them."""
import numpy as np

from signalops import (AffineSin, Amplify, Append, Bandpass, Bandstop, Chebyshev1, Filt, Highpass, Lowpass, Mix,
                       Normpower, Ramp, Sawtooth, Signal, ToFramerate, Until, dB, Hz, kHz, ms, s, sin)


def _rng(seed):
    return np.random.default_rng(seed)


def scene():
    """BASELINE config 1: README notch-noise scene at 44.1 kHz (1 s)."""
    noise = _rng(1983).standard_normal(44100)
    x = Signal(sin, ω=1 * kHz) >> Until(1 * s) >> Ramp() >> Normpower >> Amplify(-20 * dB + 5 * dB)
    y = Signal(noise) >> Until(1 * s) >> Filt(Bandstop, 0.5 * kHz, 2 * kHz) >> Normpower >> Amplify(-20 * dB)
    return Mix(x, y) >> ToFramerate(44.1 * kHz)


def iir_lowpass8():
    """BASELINE config 2, one short instance: 8th-order Butterworth low-pass + Amplify."""
    x = _rng(2).standard_normal((24000, 2))
    return Signal(x, 48 * kHz) >> Filt(Lowpass, 4 * kHz, order=8) >> Amplify(-20 * dB)


def resample_441_48():
    """BASELINE config 3, one short instance: 44.1 kHz -> 48 kHz (FIRArbitrary)."""
    x = _rng(3).standard_normal((22050, 2))
    return ToFramerate(Signal(x, 44.1 * kHz), 48 * kHz)


def readme_pipeline():
    """BASELINE config 4 scaled to 8 kHz: Append of five sounds |> Normpower |> Amplify."""
    fs = 8 * kHz
    s1 = Signal(sin, ω=1 * kHz) >> Until(1 * s) >> Ramp() >> Normpower >> Amplify(-20 * dB)
    s2 = Signal(_rng(4).standard_normal(4000), fs) >> Normpower >> Amplify(-20 * dB)
    s3 = Signal(Sawtooth(), ω=1 * kHz) >> Until(0.5 * s) >> Ramp() >> Normpower >> Amplify(-20 * dB)
    s4 = (Signal(_rng(5).standard_normal(8000), fs) >> Amplify(Signal(AffineSin(0.5, 0.5), ω=5 * Hz))
          >> Until(1 * s) >> Normpower >> Amplify(-20 * dB))
    x = Signal(sin, ω=1 * kHz) >> Until(0.25 * s) >> Ramp() >> Normpower >> Amplify(-15 * dB)
    y = (Signal(_rng(6).standard_normal(2000), fs) >> Filt(Bandstop, 0.5 * kHz, 2 * kHz) >> Normpower
         >> Amplify(-20 * dB))
    return Append(s1, s2, s3, s4, Mix(x, y)) >> Normpower >> Amplify(-20 * dB) >> ToFramerate(fs)


def am_bandpass_mix():
    """BASELINE config 5, small: AM noise |> Filt(Bandpass) |> Ramp |> Mix(tone), 4 channels."""
    x = _rng(7).standard_normal((9600, 4))
    am = Amplify(Signal(x, 96 * kHz), Signal(AffineSin(0.5, 0.5), ω=5 * Hz)) >> Until(100 * ms)
    return am >> Filt(Bandpass, 500 * Hz, 4 * kHz) >> Ramp(10 * ms) >> Mix(Signal(sin, ω=1 * kHz) >> Until(100 * ms))


def cheby_highpass():
    x = _rng(8).standard_normal((5000, 1))
    return Signal(x, 100 * Hz) >> Filt(Highpass, 8 * Hz, method=Chebyshev1(5, 1))


def resample_half():
    x = _rng(9).standard_normal((3000, 2))
    return ToFramerate(Signal(x, 1 * kHz), 500 * Hz)


CASES = {"scene": scene, "iir_lowpass8": iir_lowpass8, "resample_441_48": resample_441_48,
         "readme_pipeline": readme_pipeline, "am_bandpass_mix": am_bandpass_mix,
         "cheby_highpass": cheby_highpass, "resample_half": resample_half}
