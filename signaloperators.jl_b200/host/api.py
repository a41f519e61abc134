"""Public names — the reference's exports plus GPUSink."""
from .dspjl import (Bandpass, Bandstop, Biquad, Butterworth, Chebyshev1, Highpass,
                    Lowpass, PolynomialRatio, SecondOrderSections, ZeroPoleGain,
                    digitalfilter)
from .functors import AffineCos, AffineSin, Sawtooth
from .gpusink import Array, GPUSink, Tuple, sink, sink_batch, sink_into, sink_wav
from .wav import WavFile, WavRaw, WavSignal, read_wav, write_wav
from .graph import (AddChannel, After, Amplify, Append, Extend, FadeTo, Filt, Format, Functor,
                    Mix, Normpower, Operate, OperateOn, Pad, Prepend, Ramp, RampOff, RampOn,
                    SelectChannel, Signal, SignalError, ToChannels, ToEltype, ToFramerate,
                    Uniform, Until, Window, cos, cycle, duration, framerate, identity, inflen,
                    lastframe, mirror, nchannels, nframes, one, randn, reverse, sampletype, sin,
                    sinramp, zero)
from .lowering import LoweringError
from .philox import PhiloxRNG
from .units import Hz, dB, deg, frames, kframes, kHz, ms, rad, s

__all__ = [n for n in dir() if not n.startswith("_")]
