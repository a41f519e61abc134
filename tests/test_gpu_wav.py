"""The WAV step either side of the path: `sink(x, "file.wav")` and `Signal("file.wav")`,
src/sink.jl:139-142, src/WAV.jl:3-15.  The data chunk is produced / consumed ON THE DEVICE (csrc/k_wav.cuh:
transposition to frame-interleaved order + sample conversion); the host only handles the RIFF header.
Checked bit-exactly against scipy.io.wavfile round trips."""
import os

import numpy as np
import pytest
from scipy.io import wavfile

import oracle
from signalops import (Amplify, Filt, Hz, Lowpass, Mix, Signal, ToFramerate, WavFile, dB, kHz, sin, sink, sink_wav, Until, s)

pytestmark = pytest.mark.gpu


def rms(a):
    return float(np.sqrt(np.mean(np.asarray(a, dtype=np.float64) ** 2)))


@pytest.mark.parametrize("nch", [1, 2, 5, 40])
def test_sink_to_wav_float64_is_the_planar_result_bit_for_bit(gpu, tmp_path, nch):
    x = np.random.default_rng(nch).standard_normal((30011, nch)) * 0.2
    chain = lambda: Signal(x, 44.1 * kHz) >> Filt(Lowpass, 5 * kHz) >> Amplify(-3 * dB)   # noqa: E731
    path = str(tmp_path / "out.wav")
    fs = sink_wav(chain(), path, gpu)
    planar, fs2 = sink(chain(), gpu)
    got_fs, data = wavfile.read(path)
    assert got_fs == fs == 44100 and fs2 == 44100.0
    data = data.reshape(-1, nch)
    assert data.dtype == np.float64 and np.array_equal(data, planar)          # same kernels, only the layout differs
    want, _ = oracle.sink(chain())
    assert np.max(np.abs(data - want)) <= 1e-9 * rms(want)


def test_sink_to_wav_pcm16_and_float32(gpu, tmp_path):
    x = np.random.default_rng(7).standard_normal((20000, 2)) * 0.6             # some samples clip
    chain = lambda: Signal(x, 8 * kHz) >> Amplify(-2 * dB)                      # noqa: E731
    planar, _ = sink(chain(), gpu)
    p16 = str(tmp_path / "p16.wav")
    sink_wav(chain(), p16, gpu, encoding="pcm16")
    fs, d16 = wavfile.read(p16)
    assert fs == 8000 and d16.dtype == np.int16
    assert np.array_equal(d16, np.rint(np.clip(planar, -1, 1) * 32767).astype(np.int16))   # WAV.jl's PCM conversion
    p32 = str(tmp_path / "p32.wav")
    sink_wav(chain(), p32, gpu, encoding="float32")
    _, d32 = wavfile.read(p32)
    assert d32.dtype == np.float32 and np.array_equal(d32, planar.astype(np.float32))


@pytest.mark.parametrize("dt", [np.int16, np.float32, np.float64])
def test_signal_from_wav_file_is_decoded_on_the_device(gpu, tmp_path, dt):
    rng = np.random.default_rng(11)
    a = rng.standard_normal((25000, 2)) * 0.3
    raw = (a * 32767).astype(np.int16) if dt == np.int16 else a.astype(dt)
    path = str(tmp_path / "in.wav")
    wavfile.write(path, 48000, raw)
    decoded = raw.astype(np.float64) / 32768.0 if dt == np.int16 else raw.astype(np.float64)   # what wavread returns
    x = WavFile(path)
    assert x.framerate == 48000.0 and x.nframes == 25000 and x.nchannels == 2
    got, fs = sink(x >> Amplify(-6 * dB), gpu)
    assert fs == 48000.0 and np.array_equal(got, decoded * 10 ** (-6 / 20))
    got2, _ = sink(WavFile(path) >> Filt(Lowpass, 3 * kHz) >> Mix(Signal(sin, ω=1 * kHz) >> Until(25000 / 48000 * s)), gpu)
    want2, _ = oracle.sink(Signal(decoded, 48 * kHz) >> Filt(Lowpass, 3 * kHz) >> Mix(Signal(sin, ω=1 * kHz) >> Until(25000 / 48000 * s)))
    assert np.max(np.abs(got2 - want2)) <= 1e-9 * rms(want2)


def test_wav_round_trip_through_a_resampler(gpu, tmp_path):
    """README pipeline shape: file -> operators -> file (src/sink.jl:139-142 on both ends)."""
    a = np.random.default_rng(3).standard_normal((44100, 2)) * 0.25
    src, dst = str(tmp_path / "a.wav"), str(tmp_path / "b.wav")
    wavfile.write(src, 44100, a)
    fs = sink_wav(WavFile(src) >> ToFramerate(48 * kHz) >> Amplify(-6 * dB), dst, gpu)
    got_fs, data = wavfile.read(dst)
    want, _ = oracle.sink(Signal(a, 44.1 * kHz) >> ToFramerate(48 * kHz) >> Amplify(-6 * dB))
    assert got_fs == fs == 48000 and data.shape == want.shape == (48000, 2)
    assert np.max(np.abs(data - want)) <= 1e-9 * rms(want)
    with pytest.raises(Exception):
        WavFile(src, 48 * kHz)                      # frame-rate mismatch errors like src/WAV.jl:10-13
    assert os.path.getsize(dst) > 48000 * 2 * 8
