"""The WAV step either side of the path: `sink(x, "file.wav")` and `Signal("file.wav")`
(src/sink.jl:139-142, src/WAV.jl:3-15: `wavwrite(data, file, Fs=round(Int,fs))` / `wavread`).

The host only handles the RIFF container.  The sample data never takes the reference's detour through a
planar Float64 matrix on the host: a data chunk is frame-interleaved in the file's sample encoding, and the
library transposes / converts it on the device (csrc/k_wav.cuh, `SIGOPS_INTERLEAVED` host buffers of
include/signalops.h) — results leave the GPU in file layout, files enter it as they are on disk."""
from __future__ import annotations

import struct

import numpy as np

from . import graph as G

F32, F64, I16 = 1, 2, 4
INTERLEAVED = 0x100
_ENC = {np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.int16): I16}


class WavRaw:
    """(nframes, nchannels) C-ordered array in a file's sample encoding = the bytes of a WAV data chunk."""

    def __init__(self, raw):
        raw = np.ascontiguousarray(raw)
        if raw.ndim == 1:
            raw = raw.reshape(-1, 1)
        if raw.dtype not in _ENC:
            raise G.SignalError(f"WAV sample encoding {raw.dtype} is not supported (PCM16, Float32, Float64)")
        self.raw = raw
        self.enc = _ENC[raw.dtype]

    shape = property(lambda self: self.raw.shape)

    def decode(self):
        """What WAV.jl `wavread` returns (Float64; PCM16 / 32768) — host-side, for tests."""
        return self.raw.astype(np.float64) / 32768.0 if self.enc == I16 else self.raw.astype(np.float64)


class WavSignal(G.AbstractSignal):
    """`Signal("file.wav")` (src/WAV.jl:8-15): a data signal whose samples stay in file layout until they are
    on the device."""
    evaltrait = "data"

    def __init__(self, wav: WavRaw, fs):
        self.wav = wav
        self._fs = float(fs)

    framerate = property(lambda self: self._fs)
    nchannels = property(lambda self: self.wav.shape[1])
    sampletype = property(lambda self: np.dtype(np.float64))        # wavread's default format="double"

    def nframes_helper(self):
        return self.wav.shape[0]


def write_wav(path, raw, fs):
    """RIFF/WAVE container around an interleaved (nframes, nchannels) array of int16 / float32 / float64."""
    raw = np.ascontiguousarray(raw)
    n, c = raw.shape
    isfloat = raw.dtype.kind == "f"
    fmt = struct.pack("<HHIIHH", 3 if isfloat else 1, c, int(fs), int(fs) * c * raw.itemsize, c * raw.itemsize, 8 * raw.itemsize)
    if isfloat:
        fmt += b"\x00\x00"                              # cbSize of a non-PCM format
    chunks = b"fmt " + struct.pack("<I", len(fmt)) + fmt
    if isfloat:
        chunks += b"fact" + struct.pack("<II", 4, n)
    data_bytes = raw.nbytes
    chunks += b"data" + struct.pack("<I", data_bytes)
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 4 + len(chunks) + data_bytes + (data_bytes & 1)) + b"WAVE" + chunks)
        f.write(raw.tobytes())
        if data_bytes & 1:
            f.write(b"\x00")


def read_wav(path):
    """-> (WavRaw, framerate).  PCM16 and IEEE float 32/64, plain or WAVE_FORMAT_EXTENSIBLE."""
    with open(path, "rb") as f:
        b = f.read()
    if b[:4] != b"RIFF" or b[8:12] != b"WAVE":
        raise G.SignalError(f"{path} is not a RIFF/WAVE file")
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(b):
        cid, size = b[pos:pos + 4], struct.unpack_from("<I", b, pos + 4)[0]
        body = b[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, nch, fs, _, _, bits = struct.unpack_from("<HHIIHH", body, 0)
            if tag == 0xFFFE and size >= 26:
                tag = struct.unpack_from("<H", body, 24)[0]        # first two bytes of the sub-format GUID
            fmt = (tag, nch, fs, bits)
        elif cid == b"data":
            data = body
        pos += 8 + size + (size & 1)
    if fmt is None or data is None:
        raise G.SignalError(f"{path}: missing fmt or data chunk")
    tag, nch, fs, bits = fmt
    dt = {(1, 16): np.int16, (3, 32): np.float32, (3, 64): np.float64}.get((tag, bits))
    if dt is None:
        raise G.SignalError(f"{path}: WAV format tag {tag} with {bits} bits is not supported (PCM16, Float32, Float64)")
    raw = np.frombuffer(data, dtype=np.dtype(dt).newbyteorder("<")).astype(dt, copy=False)
    raw = raw[:len(raw) // nch * nch].reshape(-1, nch)
    return WavRaw(raw), float(fs)


def WavFile(path, fs=None):
    """`Signal("file.wav"[, fs])`: errors like src/WAV.jl:10-13 when the expected frame rate differs."""
    wav, file_fs = read_wav(path)
    if fs is not None and float(G.inHz(fs)) != file_fs:
        raise G.SignalError(f"Expected file {path} to have framerate {fs}. If you wish to convert the frame rate, "
                            "you can use `ToFramerate`.")
    return WavSignal(wav, file_fs)
