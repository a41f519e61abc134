"""CPU `sink` — TEST INFRASTRUCTURE: a numpy restatement of the reference's
block-pull materialiser, used as the parity oracle for the CUDA path.

It consumes the same lazy graph the GPU sink lowers (the graph layer stands in
for the reference's unchanged L5/L4 Julia code) and follows the reference's own
protocol: `sink!` asks the root for blocks with `nextblock(x,maxlen,skip[,prev])`
and reads each block's frames exactly once.  `frames(x, block, a, b)` below is
the reference's `frame(x, block, i)` for i = a+1..b, vectorised with numpy.

Restated (file:line in /root/reference):
  sink / sink!            src/sink.jl:87-99,115-121,158-168,225-241
  arrays                  src/arrays.jl:118-132
  numbers                 src/numbers.jl:59-64
  functions               src/functions.jl:44-60,113-114
  Until / After           src/cutting.jl:154-219
  Pad / Extend            src/padding.jl:110-235
  Append                  src/appending.jl:82-110
  MapSignal               src/mapsignal.jl:194-272
  Ramps                   src/ramps.jl:45-119
  Filt (IIR + resample)   src/filters.jl:169-262  (+ DSP.jl kernels in cpu_ref.c)
  Normpower               src/filters.jl:287-314

Deviations from the reference, both towards its intended semantics
(SURVEY.md Appendix C): D-1 Normpower blocks honour their offset; D-2 the
resampler trusts the sample count `filt!` returns instead of `outputlength`'s
estimate.  The GPU path implements the same intended semantics.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np

import signalops.graph as G
from signalops import dspjl as host_dspjl
from signalops.functors import is_sin

from . import dspjl_ref as D


# ---- block records ---------------------------------------------------------------
class Block:
    __slots__ = ("n", "kw")

    def __init__(self, n, **kw):
        self.n = int(n)
        self.kw = kw

    def __getattr__(self, k):
        try:
            return self.kw[k]
        except KeyError:
            raise AttributeError(k) from None


def nf(block):
    return 0 if block is None else block.n


_EMPTY = Block(0, empty=True)


def _inf(n):
    return G.isknowninf(n)


# ---- nextblock / frames per node type -------------------------------------------------

def nextblock(x, maxlen, skip, prev=None):
    return _NEXT[type(x)](x, maxlen, skip, prev)


def frames(x, block, a, b):
    """frame(x, block, i) for i in a+1..b, as an (b-a, nchannels) array."""
    return _FRAMES[type(x)](x, block, a, b)


# arrays (src/arrays.jl:118-132)
def _array_next(x, maxlen, skip, prev):
    offset = 0 if prev is None else prev.state + prev.n
    N = x.data.shape[0]
    if offset < N:
        ln = min(maxlen, N - offset)
        return Block(ln, state=offset)
    return None


def _array_frames(x, block, a, b):
    m = x.matrix()[block.state + a: block.state + b]
    return m.astype(np.int64) if m.dtype.kind in "iub" else m


# numbers (src/numbers.jl:59-64)
def _number_next(x, maxlen, skip, prev):
    return Block(maxlen)


def _number_frames(x, block, a, b):
    return np.full((b - a, 1), x.val, dtype=x.sampletype)


# functions (src/functions.jl:44-60,113-114)
def _function_next(x, maxlen, skip, prev):
    return Block(maxlen, offset=0 if prev is None else prev.offset + prev.n)


def _function_frames(x, block, a, b):
    if isinstance(x.fn, G.RandFn):
        if hasattr(x.fn.rng, "frames"):          # counter-based stream (PhiloxRNG): a function of the frame index
            return x.fn.rng.frames(a + 1 + block.offset, b + 1 + block.offset).reshape(-1, 1)
        return x.fn.rng.standard_normal(b - a).reshape(-1, 1)
    i = np.arange(a + 1, b + 1, dtype=np.float64) + block.offset
    fs = x.framerate
    if is_sin(x.fn):
        u = (i / fs * x.omega + x.phi) if x.omega is not None else (i / fs + x.phi)
        return _sinpi(2 * u).reshape(-1, 1)
    if x.omega is not None:
        arg = 2 * math.pi * np.fmod(i / fs * x.omega + x.phi, 1.0)
    else:
        arg = i / fs + x.phi
    out = np.empty((b - a, x.nchannels), dtype=np.float64)
    for k, v in enumerate(arg):
        r = x.fn(float(v))
        out[k, :] = r if isinstance(r, tuple) else (r,)
    return out


def _sinpi(x):
    """Julia `sinpi`: exact at integers and half-integers, argument reduced exactly."""
    x = np.asarray(x, dtype=np.float64)
    r = np.fmod(x, 2.0)
    r = np.where(r > 1.0, r - 2.0, np.where(r < -1.0, r + 2.0, r))      # [-1,1]
    r = np.where(r > 0.5, 1.0 - r, np.where(r < -0.5, -1.0 - r, r))      # [-.5,.5]
    return np.sin(np.pi * r)


# Until / After (src/cutting.jl:154-219)
def _cut_next(x, maxlen, skip, prev):
    ch = x.signal
    if x.kind == "after":
        if prev is None:
            k = x.resolvelen()
            if k == 0:
                cb = nextblock(ch, maxlen, False)
                return None if cb is None else Block(cb.n, child=cb)
            ln = max(0, k)
            cb = nextblock(ch, ln, True)
            skipped = nf(cb)
            while cb is not None and skipped < ln:
                cb = nextblock(ch, min(maxlen, ln - skipped), True, cb)
                if cb is None:
                    break
                skipped += cb.n
            if skipped < ln:
                raise G.SignalError(f"Signal is too short to skip {x.time}")
            prev = Block(0, child=cb)
        if prev.child is None:
            return None
        cb = nextblock(ch, maxlen, skip, prev.child)
        return None if cb is None else Block(cb.n, child=cb)
    # until
    if prev is None:
        prev = Block(0, child=None, left=x.resolvelen())
    nextlen = prev.left - prev.n
    if nextlen > 0:
        cb = nextblock(ch, min(nextlen, maxlen), skip, prev.child) if prev.child is not None \
            else nextblock(ch, min(nextlen, maxlen), skip)
        if cb is not None:
            return Block(cb.n, child=cb, left=nextlen)
    return None


def _cut_frames(x, block, a, b):
    return frames(x.signal, block.child, a, b)


# Pad / Extend (src/padding.jl:110-235)
def _usepad(x, last_child_block):
    p, T, C = x.pad, x.sampletype, x.signal.nchannels
    if G._isnumber(p):
        return np.full((1, C), np.dtype(T).type(p))
    if isinstance(p, (tuple, list, np.ndarray)):
        return np.asarray(p, dtype=T).reshape(1, -1)
    if p is G.lastframe:
        if last_child_block is None:
            raise G.SignalError("Signal is length zero; there is no last frame to pad with.")
        n = last_child_block.n
        return frames_again_last(x.signal, last_child_block, n)
    if p is G.cycle or p is G.mirror:
        if not isinstance(x.signal, G.ArraySignal):
            raise G.SignalError("Attemped to specify an indexing pad function for a signal which is "
                                "not known to support `getindex`.")
        return ("index", p)
    if callable(p):
        try:
            return np.full((1, C), p(np.dtype(T)))
        except TypeError:
            raise G.SignalError(f"Pad function ({p}) must take 1 or 3 arguments.") from None
    raise G.SignalError(f"unsupported padding {p!r}")


_LASTFRAME_CACHE = {}


def frames_again_last(child, block, n):
    """`usepad(...,lastframe,block) = frame(x,block,nframes(block))` re-reads the last
    frame of the previous block; blocks remember it so stateful leaves are not re-drawn."""
    v = _LASTFRAME_CACHE.get(id(block))
    if v is None:
        v = frames(child, block, n - 1, n)
    return v.reshape(1, -1)


def _pad_next(x, maxlen, skip, prev):
    if prev is None:
        cb = nextblock(x.signal, maxlen, skip)
        if cb is None:
            return Block(maxlen, pad=_usepad(x, None), child=None, offset=0)
        return Block(cb.n, pad=None, child=cb, offset=0)
    if prev.pad is None:
        cb = nextblock(x.signal, maxlen, skip, prev.child)
        if cb is None:
            return Block(maxlen, pad=_usepad(x, prev.child), child=None, offset=prev.n + prev.offset)
        return Block(cb.n, pad=None, child=cb, offset=prev.n + prev.offset)
    return Block(maxlen, pad=prev.pad, child=None, offset=prev.n + prev.offset)


def _pad_frames(x, block, a, b):
    if block.pad is None:
        out = frames(x.signal, block.child, a, b)
        if b == block.n and len(out):
            _LASTFRAME_CACHE.clear()
            _LASTFRAME_CACHE[id(block.child)] = out[-1:].copy()
        return out
    if isinstance(block.pad, tuple):
        fn = block.pad[1]
        m = x.signal.matrix()
        idx = np.arange(a, b) + block.offset            # 0-based global frame index
        N = m.shape[0]
        if fn is G.cycle:
            src = idx % N
        else:
            cnt, rem = np.divmod(idx, N)
            src = np.where(cnt % 2 == 0, rem, N - 1 - rem)
        return m[src]
    return np.repeat(block.pad, b - a, axis=0)


# Append (src/appending.jl:82-110)
def _append_next(x, maxlen, skip, prev):
    if prev is None:
        k = 0
        cb = nextblock(x.signals[0], maxlen, skip)
    else:
        k = prev.k
        cb = nextblock(x.signals[k], maxlen, skip, prev.child)
    K = len(x.signals)
    while k < K - 1 and cb is None:
        k += 1
        cb = nextblock(x.signals[k], maxlen, skip)
    return None if cb is None else Block(cb.n, child=cb, k=k)


def _append_frames(x, block, a, b):
    return frames(x.signals[block.k], block.child, a, b).astype(x.sampletype, copy=False)


# MapSignal (src/mapsignal.jl:194-272)
def _map_next(x, maxlen, skip, prev):
    sigs = x.padded_signals
    if prev is None:
        prev = Block(0, offset=0, blocks=[_EMPTY] * len(sigs), offsets=[0] * len(sigs))
    N = x.nframes
    if not _inf(N):
        maxlen = min(maxlen, N - (prev.offset + prev.n))
    if maxlen == 0:
        return None
    offsets = []
    for o, cb in zip(prev.offsets, prev.blocks):
        o += prev.n
        offsets.append(0 if o == nf(cb) else o)
    blocks = []
    for s, cb, o in zip(sigs, prev.blocks, offsets):
        if o == 0:
            blocks.append(nextblock(s, maxlen, skip) if cb is _EMPTY else nextblock(s, maxlen, skip, cb))
        else:
            blocks.append(cb)
    ln = min(maxlen, min(nf(cb) - o for cb, o in zip(blocks, offsets)))
    return Block(ln, offset=prev.offset + prev.n, blocks=blocks, offsets=offsets)


def _map_frames(x, block, a, b):
    ins = [frames(s, cb, a + o, b + o) for s, cb, o in zip(x.padded_signals, block.blocks, block.offsets)]
    fn = x.fn
    T = x.sampletype
    if x.bychannel:
        if isinstance(fn, G.ToEltypeFn):
            return ins[0].astype(fn.T)
        if x.op == "neg" or (x.op == "-" and len(ins) == 1):
            return (-ins[0]).astype(T, copy=False)
        if x.op in ("+", "-", "*", "/"):
            acc = ins[0].astype(T, copy=False)
            for v in ins[1:]:
                v = v.astype(T, copy=False)
                acc = acc + v if x.op == "+" else acc - v if x.op == "-" else acc * v if x.op == "*" else acc / v
            return acc.astype(T, copy=False)
        out = np.empty((b - a, x.nchannels), dtype=T)
        for i in range(b - a):
            for c in range(x.nchannels):
                out[i, c] = fn(*[v[i, c] for v in ins])
        return out
    if isinstance(fn, G.AsNChannels):
        return np.repeat(ins[0][:, :1], fn.ch, axis=1)
    if isinstance(fn, G.As1Channel):
        acc = ins[0][:, 0].copy()
        for c in range(1, ins[0].shape[1]):
            acc = acc + ins[0][:, c]                     # `sum(tuple)`: left to right
        return acc.reshape(-1, 1)
    if isinstance(fn, G.GetChanFn):
        return ins[0][:, fn.n - 1: fn.n]
    if isinstance(fn, G.TupleCat):
        return np.concatenate([v.astype(T, copy=False) for v in ins], axis=1)
    out = np.empty((b - a, x.nchannels), dtype=T)
    for i in range(b - a):
        out[i, :] = fn(*[tuple(v[i]) for v in ins])
    return out


# Ramps (src/ramps.jl:45-119)
def _ramp_next(x, maxlen, skip, prev):
    N = x.nframes
    L = x.resolvelen()
    big = 1 << 62
    Nn = big if _inf(N) else N
    if x.direction == "on":
        if prev is None:
            return Block(min(L, maxlen), ramp=True, marker=L, stop=Nn, offset=0)
        offset = prev.offset + prev.n
        if prev.ramp:
            ln = min(Nn - offset, maxlen, prev.marker - offset)
            if ln == 0:
                ln = min(Nn - offset, maxlen)
                return Block(ln, ramp=False, marker=prev.marker, stop=prev.stop, offset=offset)
            return Block(ln, ramp=True, marker=prev.marker, stop=prev.stop, offset=offset)
        ln = min(Nn - offset, maxlen, prev.stop - offset)
        return Block(ln, ramp=False, marker=prev.marker, stop=prev.stop, offset=offset) if ln > 0 else None
    if prev is None:
        start = Nn - L
        return Block(min(start, maxlen), ramp=False, marker=start, stop=Nn, offset=0)
    offset = prev.offset + prev.n
    if not prev.ramp:
        ln = min(Nn - offset, maxlen, prev.marker - offset)
        if ln == 0:
            ln = min(Nn - offset, maxlen)
            return Block(ln, ramp=True, marker=prev.marker, stop=prev.stop, offset=offset)
        return Block(ln, ramp=False, marker=prev.marker, stop=prev.stop, offset=offset)
    ln = min(Nn - offset, maxlen, prev.stop - offset)
    return Block(ln, ramp=True, marker=prev.marker, stop=prev.stop, offset=offset) if ln > 0 else None


def _ramp_frames(x, block, a, b):
    T = x.sampletype
    C = x.nchannels
    if not block.ramp:
        return np.ones((b - a, C), dtype=T)
    i = np.arange(a + 1, b + 1, dtype=np.float64)
    if x.direction == "on":
        arg = (i + block.offset - 1) / block.marker
    else:
        start = block.marker - block.offset
        stop = block.stop - block.offset
        arg = 1 - (i - start) / (stop - start)
    if x.fn is G.sinramp:
        v = _sinpi(0.5 * arg)
    else:
        v = np.array([x.fn(float(t)) for t in arg], dtype=np.float64)
    return np.repeat(v.astype(T).reshape(-1, 1), C, axis=1)


# Filt (src/filters.jl:169-262)
def _resolve_filter(x):
    """`resolve_filter(x.fn(framerate(x)))` with the oracle's own design code."""
    fn, fs = x.fn, x.framerate
    if isinstance(fn, G.ResamplerFn):
        return D.Resampler(fn.ratio if isinstance(fn.ratio, Fraction) else float(fn.ratio))
    if isinstance(fn, G.FilterFn):
        z, p, k = D.design_zpk(fn.design.__name__, [G.inHz(a) for a in fn.args], G.inHz(fs), fn.method.spec)
        return D.zpk2sos_dspjl(z, p, k)
    h = fn(fs)
    if isinstance(h, host_dspjl.SecondOrderSections):
        return h.coef_table(), h.g
    if isinstance(h, host_dspjl.ZeroPoleGain):
        return D.zpk2sos_dspjl(h.z, h.p, h.k)
    if isinstance(h, host_dspjl.Biquad):
        return np.array([h.astuple()]), 1.0
    if isinstance(h, host_dspjl.PolynomialRatio):
        if len(h.b) > 3 or len(h.a) > 3:
            # DSP.jl runs the order-n DF2T recurrence with an (n-1)... state vector: that is
            # scipy.signal.lfilter's structure; the state is carried between blocks
            return DirectForm(h.b, h.a)
        b = list(h.b) + [0.0] * (3 - len(h.b))
        a = list(h.a) + [0.0] * (3 - len(h.a))
        return np.array([[b[0], b[1], b[2], a[1], a[2]]]), 1.0
    if isinstance(h, host_dspjl.FIRFilter) and h.kind == "standard":
        return D.StandardFIR(h.h)
    raise TypeError(f"oracle cannot resolve filter {h!r}")


class DirectForm:
    """Order-n transposed direct form II with streaming state (DSP.jl `DF2TFilter(::PolynomialRatio)`)."""

    def __init__(self, b, a):
        from scipy import signal as sps
        n = max(len(b), len(a))
        self.b = np.concatenate([np.asarray(b, dtype=np.float64), np.zeros(n - len(b))])
        self.a = np.concatenate([np.asarray(a, dtype=np.float64), np.zeros(n - len(a))])
        self.zi = np.zeros(n - 1)
        self._lfilter = sps.lfilter

    rate = 1.0

    def outputlength(self, n):
        return n

    def filt(self, x):
        y, self.zi = self._lfilter(self.b, self.a, x, zi=self.zi)
        return y


def _filter_init(x):
    C = x.signal.nchannels
    hs = [_resolve_filter(x) for _ in range(C)]          # src/filters.jl:205 (one per channel)
    resamp = isinstance(hs[0], (D.Resampler, D.StandardFIR, DirectForm))
    N = x.nframes
    bs = x.blocksize
    if resamp and not isinstance(x.fn, G.ResamplerFn):   # single-rate streaming objects: plain blocks
        inlen = bs if _inf(N) else min(N, bs)
    elif resamp:                                         # init_length, src/filters.jl:185-199
        ratio = float(x.fn.ratio)
        n = int(max(1, (bs if _inf(N) else min(N, bs)) / ratio))
        if hs[0].outputlength(n) <= 0:
            n = int(max(1, bs / ratio))
            if hs[0].outputlength(n) <= 0:
                raise G.SignalError("Blocksize is too small for this resampling filter.")
        inlen = n
    else:
        inlen = bs if _inf(N) else min(N, bs)
        hs = [(c, g, np.zeros((c.shape[0], 2))) for (c, g) in hs]
    return Block(0, last_out=0, avail=0, hs=hs, resamp=resamp, inlen=inlen, output=None,
                 child=None, started=False, produced=0)


def _filter_next(x, maxlen, skip, prev):
    if prev is None:
        prev = _filter_init(x)
    last_out = prev.last_out + prev.n
    N = x.nframes
    if last_out < prev.avail:
        ln = min(maxlen, prev.avail - last_out)
        return Block(ln, last_out=last_out, avail=prev.avail, hs=prev.hs, resamp=prev.resamp,
                     inlen=prev.inlen, output=prev.output, child=prev.child, started=True,
                     produced=prev.produced)
    if not _inf(N) and prev.produced >= N:
        return None
    # pull the next input block from Pad(x.signal, zero)   (src/filters.jl:240-244)
    psig = G.Pad(x.signal, G.zero)
    inbuf = np.zeros((prev.inlen, x.signal.nchannels), dtype=x.signal.sampletype)
    cb = nextblock(psig, prev.inlen, False, prev.child) if prev.started else nextblock(psig, prev.inlen, False)
    cb = _sink_blocks(inbuf, psig, cb)
    outs = []
    for ch in range(inbuf.shape[1]):
        h = prev.hs[ch]
        if prev.resamp:
            outs.append(h.filt(inbuf[:, ch].astype(np.float64)))      # D-2: actual count
        else:
            coef, g, st = h
            outs.append(D.sos_filt(inbuf[:, ch].astype(np.float64), coef, g, st))
    out_len = len(outs[0])
    if out_len <= 0 and not prev.resamp:
        raise G.SignalError("Unexpected non-positive output length!")
    output = np.stack(outs, axis=1).astype(x.sampletype) if out_len else np.zeros((0, inbuf.shape[1]), x.sampletype)
    if out_len == 0:
        # a resampler block that produced nothing yet: keep pulling
        nxt = Block(0, last_out=0, avail=0, hs=prev.hs, resamp=prev.resamp, inlen=prev.inlen,
                    output=output, child=cb, started=True, produced=prev.produced)
        return _filter_next(x, maxlen, skip, nxt)
    return Block(min(maxlen, out_len), last_out=0, avail=out_len, hs=prev.hs, resamp=prev.resamp,
                 inlen=prev.inlen, output=output, child=cb, started=True,
                 produced=prev.produced + out_len)


def _filter_frames(x, block, a, b):
    return block.output[block.last_out + a: block.last_out + b]


# Normpower (src/filters.jl:287-314)
def _normed_next(x, maxlen, skip, prev):
    if prev is None:
        N = x.nframes
        if _inf(N):
            raise G.SignalError("Cannot normalize an infinite-length signal. Please "
                                "use `Until` to take a prefix of the signal")
        vals = np.zeros((N, x.nchannels), dtype=x.sampletype)
        sink_into(vals, x.signal, _force_channels=False)
        rms = np.sqrt(np.mean(vals.astype(x.sampletype) ** 2, dtype=x.sampletype))
        vals = (vals / rms).astype(x.sampletype)
        prev = Block(0, offset=0, vals=vals)
    offset = prev.offset + prev.n
    ln = min(maxlen, x.nframes - offset)
    return Block(ln, offset=offset, vals=prev.vals) if ln > 0 else None


def _normed_frames(x, block, a, b):
    return block.vals[block.offset + a: block.offset + b]       # D-1: honours the offset


_NEXT = {G.ArraySignal: _array_next, G.NumberSignal: _number_next, G.SignalFunction: _function_next,
         G.CutApply: _cut_next, G.PaddedSignal: _pad_next, G.AppendSignals: _append_next,
         G.MapSignal: _map_next, G.RampSignal: _ramp_next, G.FilteredSignal: _filter_next,
         G.NormedSignal: _normed_next}
_FRAMES = {G.ArraySignal: _array_frames, G.NumberSignal: _number_frames,
           G.SignalFunction: _function_frames, G.CutApply: _cut_frames, G.PaddedSignal: _pad_frames,
           G.AppendSignals: _append_frames, G.MapSignal: _map_frames, G.RampSignal: _ramp_frames,
           G.FilteredSignal: _filter_frames, G.NormedSignal: _normed_frames}


# ---- sink / sink! (src/sink.jl) -----------------------------------------------------------

def _sink_blocks(result, x, block):
    """sink!(result,x,::IsSignal,block), src/sink.jl:227-241"""
    written = 0
    N = result.shape[0]
    while block is not None and written < N:
        assert block.n > 0
        result[written: written + block.n] = frames(x, block, 0, block.n)     # sink_helper!
        written += block.n
        maxlen = N - written
        if maxlen > 0:
            block = nextblock(x, maxlen, False, block)
    assert written == N, f"wrote {written} of {N} frames"
    return block


def sink_into(result, x, _force_channels=True):
    """sink!(result, x), src/sink.jl:158-168"""
    x = G.Signal(x)
    n = result.shape[0]
    xn = x.nframes
    if xn is not None and not _inf(xn) and xn < n:
        raise G.SignalError(f"Signal is too short to fill buffer of length {n}.")
    view = result.reshape(-1, 1) if result.ndim == 1 else result
    if _force_channels:
        x = G.ToChannels(x, view.shape[1])
    _sink_blocks(view, x, nextblock(x, n, False))     # called even when n == 0 (errors still fire)
    return result


def sink(x, to=None):
    """sink(x[,to]), src/sink.jl:28-37,87-99,115-121. Always copies (the reference's
    zero-copy view path for raw arrays, src/sink.jl:65-76, is a CPU-only nicety)."""
    x = G.process_sink_params(x)
    result = np.zeros((x.nframes, x.nchannels), dtype=x.sampletype)
    sink_into(result, x)
    if to == "array":
        return result
    if to == "tuple" or G.result_wants_tuple(x):
        return result, x.framerate
    return result
