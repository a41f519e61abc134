"""Lazy signal graph: the host-side mirror of the reference's L5/L4 layers.

In the Julia integration this layer is the reference's own unchanged code
("L5/L4 … stay as the reference's Julia code"); `sink(x, GPUSink)`
only reads the finished graph.  Julia is not available in this image, so the
operator API is mirrored here in Python with the same names, argument meaning
and error behaviour, so that the parity tests read like test/runtests.jl.

Every class/func cites the reference definition it mirrors.  Nothing in this
file computes samples: it only builds immutable nodes and answers the trait
queries (nframes / nchannels / framerate / sampletype) plus the ToFramerate
push-down rewrites.
"""
from __future__ import annotations

import math
import operator

import numpy as np

from . import dspjl
from .units import (Quantity, gain_to_amplitude, inframes, inHz, inradians,
                    inseconds, maybeseconds, ms)

default_blocksize = 2 ** 12          # src/filters.jl:3
MAX_CHANNEL_STACK = 64               # src/mapsignal.jl:192


class SignalError(Exception):
    """Julia `error(...)` -> ErrorException."""


# ----------------------------------------------------------------------------
# length algebra (src/inflen.jl, src/signal.jl:28-37, src/numbers.jl:5-9)

class Infinite:
    def __repr__(self):
        return type(self).__name__


class InfiniteLength(Infinite):
    def __repr__(self):
        return "inflen"


class Extended(Infinite):
    def __init__(self, len_):
        self.len = len_


class NumberExtended(Infinite):
    pass


inflen = InfiniteLength()
numextend = NumberExtended()


def isknowninf(x):
    return isinstance(x, Infinite)


def cleanextend(x):
    return inflen if isinstance(x, (Extended, NumberExtended)) else x


def _tolen(x):
    if isinstance(x, Extended):
        return x.len
    if isinstance(x, NumberExtended):
        return 0
    return x


def _len_max(a, b):
    if a is None or b is None:
        return None
    if isinstance(a, Infinite) or isinstance(b, Infinite):
        return inflen
    return max(a, b)


def _maxlen(x, y):
    """src/mapsignal.jl:147-153"""
    if isinstance(x, NumberExtended) and isinstance(y, NumberExtended):
        return x
    return _len_max(_tolen(x), _tolen(y))


def _len_min(a, b):
    if a is None or b is None:
        return None
    if isinstance(a, Infinite):
        return b
    if isinstance(b, Infinite):
        return a
    return min(a, b)


# ----------------------------------------------------------------------------
# element types

def _float_of(T):
    """Julia `float(T)`"""
    T = np.dtype(T)
    return T if T.kind == "f" else np.dtype(np.float64)


def promote_type(*types):
    """Julia promote_type for the sample types that occur here (floats beat
    ints; the widest float wins; Float32 with Int64 stays Float32)."""
    types = [np.dtype(t) for t in types]
    floats = [t for t in types if t.kind == "f"]
    if floats:
        return max(floats, key=lambda t: t.itemsize)
    if all(t.kind == "b" for t in types):
        return np.dtype(np.bool_)
    return np.dtype(np.int64)


# ----------------------------------------------------------------------------
# base classes (src/signal.jl:18-48, src/wrapping.jl)

class AbstractSignal:
    evaltrait = "computed"        # src/signal.jl:174-175

    def nframes_helper(self):
        raise SignalError(f"Undefined `nframes_helper` for {type(self).__name__}")

    @property
    def nframes(self):
        return cleanextend(self.nframes_helper())

    def root(self):
        return self

    def __rshift__(self, f):
        """`x >> Until(5*s) >> Ramp()` stands in for Julia's `x |> Until(5s) |> Ramp`."""
        return f(self)


class WrappedSignal(AbstractSignal):
    """src/wrapping.jl:3-19 — single-child nodes forward traits to child()."""

    def child(self):
        raise NotImplementedError

    @property
    def evaltrait(self):
        return self.child().evaltrait

    @property
    def nchannels(self):
        return self.child().nchannels

    @property
    def framerate(self):
        return self.child().framerate

    @property
    def sampletype(self):
        return self.child().sampletype

    def nframes_helper(self):
        return self.child().nframes_helper()

    def root(self):
        return self.child().root()


def nframes(x):
    return Signal(x).nframes


def nchannels(x):
    return Signal(x).nchannels


def framerate(x):
    return Signal(x).framerate


def sampletype(x):
    return Signal(x).sampletype


def duration(x):
    x = Signal(x)
    n, fs = x.nframes, x.framerate
    if isinstance(n, Infinite):
        return inflen
    if n is None or fs is None:
        return None
    return n / fs


# ----------------------------------------------------------------------------
# leaves

class ArraySignal(AbstractSignal):
    """Arrays and (array, fs) tuples as data signals (src/arrays.jl:35-98)."""
    evaltrait = "data"

    def __init__(self, data, fs=None):
        data = np.asarray(data)
        if data.ndim not in (1, 2):
            raise SignalError("To treat an array as a signal it must have 1 or 2 dimensions")
        if data.dtype.kind not in "fiub":
            raise SignalError(f"Don't know how create a signal from array of {data.dtype}.")
        self.data = data
        self._fs = None if fs is None else float(inHz(fs))

    framerate = property(lambda self: self._fs)
    nchannels = property(lambda self: 1 if self.data.ndim == 1 else self.data.shape[1])
    sampletype = property(lambda self: self.data.dtype if self.data.dtype.kind == "f"
                          else np.dtype(np.int64))

    def nframes_helper(self):
        return self.data.shape[0]

    def matrix(self):
        d = self.data
        return d.reshape(-1, 1) if d.ndim == 1 else d


class NumberSignal(AbstractSignal):
    """src/numbers.jl:1-11,47-49 — constant, infinite, mono."""

    def __init__(self, val, fs=None, dB=False):
        self.val = val
        self._fs = None if fs is None else float(inHz(fs))
        self.dB = dB

    framerate = property(lambda self: self._fs)
    nchannels = 1

    @property
    def sampletype(self):
        if isinstance(self.val, (np.floating, np.integer)):
            return np.dtype(type(self.val))
        return np.dtype(np.float64) if isinstance(self.val, float) else np.dtype(np.int64)

    def nframes_helper(self):
        return numextend


class Functor:
    """src/functions.jl:65-66 — hook for callable objects usable as generators."""

    def __call__(self, x):
        raise NotImplementedError


class RandFn:
    """src/functions.jl:98-114 — `Signal(randn; rng)`; one draw per evaluated frame."""

    def __init__(self, rng):
        self.rng = rng


def randn(*a, **k):
    raise SignalError("`randn` is only a marker for Signal(randn, ...)")


def sin(x):
    """Marker for Julia's `sin` (gets the `sinpi` special case, src/functions.jl:57-60)."""
    return np.sin(x)


def cos(x):
    return np.cos(x)


class SignalFunction(AbstractSignal):
    """src/functions.jl:11-28"""

    def __init__(self, fn, first, omega, phi, fs=None):
        self.fn = fn
        self.first = first
        self.omega = omega
        self.phi = float(phi)
        self._fs = None if fs is None else float(fs)

    framerate = property(lambda self: self._fs)
    nchannels = property(lambda self: len(self.first))

    @property
    def sampletype(self):
        v = self.first[0]
        if isinstance(v, (np.floating, np.integer)):
            return np.dtype(type(v))
        return np.dtype(np.float64) if isinstance(v, float) else np.dtype(np.int64)

    def nframes_helper(self):
        return inflen


def _astuple(v):
    if isinstance(v, tuple):
        return v
    if isinstance(v, (int, float, np.number)) and not isinstance(v, bool):
        return (v,)
    if isinstance(v, np.ndarray) and v.ndim == 0:
        return (v[()],)
    raise SignalError("Function must return number or tuple of numbers.")


# ----------------------------------------------------------------------------
# Signal() coercion (src/signal.jl:96-149, arrays.jl:35-45,76-84, numbers.jl:47-49,
# functions.jl:88-96,110-111)

def _isnumber(x):
    return isinstance(x, (int, float, np.number)) and not isinstance(x, (bool, np.bool_))


def Signal(x=None, fs=None, **kwds):
    if x is None:
        return lambda y: Signal(y, fs, **kwds)
    if isinstance(x, Quantity) and x.kind == "freq" and fs is None:
        return lambda y: Signal(y, x, **kwds)
    fs_hz = None if fs is None else float(inHz(fs))
    if isinstance(x, AbstractSignal):
        if kwds:
            raise TypeError("keyword arguments only apply to function signals")
        if x.framerate is None:
            return ToFramerate(x, fs_hz)
        if fs_hz is not None and x.framerate != fs_hz:
            raise SignalError(f"Signal expected to have frame rate of {fs_hz} Hz.")
        return x
    if isinstance(x, tuple) and len(x) == 2 and isinstance(x[0], (np.ndarray, list, range)):
        afs = float(inHz(x[1]))
        if fs_hz is not None and afs != fs_hz:
            raise SignalError(f"Signal expected to have frame rate of {fs_hz} Hz.")
        return ArraySignal(x[0], afs)
    if isinstance(x, (np.ndarray, list, range)):
        return ArraySignal(x, fs_hz)
    if isinstance(x, Quantity):
        if x.kind == "gain":
            return NumberSignal(gain_to_amplitude(x), fs_hz, dB=True)
        raise SignalError(f"Don't know how create a signal from {x!r}.")
    if _isnumber(x):
        return NumberSignal(x, fs_hz)
    if x is randn:
        rng = kwds.pop("rng", None)
        if kwds:
            raise TypeError(f"unexpected keyword arguments {list(kwds)}")
        if rng is None:
            rng = np.random.default_rng()
        return SignalFunction(RandFn(rng), (0.0,), None, 0.0, fs_hz)
    if callable(x):
        omega = kwds.pop("ω", kwds.pop("omega", None))
        omega = kwds.pop("frequency", omega)
        phi = 0
        for key in ("\u03d5", "\u03c6", "phi", "phase"):   # Python NFKC-normalises ϕ to φ in keywords
            if key in kwds:
                phi = kwds.pop(key)
        if kwds:
            raise TypeError(f"unexpected keyword arguments {list(kwds)}")
        try:
            first = _astuple(x(0.0))
            if omega is None:
                p = inseconds(phi)
            else:
                p = inradians(phi, omega) / (2 * math.pi)
        except ValueError as e:
            raise SignalError(str(e)) from None
        return SignalFunction(x, first, inHz(omega), p, fs_hz)
    if isinstance(x, str):
        raise SignalError("No backend loaded for files: file IO is outside the GPU sink path.")
    raise SignalError(f"Don't know how create a signal from {x!r}.")


def _nosignal(x):
    raise SignalError(f"Value is not a signal: {x!r}")


# ----------------------------------------------------------------------------
# cutting (src/cutting.jl)

class CutApply(WrappedSignal):
    def __init__(self, signal, time, kind):
        self.signal = signal
        self.time = time
        self.kind = kind          # "until" | "after"

    def child(self):
        return self.signal

    @property
    def evaltrait(self):
        # src/cutting.jl:138 — After is always a DataSignal
        return "data" if self.kind == "after" else self.signal.evaltrait

    def resolvelen(self):
        """src/cutting.jl:32"""
        try:
            return inframes(maybeseconds(self.time), self.framerate)
        except ValueError as e:
            raise SignalError(str(e)) from None

    def nframes_helper(self):
        n = self.signal.nframes_helper()
        k = self.resolvelen()
        if k is None:
            return None
        if self.kind == "until":
            return _len_min(n, max(0, k))            # :130
        if n is None:
            return None
        if isinstance(n, Infinite):
            return inflen if not isinstance(n, Extended) else n
        return min(max(n - k, 0), n)                 # :134 clamp(n-k,0,n)


def _curry_time(args, make):
    if len(args) == 1:
        return lambda x: make(x, args[0])
    return make(*args)


def Until(*args):
    return _curry_time(args, lambda x, t: CutApply(Signal(x), t, "until"))


def After(*args):
    return _curry_time(args, lambda x, t: CutApply(Signal(x), t, "after"))


def Window(x=None, *, from_=None, to=None, at=None, width=None, **kw):
    """src/cutting.jl:45-57 (`from` is a Python keyword -> `from_`; also accepted
    through **{'from': ...})."""
    if "from" in kw:
        from_ = kw.pop("from")
    if x is None:
        return lambda y: Window(y, from_=from_, to=to, at=at, width=width)
    if (at is None) != (width is None) or (from_ is None) != (to is None) or \
            (at is None) == (from_ is None):
        raise SignalError("`Window` must either use the two keywords `at` and `width` OR"
                          "the two keywords `from` and `to`.")
    if from_ is None:
        after, until = at - width / 2, width
    else:
        after, until = from_, to - from_
    return Until(After(x, after), until)


def _stretchtime(t, scale):
    """src/cutting.jl:140-141"""
    if isinstance(t, Quantity) and t.kind == "frames":
        return Quantity(int(math.floor(t.value * scale)), "frames", "frames")
    return t


# ----------------------------------------------------------------------------
# padding (src/padding.jl)

def zero(T):
    return np.dtype(T).type(0)


def one(T):
    return np.dtype(T).type(1)


def lastframe(x):
    raise SignalError("Must be passed as argument to `Pad`.")


def cycle(x, i, j):
    """0-based restatement of src/padding.jl:132"""
    return x[i % x.shape[0], j]


def mirror(x, i, j):
    """0-based restatement of src/padding.jl:142-148"""
    n = x.shape[0]
    count, rem = divmod(i, n)
    return x[rem if count % 2 == 0 else n - 1 - rem, j]


class PaddedSignal(WrappedSignal):
    def __init__(self, signal, pad, extending=False):
        self.signal = signal
        self.pad = pad
        self.extending = extending

    def child(self):
        return self.signal

    def nframes_helper(self):
        return Extended(self.signal.nframes) if self.extending else inflen


def Pad(*args):
    if len(args) == 1:
        return lambda x: Pad(x, args[0])
    x, p = args
    x = Signal(x)
    return x if isknowninf(x.nframes) else PaddedSignal(x, p)


def Extend(*args):
    if len(args) == 1:
        return lambda x: Extend(x, args[0])
    x, p = args
    x = Signal(x)
    return x if isknowninf(x.nframes) else PaddedSignal(x, p, True)


# ----------------------------------------------------------------------------
# appending (src/appending.jl)

class AppendSignals(WrappedSignal):
    def __init__(self, signals, len_, T):
        self.signals = tuple(signals)
        self.len = len_
        self._T = T

    def child(self):
        return self.signals[0]

    sampletype = property(lambda self: self._T)

    def nframes_helper(self):
        return self.len

    def root(self):
        return _mergeroots([s.root() for s in self.signals])


def Append(*xs):
    if len(xs) == 1:
        y = xs[0]
        return lambda x: Append(x, y)
    xs = Uniform(xs, channels=True)
    if any(isknowninf(x.nframes) for x in xs[:-1]):
        raise SignalError("Cannot Append to the end of an infinite signal")
    El = promote_type(*[x.sampletype for x in xs])
    xs = [x if x.sampletype == El else ToEltype(x, El) for x in xs]
    if any(isknowninf(x.nframes) for x in xs):
        n = inflen
    elif any(x.nframes is None for x in xs):
        n = None
    else:
        n = sum(x.nframes for x in xs)
    return AppendSignals(xs, n, El)


def Prepend(*xs):
    if len(xs) == 1:
        x = xs[0]
        return lambda y: Append(x, y)
    return Append(*reversed(xs))


# ----------------------------------------------------------------------------
# filters (src/filters.jl)

class FilterFn:
    """src/filters.jl:5-12"""

    def __init__(self, design, method, args):
        self.design, self.method, self.args = design, method, args

    def __call__(self, fs):
        return dspjl.digitalfilter(self.design(*[inHz(a) for a in self.args], fs=inHz(fs)),
                                   self.method)


class RawFilterFn:
    """src/filters.jl:89-92"""

    def __init__(self, h):
        self.h = h

    def __call__(self, fs):
        return self.h


class ResamplerFn:
    """src/util.jl:12-15 + src/reformatting.jl:92-99"""

    def __init__(self, ratio, fs):
        self.ratio, self.fs = ratio, fs

    def __call__(self, fs):
        h = dspjl.resample_filter(self.ratio)
        f = dspjl.FIRFilter(h, self.ratio)
        f.setphase(f.timedelay())
        return f


class FilteredSignal(WrappedSignal):
    """src/filters.jl:98-114"""
    evaltrait = "computed"

    def __init__(self, signal, fn, blocksize, newfs):
        self.signal = signal
        self.fn = fn
        self.blocksize = int(blocksize)
        self._fs = newfs

    def child(self):
        return self.signal

    framerate = property(lambda self: self._fs)
    sampletype = property(lambda self: _float_of(self.signal.sampletype))

    def nframes_helper(self):
        """src/filters.jl:159-167"""
        cfs = self.signal.framerate
        if cfs is None:
            return None
        n = self.signal.nframes_helper()
        if self._fs == cfs:
            return n
        if n is None or isinstance(n, Infinite):
            return n
        return int(math.ceil(n * self._fs / cfs))


def _nyquist_check(x, hz):
    if x.framerate is not None and inHz(hz) >= 0.5 * x.framerate:
        raise SignalError(f"The frequency {hz} cannot be represented at a sampling rate "
                          f"of {x.framerate} Hz. Increase the frame rate or lower the frequency.")


def _is_filtertype(a):
    return isinstance(a, type) and issubclass(a, dspjl.FilterType)


def Filt(*args, blocksize=default_blocksize, order=5, method=None, newfs=None):
    """src/filters.jl:54-66,96-97"""
    if args and _is_filtertype(args[0]):
        return lambda x: Filt(x, *args, blocksize=blocksize, order=order, method=method)
    if len(args) == 1:
        h = args[0]
        return lambda x: Filt(x, h, blocksize=blocksize)
    x = Signal(args[0])
    if _is_filtertype(args[1]):
        bounds = args[2:]
        for b in bounds:
            _nyquist_check(x, b)
        m = method if method is not None else dspjl.Butterworth(order)
        fn = FilterFn(args[1], m, bounds)
    elif isinstance(args[1], (FilterFn, ResamplerFn, RawFilterFn)) or callable(args[1]):
        fn = args[1]
    else:
        fn = RawFilterFn(args[1])
    return FilteredSignal(x, fn, blocksize, x.framerate if newfs is None else newfs)


class NormedSignal(WrappedSignal):
    """src/filters.jl:266-285"""

    def __init__(self, signal):
        self.signal = signal

    def child(self):
        return self.signal

    sampletype = property(lambda self: _float_of(self.signal.sampletype))


def Normpower(x):
    """src/filters.jl:323-326"""
    return NormedSignal(Signal(x))


# ----------------------------------------------------------------------------
# maps (src/mapsignal.jl, src/reformatting.jl:132-184)

class ToEltypeFn:
    def __init__(self, T):
        self.T = np.dtype(T)


class AsNChannels:
    def __init__(self, ch):
        self.ch = int(ch)


class As1Channel:
    pass


class GetChanFn:
    def __init__(self, n):
        self.n = int(n)      # 1-based, as in the reference


class TupleCat:
    pass


tuplecat = TupleCat()


def reverse(frame):
    """Julia's `reverse` on a frame tuple: `OperateOn(reverse, x, bychannel=false)` swaps the channel order
    (test/runtests.jl:273-276).  One of the enumerated whole-frame functions the GPU sink lowers."""
    return tuple(reversed(frame))

_ARITH = {operator.add: "+", operator.mul: "*", operator.sub: "-", operator.truediv: "/",
          operator.neg: "neg", "+": "+", "*": "*", "-": "-", "/": "/"}


def default_pad(fn):
    """src/mapsignal.jl:274-276"""
    return one if _ARITH.get(fn) in ("*", "/") else zero


class MapSignal(AbstractSignal):
    """src/mapsignal.jl:8-38"""

    def __init__(self, fn, signals, fs, padding, blocksize, bychannel):
        self.fn = fn
        self.op = _ARITH.get(fn) if not isinstance(fn, (ToEltypeFn, AsNChannels, As1Channel,
                                                          GetChanFn, TupleCat)) else None
        self.signals = tuple(signals)
        self._fs = fs
        self.padding = padding
        self.padded_signals = tuple(Extend(s, padding) for s in signals)
        self.blocksize = blocksize
        self.bychannel = bychannel
        self._nch, self._T = self._infer()

    def _infer(self):
        """type/arity of `fn` applied to zero samples (src/mapsignal.jl:139-144)"""
        sigs, fn = self.signals, self.fn
        types = [s.sampletype for s in sigs]
        if self.bychannel:
            nch = sigs[0].nchannels
            if isinstance(fn, ToEltypeFn):
                return nch, fn.T
            if self.op in ("+", "*", "-", "neg"):
                return nch, promote_type(*types)
            if self.op == "/":
                return nch, _float_of(promote_type(*types))
            v = fn(*[np.dtype(t).type(0) for t in types])
            return nch, np.asarray(v).dtype
        if isinstance(fn, AsNChannels):
            return fn.ch, types[0]
        if isinstance(fn, As1Channel):
            return 1, types[0]
        if isinstance(fn, GetChanFn):
            return 1, types[0]
        if isinstance(fn, TupleCat):
            return sum(s.nchannels for s in sigs), promote_type(*types)
        v = _astuple_seq(fn(*[tuple(np.dtype(t).type(0) for _ in range(s.nchannels))
                               for s, t in zip(sigs, types)]))
        return len(v), promote_type(*[np.asarray(e).dtype for e in v])

    framerate = property(lambda self: self._fs)
    nchannels = property(lambda self: self._nch)
    sampletype = property(lambda self: self._T)

    def nframes_helper(self):
        out = None
        for i, s in enumerate(self.signals):
            n = s.nframes_helper()
            out = n if i == 0 else _maxlen(out, n)
        return out

    def root(self):
        return _mergeroots([s.root() for s in self.signals])


def _astuple_seq(v):
    if isinstance(v, (tuple, list)):
        return tuple(v)
    return _astuple(v)


def OperateOn(fn, *xs, padding=None, bychannel=True, blocksize=default_blocksize):
    """src/mapsignal.jl:131-145"""
    if padding is None:
        padding = default_pad(fn)
    xs = Uniform(xs, channels=bychannel)
    return MapSignal(fn, xs, xs[0].framerate, padding, blocksize, bychannel)


def Operate(fn, *xs, **kw):
    return lambda x: OperateOn(fn, x, *xs, **kw)


def _curried_nary(fn, xs, **kw):
    if len(xs) == 1:
        y = xs[0]
        return lambda x: OperateOn(fn, x, y, **kw)
    return OperateOn(fn, *xs, **kw)


def Mix(*xs):
    return _curried_nary(operator.add, xs)          # src/mapsignal.jl:307-308


def Amplify(*xs):
    return _curried_nary(operator.mul, xs)          # src/mapsignal.jl:332-333


def AddChannel(*xs):
    return _curried_nary(tuplecat, xs, bychannel=False)   # :359-362


def SelectChannel(*args):
    if len(args) == 1:
        return lambda x: SelectChannel(x, args[0])
    return OperateOn(GetChanFn(args[1]), args[0], bychannel=False)    # :388-391


def ToChannels(*args):
    """src/reformatting.jl:139-170"""
    if len(args) == 1:
        return lambda x: ToChannels(x, args[0])
    x, ch = Signal(args[0]), args[1]
    if ch == x.nchannels:
        return x
    if ch == 1:
        return OperateOn(As1Channel(), x, bychannel=False)
    if x.nchannels == 1:
        return OperateOn(AsNChannels(ch), x, bychannel=False)
    raise SignalError(f"No rule to convert signal with {x.nchannels} channels to"
                      f" a signal with {ch} channels.")


def ToEltype(*args):
    if len(args) == 1:
        return lambda x: ToEltype(x, args[0])
    return OperateOn(ToEltypeFn(args[1]), args[0])      # src/reformatting.jl:184


def Format(x, fs, ch=None):
    """src/reformatting.jl:207-213"""
    x = Signal(x)
    if ch is None:
        ch = x.nchannels
    if ch > 1 and x.nchannels == 1:
        return ToChannels(ToFramerate(x, fs), ch)
    return ToFramerate(ToChannels(x, ch), fs)


def Uniform(xs, channels=False):
    """src/reformatting.jl:241-254"""
    xs = [Signal(x) for x in xs]
    rates = [x.framerate for x in xs if x.framerate is not None]
    fs = max(rates) if rates else None
    if not channels:
        return [Format(x, fs) for x in xs]
    ch = max(x.nchannels for x in xs)
    return [Format(x, fs, ch) for x in xs]


# ----------------------------------------------------------------------------
# ramps (src/ramps.jl)

def sinramp(x):
    """src/ramps.jl:4 — sinpi(0.5x)"""
    return math.sin(math.pi * 0.5 * x) if x != 1 else 1.0


def identity(x):
    return x


class RampSignal(WrappedSignal):
    """src/ramps.jl:6-26"""

    def __init__(self, direction, signal, time, fn):
        self.direction = direction     # "on" | "off"
        self.signal = signal
        self.time = time
        self.fn = fn

    def child(self):
        return self.signal

    sampletype = property(lambda self: _float_of(self.signal.sampletype))

    def resolvelen(self):
        k = inframes(maybeseconds(self.time), self.framerate)
        return None if k is None else max(1, k)


def _ramp_args(args, default_len=None):
    """Split (x?, len?, fn?) the way the reference's method table does
    (src/ramps.jl:156-161): a leading Number/Function is never the signal."""
    default_len = 10 * ms if default_len is None else default_len
    args = list(args)
    x = None
    if args and not (_isnumber(args[0]) or isinstance(args[0], Quantity) or
                     (callable(args[0]) and not isinstance(args[0], AbstractSignal)
                      and args[0] not in (sin, cos, randn))):
        x = args.pop(0)
    elif args and args[0] in (sin, cos, randn):
        x = args.pop(0)
    length, fn = default_len, sinramp
    if args and (_isnumber(args[0]) or isinstance(args[0], Quantity)):
        length = args.pop(0)
    if args and callable(args[0]):
        fn = args.pop(0)
    if args:
        raise TypeError("unexpected ramp arguments")
    return x, length, fn


def RampOn(*args):
    x, length, fn = _ramp_args(args)
    if x is None:
        return lambda y: RampOn(y, length, fn)
    x = Signal(x)
    return Amplify(x, RampSignal("on", x, length, fn))


def RampOff(*args):
    x, length, fn = _ramp_args(args)
    if x is None:
        return lambda y: RampOff(y, length, fn)
    x = Signal(x)
    return Amplify(x, RampSignal("off", x, length, fn))


def Ramp(*args):
    x, length, fn = _ramp_args(args)
    if x is None:
        return lambda y: Ramp(y, length, fn)
    x = Signal(x)
    return RampOff(RampOn(x, length, fn), length, fn)


def FadeTo(*args):
    """src/ramps.jl:261-273"""
    args = list(args)
    if len(args) >= 2 and not (_isnumber(args[1]) or isinstance(args[1], Quantity)
                               or (callable(args[1]) and not isinstance(args[1], AbstractSignal))):
        x, y = args[0], args[1]
        _, length, fn = _ramp_args(args[2:])
    else:
        y = args[0]
        _, length, fn = _ramp_args(args[1:])
        return lambda x: FadeTo(x, y, length, fn)
    x, y = Uniform((x, y))
    if x.framerate is None:
        raise SignalError("Unknown frame rate is not supported by `FadeTo`.")
    n = inframes(maybeseconds(length), x.framerate)
    silence = Until(Signal(np.dtype(y.sampletype).type(0)),
                    Quantity(x.nframes - n, "frames", "frames"))
    return Mix(RampOff(x, length, fn), Append(silence, RampOn(y, length, fn)))


# ----------------------------------------------------------------------------
# ToFramerate and its push-down rules (src/reformatting.jl:64-122 + per node)

def maybe_rationalize(r):
    """src/reformatting.jl:103-111"""
    q = dspjl.rationalize(r)
    return q if q is not None else r


def _resample(x, fs, blocksize):
    """src/reformatting.jl:113-122 (`__ToFramerate__`)"""
    ratio = maybe_rationalize(fs / x.framerate)
    if ratio == 1:
        return x
    return FilteredSignal(x, ResamplerFn(ratio, fs), blocksize, fs)


def ToFramerate(*args, blocksize=default_blocksize):
    if len(args) == 1:
        fs = args[0]
        return lambda x: ToFramerate(x, fs, blocksize=blocksize)
    x, fs = args
    if not isinstance(x, AbstractSignal):
        x = _coerce_no_rate(x)
    fs = None if fs is None else float(inHz(fs))
    if fs is None and x.framerate is None:
        return x
    if fs is not None and x.framerate is not None and fs == x.framerate:
        return x
    if fs is None:
        return x                                           # reformatting.jl:85-86
    if x.framerate is None:
        return _toframerate_missing(x, fs, blocksize)
    if x.evaltrait == "data":
        return _resample(x, fs, blocksize)                 # reformatting.jl:88-90
    return _toframerate_computed(x, fs, blocksize)


def _coerce_no_rate(x):
    return Signal(x)


def _toframerate_missing(x, fs, bs):
    """The `IsSignal{<:Any,Missing}` methods."""
    if isinstance(x, ArraySignal):
        return ArraySignal(x.data, fs)                                     # arrays.jl:47-48
    if isinstance(x, NumberSignal):
        return NumberSignal(x.val, fs, x.dB)                               # numbers.jl:56-57
    if isinstance(x, SignalFunction):
        return SignalFunction(x.fn, x.first, x.omega, x.phi, fs)           # functions.jl:62-63
    if isinstance(x, CutApply):
        return CutApply(ToFramerate(x.signal, fs, blocksize=bs), x.time, x.kind)   # cutting.jl:147-152
    if isinstance(x, PaddedSignal):
        return PaddedSignal(ToFramerate(x.signal, fs, blocksize=bs), x.pad)        # padding.jl:18-19
    if isinstance(x, AppendSignals):
        return Append(*[ToFramerate(s, fs, blocksize=bs) for s in x.signals])      # appending.jl:79-80
    if isinstance(x, FilteredSignal):
        return FilteredSignal(ToFramerate(x.signal, fs, blocksize=bs), x.fn, x.blocksize, fs)  # filters.jl:153-157
    if isinstance(x, NormedSignal):
        return NormedSignal(ToFramerate(x.signal, fs, blocksize=bs))               # filters.jl:281-285
    if isinstance(x, MapSignal):
        return OperateOn(x.fn, *[ToFramerate(s, fs, blocksize=bs) for s in x.signals],
                         padding=x.padding, bychannel=x.bychannel, blocksize=x.blocksize)  # mapsignal.jl:62-64
    if isinstance(x, RampSignal):
        return RampSignal(x.direction, ToFramerate(x.signal, fs, blocksize=bs), x.time, x.fn)  # ramps.jl:37-43
    _nosignal(x)


def _toframerate_computed(x, fs, bs):
    """The `(::IsSignal{<:Any,<:Number}, ::ComputedSignal)` methods."""
    if isinstance(x, NumberSignal):
        return NumberSignal(x.val, fs, x.dB)
    if isinstance(x, SignalFunction):
        return SignalFunction(x.fn, x.first, x.omega, x.phi, fs)
    if isinstance(x, CutApply):      # only Until reaches here (After is data)
        t = _stretchtime(x.time, fs / x.framerate)                                # cutting.jl:142-146
        return CutApply(ToFramerate(x.signal, fs, blocksize=bs), t, "until")
    if isinstance(x, PaddedSignal):
        return PaddedSignal(ToFramerate(x.signal, fs, blocksize=bs), x.pad)       # padding.jl:16-17
    if isinstance(x, AppendSignals):
        return Append(*[ToFramerate(s, fs, blocksize=bs) for s in x.signals])     # appending.jl:77-78
    if isinstance(x, FilteredSignal):
        if x.framerate == x.signal.framerate:                                     # filters.jl:143-152
            return FilteredSignal(ToFramerate(x.signal, fs, blocksize=bs), x.fn, x.blocksize, fs)
        return _resample(x.signal, fs, bs)
    if isinstance(x, NormedSignal):
        return NormedSignal(ToFramerate(x.signal, fs, blocksize=bs))              # filters.jl:276-280
    if isinstance(x, MapSignal):
        if fs < x.framerate:                                                      # mapsignal.jl:46-58
            return OperateOn(x.fn, *[ToFramerate(s, fs, blocksize=bs) for s in x.signals],
                             padding=x.padding, bychannel=x.bychannel, blocksize=x.blocksize)
        return _resample(x, fs, bs)
    if isinstance(x, RampSignal):
        return RampSignal(x.direction, ToFramerate(x.signal, fs, blocksize=bs), x.time, x.fn)  # ramps.jl:28-36
    _nosignal(x)


# ----------------------------------------------------------------------------
# sink result-type selection (src/sink.jl:28-50)

def _mergepriority(r):
    if isinstance(r, ArraySignal):
        return 1
    return 0


def _mergeroots(roots):
    best = roots[0]
    for r in roots[1:]:
        if _mergepriority(r) > _mergepriority(best):
            best = r
    return best


def result_wants_tuple(x):
    """refineroot(root(x)) (src/sink.jl:31-37): a plain array root with unknown
    rate -> bare Array; everything else -> (Array, framerate)."""
    r = x.root()
    if isinstance(r, ArraySignal) and r.framerate is None:
        return False
    return True


def process_sink_params(x):
    """src/sink.jl:94-99"""
    x = Signal(x)
    if x.nframes is None:
        raise SignalError("Unknown number of frames in signal.")
    if isknowninf(x.nframes):
        raise SignalError("Cannot store infinite signal.")
    return x
