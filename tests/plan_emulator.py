"""Test-only numpy interpreter of the plan byte format of include/signalops.h.

It decodes exactly the bytes `libsignalops_cuda.so` receives and evaluates them
with straightforward numpy, so the host-side lowering (offsets, lengths, pads,
stage order, fusion) can be checked against the oracle on a machine without a
GPU.  It is not a fallback: nothing in the product imports it.
"""
import math
import struct

import numpy as np
from scipy import signal as sps

F32, F64, I64 = 1, 2, 3
NP = {F32: np.float32, F64: np.float64, I64: np.int64}


def _sinpi(x):
    r = np.fmod(x, 2.0)
    r = np.where(r > 1.0, r - 2.0, np.where(r < -1.0, r + 2.0, r))
    r = np.where(r > 0.5, 1.0 - r, np.where(r < -0.5, -1.0 - r, r))
    return np.sin(np.pi * r)


def _fn(fn, x, a, b):
    return {1: np.sin, 2: np.cos, 3: lambda v: v / math.pi - 1.0, 4: lambda v: a * np.sin(v) + b,
            5: lambda v: a * np.cos(v) + b, 6: lambda v: v, 7: lambda v: _sinpi(0.5 * v)}[fn](x)


class Emulator:
    def __init__(self, blob):
        o = 0
        (magic, ver, self.n_in, self.n_tmp, self.n_out, self.n_scal, n_tab, n_ins, n_pc, n_st,
         n_dbl) = struct.unpack_from("<10IQ", blob, o)
        o += 48
        assert magic == 0x504F4753 and ver == 1
        nb = self.n_in + self.n_tmp + self.n_out
        self.bufdesc = [struct.unpack_from("<q2i", blob, o + 16 * k) for k in range(nb)]
        o += 16 * nb
        tabs = [struct.unpack_from("<2q", blob, o + 16 * k) for k in range(n_tab)]
        o += 16 * n_tab
        self.instrs = [struct.unpack_from("<4B3i3q5d", blob, o + 80 * k) for k in range(n_ins)]
        o += 80 * n_ins
        self.pieces = [struct.unpack_from("<2q4i", blob, o + 32 * k) for k in range(n_pc)]
        o += 32 * n_pc
        self.stages = [struct.unpack_from("<10i2q2id8iq2d", blob, o + 128 * k) for k in range(n_st)]
        o += 128 * n_st
        dbl = np.frombuffer(blob, dtype="<f8", count=n_dbl, offset=o)
        assert o + 8 * n_dbl == len(blob)
        self.tables = [dbl[a:a + c] for a, c in tabs]

    # ---- leaves / programs ---------------------------------------------------------
    def leaf(self, I, n, c, stage):
        (op, leaf, fn, flags, buf, c_mul, c_off, i0, i1, i2, d0, d1, d2, d3, d4) = I
        if leaf == 1:
            return np.full(n.shape, d0)
        if leaf == 2:
            idx = n + i0
            out = np.full(n.shape, d0, dtype=np.float64)
            pad = (flags >> 1) & 3
            ok = (idx >= 0) & (idx < i1)
            src = idx.copy()
            if pad and i1 > 0:
                hi = idx >= i1
                if pad == 1:
                    src[hi] = idx[hi] % i1
                elif pad == 2:
                    cnt, rem = np.divmod(idx[hi], i1)
                    src[hi] = np.where(cnt % 2 == 1, i1 - 1 - rem, rem)
                else:
                    src[hi] = i1 - 1
                ok = ok | hi
            b = self.bufs[buf]
            out[ok] = b[src[ok], c * c_mul + c_off]
            return out
        if leaf == 3:
            idx = n + i0
            ok = (idx >= 0) & (idx < i1)
            out = np.full(n.shape, d0, dtype=np.float64)
            b = self.bufs[buf].astype(np.float64)
            acc = b[idx[ok], 0].copy()
            for ch in range(1, i2):
                acc = acc + b[idx[ok], ch]
            out[ok] = acc
            return out
        if leaf == 4:
            k = (n + i0).astype(np.float64)
            t = k / d0
            if flags & 1:
                u = t * d1 + d2
                return _sinpi(2 * u) if fn == 1 else _fn(fn, 2 * math.pi * np.fmod(u, 1.0), d3, d4)
            return _sinpi(2 * (t + d2)) if fn == 1 else _fn(fn, t + d2, d3, d4)
        if leaf == 5:
            k = n + i0
            return np.where(k > i1, 1.0, _fn(fn, (k - 1) / float(i1), 0, 0))
        if leaf == 6:
            k = n + i0
            return np.where(k <= i1, 1.0, _fn(fn, 1.0 - (k - i1) / float(i2), 0, 0))
        if leaf == 7:
            return np.full(n.shape, math.sqrt(self.scalars[buf] / d0))
        if leaf == 8:
            return stage
        if leaf == 9:                      # LEAF_RANDN: Philox noise, stream = i2 + index of the instance in the call
            from signalops.philox import PhiloxRNG
            k = n + i0
            if k.size == 0:
                return np.zeros(0)
            assert np.all(np.diff(k) == 1)
            return PhiloxRNG(i1, i2 + self.inst).frames(int(k[0]), int(k[-1]) + 1)
        raise ValueError(leaf)

    def run_prog(self, start, ln, n, c, stage=None):
        acc, stack = None, []
        for I in self.instrs[start:start + ln]:
            op = I[0]
            if op <= 5:
                v = self.leaf(I, n, c, stage)
                acc = v if op == 1 else acc + v if op == 2 else acc - v if op == 3 else acc * v if op == 4 else acc / v
            elif op == 6:
                stack.append(acc)
            elif op <= 10:
                l = stack.pop()
                acc = l + acc if op == 7 else l - acc if op == 8 else l * acc if op == 9 else l / acc
            elif op == 11:
                acc = -acc
            elif op == 12:
                acc = acc.astype(np.float32).astype(np.float64)
            elif op == 13:
                acc = acc.astype(np.int64).astype(np.float64)
        return acc

    # ---- stages -----------------------------------------------------------------------
    def run(self, inputs, inst=0):
        self.inst = inst                  # index of this instance in the call (noise streams)
        # (WavRaw inputs — frame-interleaved host buffers, also what a C-ordered numpy matrix is passed as — are decoded
        #  the way the device's k_wav does: PCM16 / 32768, floats as they are)
        inputs = [a.decode().astype(np.float32 if a.raw.dtype == np.float32 else np.float64) if hasattr(a, "raw") else a
                  for a in inputs]
        self.bufs = [np.asarray(a).reshape(len(a), -1) for a in inputs]
        for k in range(self.n_in, len(self.bufdesc)):
            n, c, dt = self.bufdesc[k]
            self.bufs.append(np.zeros((n, c), dtype=NP[dt]))
        self.scalars = np.zeros(max(1, self.n_scal))
        for st in self.stages:
            (kind, out_buf, slot, p0, npc, in0, inl, ep0, epl, nch, n_in, n_out, M, ctab, gain,
             fkind, nphi, tapsper, pfb_t, dpfb_t, interp, decim, _r, deficit, rate, phase0) = st
            ob = self.bufs[out_buf]

            def store(n, c, v):
                w = v.astype(ob.dtype)
                ob[n, c] = w
                if slot >= 0:
                    self.scalars[slot] += float(np.sum(w.astype(np.float64) ** 2))

            if kind == 1:
                for (lo, ln, c0, cc, ps, pl) in self.pieces[p0:p0 + npc]:
                    n = np.arange(lo, lo + ln)
                    for c in range(c0, c0 + cc):
                        store(n, c, self.run_prog(ps, pl, n, c))
                continue
            for c in range(nch):
                x = self.run_prog(in0, inl, np.arange(n_in), c) if n_in else np.zeros(0)
                if kind == 2:
                    coef = self.tables[ctab].reshape(M, 5)
                    sos = np.column_stack([coef[:, :3], np.ones(M), coef[:, 3:]])
                    y = sps.sosfilt(sos, x) * gain
                else:
                    y = self.fir(st, x)
                n = np.arange(n_out)
                store(n, c, self.run_prog(ep0, epl, n, c, y) if epl else y)
        return self.bufs[self.n_in + self.n_tmp:]

    def fir(self, st, x):
        (kind, out_buf, slot, p0, npc, in0, inl, ep0, epl, nch, n_in, n_out, M, ctab, gain,
         fkind, nphi, T, pfb_t, dpfb_t, interp, decim, _r, deficit, rate, phase0) = st
        pfb = self.tables[pfb_t].reshape(nphi, T)
        dpfb = self.tables[dpfb_t].reshape(nphi, T) if dpfb_t >= 0 else None
        xp = np.concatenate([np.zeros(T), x, np.zeros(T + 8)])
        y = np.zeros(n_out)
        xi, acc, ph = deficit, phase0, int(phase0)
        for m in range(n_out):
            p = xi - 1                          # newest sample, 0-based
            lo = p - T + 1 + T                  # index into xp
            if lo + T > len(xp):
                xp = np.concatenate([xp, np.zeros(lo + T - len(xp) + 1024)])
            w = xp[lo:lo + T]
            if fkind == 1:
                phi = int(math.floor(acc))
                a = acc - phi
                y[m] = np.dot(pfb[phi - 1] + a * dpfb[phi - 1], w)
                acc += nphi / rate
                if acc > nphi:
                    xi += int(math.floor((acc - 1.0) / nphi))
                    acc = math.fmod(acc - 1.0, nphi) + 1.0
            elif fkind == 2:
                y[m] = np.dot(pfb[ph - 1], w)
                xi += (ph + decim - 1) // interp
                v = ph + decim % interp
                ph = v - interp if v > interp else v
            else:
                y[m] = np.dot(pfb[0], w)
                xi += decim
        return y
