"""Counter-based normal noise for `Signal(randn; rng)` on the device (src/functions.jl:98-114).

The reference draws `randn(rng)` once per evaluated frame, in pull order, from a stateful generator: which
number a frame gets depends on block scheduling, and a stateful stream cannot be shared between thousands of
GPU threads.  `PhiloxRNG` is the stateless counterpart: frame k (1-based) of stream s is a pure function of
(seed, s, k) — Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; the same
generator cuRAND and torch use) followed by Box-Muller — so the device kernels, this numpy implementation (used by
the CPU sink / oracle and by the host when a leaf has to be materialised) and the plan interpreter of the
tests all produce the same numbers whatever the block, wave or device split.

    counter = (k_pair lo, k_pair hi, stream lo, stream hi),  key = (seed lo, seed hi),  k_pair = (k - 1) >> 1
    u1 = (x0 + x1 * 2^32 >> 11 + 0.5) * 2^-53,  u2 likewise from (x2, x3)          (both in (0, 1))
    r = sqrt(-2 ln u1);  frame k odd -> r cos(2 pi u2),  even -> r sin(2 pi u2)

The device side is `randn_value` in csrc/interp.cuh (leaf SIGOPS_LEAF_RANDN).
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """Philox4x32 with 10 rounds.  counter: (..., 4) uint32, key: (..., 2) uint32 (broadcastable) -> (..., 4) uint32."""
    c = np.asarray(counter, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    c0, c1, c2, c3 = (c[..., i].astype(np.uint64) for i in range(4))
    k0 = np.broadcast_to(k[..., 0], c[..., 0].shape).astype(np.uint32)
    k1 = np.broadcast_to(k[..., 1], c[..., 0].shape).astype(np.uint32)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0
            p1 = _M1 * c2
            hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
            hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
            n0 = hi1 ^ c1 ^ k0.astype(np.uint64)
            n2 = hi0 ^ c3 ^ k1.astype(np.uint64)
            c0, c1, c2, c3 = n0, lo1, n2, lo0
            k0 = (k0 + _W0).astype(np.uint32)
            k1 = (k1 + _W1).astype(np.uint32)
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def _unit(lo, hi):
    """53 random bits -> a double in (0, 1): ((hi:lo) >> 11 + 0.5) * 2^-53."""
    x = (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
    return ((x >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


class PhiloxRNG:
    """`rng` argument of `Signal(randn, rng=PhiloxRNG(seed))`: noise generated on the device.  In a batch, element i
    of the call must carry `stream = stream of element 0 + i` (every signal gets its own stream)."""

    def __init__(self, seed=0, stream=0):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.stream = int(stream)

    def frames(self, k_lo, k_hi, stream=None):
        """Frames k_lo .. k_hi - 1 (1-based, like the reference's frame index) of the stream, Float64."""
        s = self.stream if stream is None else int(stream)
        k = np.arange(int(k_lo), int(k_hi), dtype=np.int64)
        if k.size == 0:
            return np.empty(0)
        pair = ((k - 1) >> 1).astype(np.uint64)
        ctr = np.empty((k.size, 4), dtype=np.uint32)
        ctr[:, 0] = (pair & _MASK).astype(np.uint32)
        ctr[:, 1] = (pair >> np.uint64(32)).astype(np.uint32)
        su = np.uint64(s & 0xFFFFFFFFFFFFFFFF)
        ctr[:, 2] = np.uint32(su & _MASK)
        ctr[:, 3] = np.uint32(su >> np.uint64(32))
        key = np.array([self.seed & 0xFFFFFFFF, self.seed >> 32], dtype=np.uint32)
        x = philox4x32_10(ctr, key)
        u1, u2 = _unit(x[:, 0], x[:, 1]), _unit(x[:, 2], x[:, 3])
        r = np.sqrt(-2.0 * np.log(u1))
        ang = 2.0 * u2                      # in units of pi, like the device's sincospi
        odd = (k & 1) == 1
        return np.where(odd, r * np.cos(np.pi * ang), r * np.sin(np.pi * ang))

    def standard_normal(self, n):
        """numpy.Generator-style draw (frames 1..n of the stream): lets the object stand in wherever the host code
        expects a numpy generator."""
        return self.frames(1, int(n) + 1)
