"""Counter-based randomness shared verbatim by the oracle and the CUDA path (SURVEY §8d "Randomness").

TEST INFRASTRUCTURE for the oracle side; the same arithmetic is restated in
ctrm_b200/csrc/rng.cuh for the device.  The reference consumes two global streams in a
data-dependent order (numpy MT19937 for the gate / random-walk draws, learned_sampler.py:93,:137-138,
:150-155; torch's CPU generator for the Gumbel uniforms, cvae.py:133-136).  A GPU cannot reproduce a
stream order, so every draw is keyed instead:

    Philox4x32-10(key=(seed_lo, seed_hi), ctr=(instance, k_traj<<16 | t, index, stream<<8 | sub))

    stream 0: gate draw u in [0,1)           index = agent                (learned_sampler.py:150-155)
    stream 1: random walk (u_mag, u_theta)   index = agent, sub = attempt (learned_sampler.py:137-138)
    stream 2: per-step draw                  index = 0                    (learned_sampler.py:93)
    stream 3: randomised indicator           index = row  (word0 % dim_ind)   (learned_sampler.py:94-97)
    stream 8: Gumbel uniforms, 4 per call    index = row (k*N+i), sub = column/4  (cvae.py:133-136)

f64 uniforms: ((w0<<32 | w1) >> 11) * 2^-53 ; f32 uniforms: (w >> 8) * 2^-24 (what torch.rand does on
CPU for float32).

sin/cos of the random-walk heading: numpy's libm/SVML values are not reproducible on a GPU (and differ
between CPUs), so table mode uses det_sincos_2pi below: exact quadrant reduction of u, one rounded
multiply by pi/2, Taylor polynomials evaluated with individually rounded * and + (no FMA).  It agrees
with np.sin/np.cos to ~1 ulp; the stream-mode oracle (pinned against the reference) keeps np.sin/np.cos.
"""
from __future__ import annotations

import math
import struct

import numpy as np

PHILOX_M0 = 0xD2511F53
PHILOX_M1 = 0xCD9E8D57
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85

STREAM_GATE = 0
STREAM_RW = 1
STREAM_STEP = 2
STREAM_RANDIND = 3
STREAM_GUMBEL = 8

_M32 = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1):
    """vectorised Philox4x32-10; all inputs broadcastable uint32-valued arrays; returns 4 uint32 arrays"""
    c0 = np.asarray(c0, dtype=np.uint64) & _M32
    c1 = np.asarray(c1, dtype=np.uint64) & _M32
    c2 = np.asarray(c2, dtype=np.uint64) & _M32
    c3 = np.asarray(c3, dtype=np.uint64) & _M32
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint64(k0 & 0xFFFFFFFF)
    k1 = np.uint64(k1 & 0xFFFFFFFF)
    m0 = np.uint64(PHILOX_M0)
    m1 = np.uint64(PHILOX_M1)
    s32 = np.uint64(32)
    for _ in range(10):
        p0 = m0 * c0
        p1 = m1 * c2
        hi0, lo0 = p0 >> s32, p0 & _M32
        hi1, lo1 = p1 >> s32, p1 & _M32
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(PHILOX_W0)) & _M32
        k1 = (k1 + np.uint64(PHILOX_W1)) & _M32
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def _ctr1(k_traj: int, t: int) -> int:
    return ((k_traj & 0xFFFF) << 16) | (t & 0xFFFF)


def _ctr3(stream: int, sub: int) -> int:
    return ((stream & 0xFF) << 8) | (sub & 0xFF)


def u64_to_f64(w0, w1):
    x = (np.asarray(w0, dtype=np.uint64) << np.uint64(32)) | np.asarray(w1, dtype=np.uint64)
    return (x >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def u32_to_f32(w):
    return ((np.asarray(w, dtype=np.uint32) >> np.uint32(8)).astype(np.float32)) * np.float32(2.0 ** -24)


# ---------------------------------------------------------------- deterministic sin/cos

def _bits(x: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", x))[0]


HALF_PI = math.pi / 2.0  # 0x3FF921FB54442D18
SIN_COEF = [(-1.0) ** k / math.factorial(2 * k + 1) for k in range(9)]   # x^(2k+1)
COS_COEF = [(-1.0) ** k / math.factorial(2 * k) for k in range(10)]      # x^(2k)
SIN_COEF_BITS = [_bits(c) for c in SIN_COEF]
COS_COEF_BITS = [_bits(c) for c in COS_COEF]


def det_sincos_2pi(u: float):
    """(sin(2*pi*u), cos(2*pi*u)) for u in [0,1), bit-reproducible with plain IEEE * and + (no FMA)."""
    u = float(u)
    a = u * 4.0                 # exact
    q = int(a)                  # 0..3
    f = a - q                   # exact, [0,1)
    swap = f > 0.5
    g = (1.0 - f) if swap else f   # exact; [0, 0.5]
    x = g * HALF_PI             # one rounding; [0, pi/4]
    x2 = x * x
    s = SIN_COEF[8]
    for k in range(7, -1, -1):
        s = s * x2
        s = s + SIN_COEF[k]
    s = s * x
    c = COS_COEF[9]
    for k in range(8, -1, -1):
        c = c * x2
        c = c + COS_COEF[k]
    if swap:
        s, c = c, s
    if q == 0:
        return s, c
    if q == 1:
        return c, -s
    if q == 2:
        return -s, -c
    return -c, s


class PhiloxTables:
    """table-mode randomness for ONE instance (instance id is part of the counter)"""

    mode = "table"

    def __init__(self, seed: int, instance: int):
        self.k0 = seed & 0xFFFFFFFF
        self.k1 = (seed >> 32) & 0xFFFFFFFF
        self.instance = instance & 0xFFFFFFFF

    def _f64(self, k_traj, t, index, stream, sub, second=False):
        w = philox4x32(self.instance, _ctr1(k_traj, t), index, _ctr3(stream, sub), self.k0, self.k1)
        if second:
            return float(u64_to_f64(w[2], w[3]))
        return float(u64_to_f64(w[0], w[1]))

    def step_draw(self, k_traj, t):
        return self._f64(k_traj, t, 0, STREAM_STEP, 0)

    def gate(self, k_traj, t, agent):
        return self._f64(k_traj, t, agent, STREAM_GATE, 0)

    def random_walk(self, k_traj, t, agent, attempt):
        """returns (u_mag, sin, cos)"""
        w = philox4x32(self.instance, _ctr1(k_traj, t), agent, _ctr3(STREAM_RW, attempt), self.k0, self.k1)
        u_mag = float(u64_to_f64(w[0], w[1]))
        u_th = float(u64_to_f64(w[2], w[3]))
        s, c = det_sincos_2pi(u_th)
        return u_mag, s, c

    def rand_indicator(self, k_traj, t, rows, dim_ind):
        w = philox4x32(self.instance, _ctr1(k_traj, t), np.arange(rows), _ctr3(STREAM_RANDIND, 0), self.k0, self.k1)
        return (w[0] % np.uint32(dim_ind)).astype(np.int64)

    def gumbel_uniforms(self, k_traj, t, rows, dim_latent):
        assert dim_latent % 4 == 0
        r = np.arange(rows).reshape(-1, 1)
        sub = np.arange(dim_latent // 4).reshape(1, -1)
        c3 = (np.uint64(STREAM_GUMBEL) << np.uint64(8)) | sub.astype(np.uint64)
        w = philox4x32(self.instance, _ctr1(k_traj, t), r, c3, self.k0, self.k1)
        out = np.stack([u32_to_f32(x) for x in w], axis=-1)  # (rows, L/4, 4)
        return out.reshape(rows, dim_latent)


class StreamRNG:
    """stream-mode randomness: the reference's own global generators in the reference's own call order.
    Only used to pin the oracle against the reference (oracle/gen_golden.py, tests)."""

    mode = "stream"

    def step_draw(self, k_traj, t):
        return float(np.random.rand())

    def gate(self, k_traj, t, agent):
        return float(np.random.rand())

    def random_walk(self, k_traj, t, agent, attempt):
        u_mag = np.random.rand()
        theta = np.random.rand() * np.pi * 2
        return u_mag, np.sin(theta), np.cos(theta)

    def rand_indicator(self, k_traj, t, rows, dim_ind):
        import torch

        return torch.randint(0, dim_ind, (rows,)).numpy()

    def gumbel_uniforms(self, k_traj, t, rows, dim_latent):
        import torch

        return torch.rand((rows, dim_latent), dtype=torch.float32).numpy()


class ReplayRNG:
    """replays a recorded reference run: np.random.rand draws in call order, the np.sin/np.cos values the
    random walk took (libm/SVML results are machine-dependent), and optionally the torch.rand uniforms
    per step (torch's generator is also consumed by nn.Linear initialisation inside `reconstruct`, so a
    seed alone does not reproduce them)."""

    mode = "stream"

    def __init__(self, draws, sincos=None, uniforms=None):
        self.draws = np.asarray(draws, dtype=np.float64)
        self.sincos = None if sincos is None else np.asarray(sincos, dtype=np.float64)
        self.uniforms = uniforms
        self.i = 0
        self.j = 0
        self.s = 0

    def _next(self):
        v = float(self.draws[self.i])
        self.i += 1
        return v

    def step_draw(self, k_traj, t):
        return self._next()

    def gate(self, k_traj, t, agent):
        return self._next()

    def random_walk(self, k_traj, t, agent, attempt):
        u_mag = self._next()
        u_th = self._next()
        if self.sincos is None:
            theta = u_th * np.pi * 2
            return u_mag, np.sin(theta), np.cos(theta)
        s, c = float(self.sincos[self.j]), float(self.sincos[self.j + 1])
        self.j += 2
        return u_mag, s, c

    def rand_indicator(self, k_traj, t, rows, dim_ind):
        raise NotImplementedError("replay of randomize_indicator is not recorded")

    def gumbel_uniforms(self, k_traj, t, rows, dim_latent):
        if self.uniforms is None:
            return np.full((rows, dim_latent), 0.5, dtype=np.float32)  # teacher-forced runs ignore the NN
        u = self.uniforms[self.s]
        self.s += 1
        return u
