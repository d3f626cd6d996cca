"""
Philox4x32-10 and the per-chain draw layout of libgeosss_cuda, restated in
numpy (oracle; test infrastructure only).

The reference draws from numpy's PCG64 (geosss/mcmc.py:45); a counter-based
stream cannot reproduce that sequence, so free-running GPU chains are checked by
feeding the SAME Philox draws to the reference algorithm on the CPU:
``PhiloxChainRNG`` exposes the three calls the samplers make and yields, call by
call, exactly what chain ``chain`` of the CUDA kernels consumes.

Layout (DESIGN.md "RNG"): key = (seed lo, seed hi); counter =
(chain lo, chain hi, step lo, (step hi16 << 16) | block).  Within one transition
blocks 0..nb-1, nb = ceil(d/4), give the d normals (Box-Muller on word pairs (0,1)
and (2,3)).  The slice-level and shrink-bracket uniforms ("aux") are words 2,3 of
block nb-1 when d % 4 is 1 or 2 (that pair would be unused), else words 0,1 of
block nb.  Proposals start at block pb (= nb, resp. nb + 1): proposal j is word
j % 4 of block pb + j // 4.  The reject sampler skips the bracket word.
"""

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: 4 ints, key: 2 ints -> 4 uint32 words (Salmon et al. 2011)."""
    c = [np.uint64(int(v) & 0xFFFFFFFF) for v in counter]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return [int(v) for v in c]


def u01(word, dtype):
    """uint32 -> uniform strictly inside (0,1).
    float32 path: top 23 bits, ((w >> 9) + 0.5) * 2^-23 (exact in fp32).
    float64 path: all 32 bits, (w + 0.5) * 2^-32."""
    if np.dtype(dtype) == np.float32:
        return ((word >> 9) + 0.5) * 2.0**-23
    return (word + 0.5) * 2.0**-32


def box_muller(wa, wb, dtype):
    u1, u2 = u01(wa, dtype), u01(wb, dtype)
    r = np.sqrt(-2.0 * np.log(u1))
    return r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)


class PhiloxChainRNG:
    """The draws of one chain, served through the reference's RNG calls.

    ``standard_normal(d)`` starts a new transition (step += 1 after the first);
    ``random()`` is the slice-level uniform; ``uniform(a, b)`` serves the shrink
    bracket draw and then the proposals, in the layout above.
    """

    def __init__(self, seed, chain, kind, d, dtype=np.float64, first_step=0):
        self.key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        self.chain, self.kind, self.d = int(chain), int(kind), int(d)
        self.dtype = np.dtype(dtype)
        self.step = first_step - 1
        self.nb = (d + 3) // 4
        self.aux_in_last = d % 4 in (1, 2)
        self.pb = self.nb if self.aux_in_last else self.nb + 1
        self._cache = {}
        self.prop = 0
        self.bracket_pending = False

    def _block(self, b):
        got = self._cache.get(b)
        if got is None:
            assert b < 65536
            ctr = (self.chain & 0xFFFFFFFF, self.chain >> 32,
                   self.step & 0xFFFFFFFF, ((self.step >> 32) << 16) | b)
            got = self._cache[b] = philox4x32_10(ctr, self.key)
        return got

    def _aux(self, i):
        return self._block(self.nb - 1)[2 + i] if self.aux_in_last else self._block(self.nb)[i]

    def standard_normal(self, n):
        assert n == self.d
        self.step += 1
        self._cache = {}
        self.prop = 0
        self.bracket_pending = self.kind == 1
        z = np.empty(4 * self.nb)
        for b in range(self.nb):
            w = self._block(b)
            z[4 * b], z[4 * b + 1] = box_muller(w[0], w[1], self.dtype)
            z[4 * b + 2], z[4 * b + 3] = box_muller(w[2], w[3], self.dtype)
        return z[:n]

    def random(self):
        return u01(self._aux(0), self.dtype)

    def _next_uniform(self):
        if self.bracket_pending:
            self.bracket_pending = False
            return u01(self._aux(1), self.dtype)
        j = self.prop
        self.prop += 1
        return u01(self._block(self.pb + j // 4)[j % 4], self.dtype)

    def uniform(self, a, b):
        return a + (b - a) * self._next_uniform()


class PhiloxSeqRNG:
    """Sequential reader of one chain's Philox stream, as the baseline-sampler kernels
    (csrc/kernels_mh.cuh: PhiloxSeq) consume it: within a transition the words of blocks 0, 1, ...
    are used in order; ``random()`` takes one word, normals come in Box-Muller pairs (two words, the
    second normal is kept for the next request), ``gamma(d/2)`` is half the sum of squares of ``d``
    normals.  ``begin_step()`` moves to the next transition and drops the kept normal."""

    def __init__(self, seed, chain, dtype=np.float64, first_step=0):
        self.key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        self.chain = int(chain)
        self.dtype = np.dtype(dtype)
        self.step = first_step - 1
        self.begin_step()
        self.step = first_step - 1

    def begin_step(self):
        self.step += 1
        self.word, self.block, self.spare = 0, None, None

    def _next_word(self):
        b, i = divmod(self.word, 4)
        if i == 0 or self.block is None:
            assert b < 65536
            ctr = (self.chain & 0xFFFFFFFF, self.chain >> 32, self.step & 0xFFFFFFFF, ((self.step >> 32) << 16) | b)
            self.block = philox4x32_10(ctr, self.key)
        self.word += 1
        return self.block[i]

    def _normal(self):
        if self.spare is not None:
            z, self.spare = self.spare, None
            return z
        wa = self._next_word()
        wb = self._next_word()
        z, self.spare = box_muller(wa, wb, self.dtype)
        return z

    def standard_normal(self, n):
        return np.array([self._normal() for _ in range(n)], dtype=float)

    def random(self):
        return float(u01(self._next_word(), self.dtype))

    def gamma(self, shape):
        z = np.array([self._normal() for _ in range(int(round(2 * shape)))], dtype=self.dtype)
        acc = self.dtype.type(0)
        for v in z:  # same left-to-right accumulation as the kernel
            acc = acc + v * v
        return float(self.dtype.type(0.5) * acc)
