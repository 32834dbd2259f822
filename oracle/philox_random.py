"""TEST INFRASTRUCTURE — not product code.

Counter-based Philox4x32-10 stream presented as a ``random.Random`` subclass.

Why it exists: the reference seeds the process-global Mersenne Twister
(simcore/simulator_paper_multi.py:71) and draws from it at arrivals.py:8,11,15,44
and simulator_paper_multi.py:576.  A GPU replica cannot carry 2.5 kB of MT state, so
"same seed" for the batched engine means "same Philox key".  This class is injected
into the *unmodified* reference (see oracle/ref_harness.py) so that reference, C oracle
and CUDA kernel consume one and the same word stream.

Stream definition (shared by oracle/dcsim_oracle.c and csrc/dcsim_kernel.cu):
  key      = (seed & 0xffffffff, seed >> 32)              seed taken mod 2**64
  block b  = Philox4x32-10(counter=(b & 0xffffffff, b >> 32, 0, 0), key) -> 4 words
  word  w  = block[w // 4][w % 4]                          consumed strictly in order
  random()        takes 2 words a,b -> ((a >> 5) * 2**26 + (b >> 6)) / 2**53   (CPython's own rule,
                  Modules/_randommodule.c genrand_res53)
  getrandbits(k)  k <= 32: takes 1 word -> word >> (32 - k)                     (CPython's rule for k<=32)
Everything else (expovariate, normalvariate, lognormvariate, choice/_randbelow) is inherited
from /usr/lib/python3.12/random.py unchanged, so draw order and counts are CPython's.
"""
import random as _random

import numpy as np

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    """Scalar Philox4x32-10 (Salmon et al., SC'11). ctr: 4 words, key: 2 words -> 4 words."""
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK32, p1 & MASK32, ((p0 >> 32) ^ c3 ^ k1) & MASK32, p0 & MASK32
        k0 = (k0 + W0) & MASK32
        k1 = (k1 + W1) & MASK32
    return c0, c1, c2, c3


def philox_blocks(seed, first_block, n_blocks):
    """Vectorised: words of blocks [first_block, first_block+n_blocks) as a flat uint32 array."""
    seed &= (1 << 64) - 1
    b = np.arange(first_block, first_block + n_blocks, dtype=np.uint64)
    c0 = b & np.uint64(MASK32)
    c1 = b >> np.uint64(32)
    c2 = np.zeros_like(b)
    c3 = np.zeros_like(b)
    k0 = seed & MASK32
    k1 = seed >> 32
    m32 = np.uint64(MASK32)
    s32 = np.uint64(32)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        c0, c1, c2, c3 = ((p1 >> s32) ^ c1 ^ np.uint64(k0)), p1 & m32, ((p0 >> s32) ^ c3 ^ np.uint64(k1)), p0 & m32
        k0 = (k0 + W0) & MASK32
        k1 = (k1 + W1) & MASK32
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.uint32).reshape(-1)


class PhiloxRandom(_random.Random):
    """random.Random whose bit source is the Philox word stream described in the module docstring."""

    CHUNK_BLOCKS = 2048

    def __init__(self, seed=0):
        self._seed = 0
        self._pos = 0          # words consumed so far
        self._buf = []
        self._buf_first = 0    # stream index of _buf[0]
        self.n_random = 0
        self.n_getrandbits = 0
        super().__init__(seed)

    # --- state ---------------------------------------------------------
    def seed(self, a=0, version=2):
        if not isinstance(a, int):
            raise TypeError("PhiloxRandom.seed wants an int")
        self._seed = a & ((1 << 64) - 1)
        self._pos = 0
        self._buf = []
        self._buf_first = 0
        self.n_random = 0
        self.n_getrandbits = 0

    def getstate(self):
        return (self._seed, self._pos)

    def setstate(self, state):
        self._seed, self._pos = state
        self._buf = []
        self._buf_first = 0

    @property
    def words_consumed(self):
        return self._pos

    # --- bit source ----------------------------------------------------
    def _word(self):
        i = self._pos - self._buf_first
        if i >= len(self._buf) or i < 0:
            first_block = self._pos // 4
            self._buf = philox_blocks(self._seed, first_block, self.CHUNK_BLOCKS).tolist()
            self._buf_first = first_block * 4
            i = self._pos - self._buf_first
        self._pos += 1
        return self._buf[i]

    def random(self):
        self.n_random += 1
        a = self._word() >> 5
        b = self._word() >> 6
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0)

    def getrandbits(self, k):
        if k < 0:
            raise ValueError("number of bits must be non-negative")
        self.n_getrandbits += 1
        if k == 0:
            return 0
        if k <= 32:
            return self._word() >> (32 - k)
        # CPython fills 32-bit chunks least-significant first; the top chunk keeps its high bits.
        out, shift = 0, 0
        while k > 0:
            w = self._word()
            if k < 32:
                w >>= (32 - k)
            out |= w << shift
            shift += 32
            k -= 32
        return out
