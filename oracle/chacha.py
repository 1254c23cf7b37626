"""ChaCha20 RNG and the arkworks sampling routines built on it.

Test infrastructure (see oracle/__init__.py).  Restates
* `rand_chacha 0.3.1::ChaChaRng` (Cargo.lock:2927-2928 of the reference): ChaCha20,
  256-bit seed as key, 64-bit block counter from 0, 64-bit stream id 0; output is
  the keystream read as little-endian u32 words; `next_u64` = low word then high word
  (the 4-block buffering of `BlockRng` never skips words);
* `setup_utils::derive_rng_from_seed` — reached from src/bin/contribute.rs:789,
  src/bin/verify_transcript.rs:675, src/bin/control.rs:791 — as
  `ChaChaRng::from_seed(seed[..32])` [UP];
* `ark-ff 0.4.2` `Fp::rand` and `ark-ec 0.4.2` `Projective::rand`.
"""
from __future__ import annotations

import struct

from .curves import Group
from .fields import Fp

_M32 = 0xFFFFFFFF


def _rotl(v, n):
    return ((v << n) & _M32) | (v >> (32 - n))


def _qr(s, a, b, c, d):
    s[a] = (s[a] + s[b]) & _M32; s[d] = _rotl(s[d] ^ s[a], 16)
    s[c] = (s[c] + s[d]) & _M32; s[b] = _rotl(s[b] ^ s[c], 12)
    s[a] = (s[a] + s[b]) & _M32; s[d] = _rotl(s[d] ^ s[a], 8)
    s[c] = (s[c] + s[d]) & _M32; s[b] = _rotl(s[b] ^ s[c], 7)


def chacha20_block(key_words, counter: int, stream: int = 0):
    init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + [
        counter & _M32, (counter >> 32) & _M32, stream & _M32, (stream >> 32) & _M32]
    s = list(init)
    for _ in range(10):
        _qr(s, 0, 4, 8, 12); _qr(s, 1, 5, 9, 13); _qr(s, 2, 6, 10, 14); _qr(s, 3, 7, 11, 15)
        _qr(s, 0, 5, 10, 15); _qr(s, 1, 6, 11, 12); _qr(s, 2, 7, 8, 13); _qr(s, 3, 4, 9, 14)
    return [(x + y) & _M32 for x, y in zip(s, init)]


class ChaChaRng:
    def __init__(self, seed: bytes):
        assert len(seed) == 32
        self.key = struct.unpack("<8I", seed)
        self.counter = 0
        self.buf: list[int] = []

    def next_u32(self) -> int:
        if not self.buf:
            self.buf = chacha20_block(self.key, self.counter)
            self.counter += 1
        return self.buf.pop(0)

    def next_u64(self) -> int:
        lo = self.next_u32()
        hi = self.next_u32()
        return lo | (hi << 32)

    def gen_bool(self) -> bool:
        # rand 0.8 `Standard` for bool: sign bit of one u32
        return bool(self.next_u32() >> 31)

    def fill_bytes(self, n: int) -> bytes:
        out = b""
        while len(out) < n:
            out += struct.pack("<I", self.next_u32())
        return out[:n]


def derive_rng_from_seed(seed: bytes) -> ChaChaRng:
    return ChaChaRng(bytes(seed[:32]))


def fp_rand(F: Fp, rng: ChaChaRng) -> int:
    """ark-ff `Fp::rand`: sample limbs, shave the top bits, read them AS the Montgomery
    representation, reject if >= p.  Returns the canonical value (limbs * R^-1 mod p)."""
    n = F.limbs64
    shave = 64 * n - F.bits
    rinv = pow(F.R, -1, F.p)
    while True:
        limbs = [rng.next_u64() for _ in range(n)]
        limbs[-1] &= (1 << (64 - shave)) - 1 if shave < 64 else 0
        v = sum(l << (64 * i) for i, l in enumerate(limbs))
        if v < F.p:
            return v * rinv % F.p


def field_rand(F, rng: ChaChaRng):
    if F.deg == 1:
        return fp_rand(F, rng)
    return tuple(fp_rand(F.base, rng) for _ in range(F.deg))


def group_rand(G: Group, rng: ChaChaRng):
    """ark-ec `Projective::rand`: x <- F::rand, greatest <- bool, retry until on curve,
    then clear the cofactor by multiplication."""
    while True:
        x = field_rand(G.F, rng)
        greatest = rng.gen_bool()
        P = G.point_from_x(x, greatest)
        if P is not None:
            return G.mul(P, G.cofactor)
