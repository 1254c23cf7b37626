"""`Phase1Parameters` size / offset arithmetic.

Test infrastructure (see oracle/__init__.py).  Restates `phase1::Phase1Parameters`
as constructed by the reference at src/utils.rs:326-352 (`new_chunk`, `new_full`)
and src/bin/new_setup.rs:95-102, 265-277 (number of chunks), with the size formulas
of SURVEY.md Appendix A.4.  Groth16 proving system only (the ceremony's).

Chunk file layout (both challenge and response):
    hash[64] || tauG1[g1n] || tauG2[on] || alphaG1[on] || betaG1[on] || betaG2[1] (|| pubkey)
"""
from __future__ import annotations

from dataclasses import dataclass

from .curves import Curve, get_curve

HASH_SIZE = 64


@dataclass(frozen=True)
class Phase1Params:
    curve: Curve
    power: int
    chunk_index: int = 0
    chunk_size: int = 0          # 0 = full mode
    batch_size: int = 0

    # -- reference constructors ---------------------------------------------------------
    @staticmethod
    def new_chunk(curve, chunk_index: int, chunk_size: int, power: int, batch_size: int):
        c = get_curve(curve) if isinstance(curve, str) else curve
        return Phase1Params(c, power, chunk_index, chunk_size, batch_size)

    @staticmethod
    def new_full(curve, power: int, batch_size: int):
        c = get_curve(curve) if isinstance(curve, str) else curve
        return Phase1Params(c, power, 0, 0, batch_size)

    # -- derived quantities ----------------------------------------------------------------
    @property
    def powers_length(self) -> int:
        return 1 << self.power

    @property
    def powers_g1_length(self) -> int:
        return (1 << (self.power + 1)) - 1

    @property
    def full(self) -> bool:
        return self.chunk_size == 0

    @property
    def start(self) -> int:
        return 0 if self.full else self.chunk_index * self.chunk_size

    @property
    def g1_count(self) -> int:
        if self.full:
            return self.powers_g1_length
        return max(0, min(self.chunk_size, self.powers_g1_length - self.start))

    @property
    def other_count(self) -> int:
        if self.full:
            return self.powers_length
        return max(0, min(self.chunk_size, self.powers_length - self.start))

    @property
    def num_chunks(self) -> int:
        # src/bin/new_setup.rs:274-276
        cs = self.chunk_size
        return (self.powers_g1_length + cs - 1) // cs

    def sizes(self, compressed: bool):
        g1 = self.curve.g1.F.nbytes * (1 if compressed else 2)
        g2 = self.curve.g2.F.nbytes * (1 if compressed else 2)
        return g1, g2

    @property
    def public_key_size(self) -> int:
        g1u, g2u = self.sizes(False)
        return 3 * g2u + 6 * g1u

    def _size(self, compressed: bool) -> int:
        g1, g2 = self.sizes(compressed)
        return self.g1_count * g1 + self.other_count * (g2 + 2 * g1) + g2 + HASH_SIZE

    @property
    def accumulator_size(self) -> int:
        """challenge file (uncompressed)"""
        return self._size(False)

    @property
    def contribution_size(self) -> int:
        """response file (compressed + public key)"""
        return self._size(True) + self.public_key_size

    def offsets(self, compressed: bool):
        """byte offsets of (tauG1, tauG2, alphaG1, betaG1, betaG2, end) in a chunk file"""
        g1, g2 = self.sizes(compressed)
        o_tau1 = HASH_SIZE
        o_tau2 = o_tau1 + self.g1_count * g1
        o_al = o_tau2 + self.other_count * g2
        o_be = o_al + self.other_count * g1
        o_b2 = o_be + self.other_count * g1
        return o_tau1, o_tau2, o_al, o_be, o_b2, o_b2 + g2
