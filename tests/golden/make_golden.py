#!/usr/bin/env python3
"""Regenerates the fixtures in tests/golden/.

1. circuit_<curve>.bin — byte-for-byte copies of the reference's e2e R1CS fixtures
   (/root/reference/e2e/circuit_{bls12_377,bw6,mnt4_753,mnt6_753}); they are DATA, not source,
   and are the only files in the reference that pin a byte format on this path: ark-serialize
   `Matrices` whose Fr coefficients are 32/48/95/95-byte little-endian canonical integers
   (SURVEY.md §8c.1).  Copied only when /root/reference exists (this container).
2. p1_<curve>_c<k>.{challenge,response}.bin — oracle outputs for tiny phase-1 chunks
   (power 3, chunk size 4, chunk index k in {0, 3}: a full chunk and a G1-only tail chunk)
   with the synthetic accumulator of oracle/synth.py and the contributor scalars from
   ChaCha20(seed = bytes(range(32))); public key block zero-filled (RNG-free core).
   The GPU tests compare libsso_b200.so against these bytes, so the B200 box needs neither
   /root/reference nor minutes of Python big-int arithmetic.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import phase1, synth  # noqa: E402
from oracle.curves import CURVE_NAMES  # noqa: E402
from oracle.params import Phase1Params  # noqa: E402

REF = "/root/reference/e2e"
POWER, CHUNK = 3, 4


def main():
    if os.path.isdir(REF):
        for src, dst in (("circuit_bls12_377", "circuit_bls12_377.bin"), ("circuit_bw6", "circuit_bw6_761.bin"),
                         ("circuit_mnt4_753", "circuit_mnt4_753.bin"), ("circuit_mnt6_753", "circuit_mnt6_753.bin")):
            shutil.copyfile(os.path.join(REF, src), os.path.join(HERE, dst))
    for name in CURVE_NAMES:
        for k in (0, 3):
            p = Phase1Params.new_chunk(name, k, CHUNK, POWER, CHUNK)
            ch = synth.synthetic_challenge(p)
            key = synth.contributor_key(p.curve)
            resp = phase1.contribute_with_key(p, ch, key, bytes(p.public_key_size))
            base = os.path.join(HERE, "p1_%s_c%d" % (name, k))
            open(base + ".challenge.bin", "wb").write(ch)
            open(base + ".response.bin", "wb").write(resp)
            print(name, k, len(ch), len(resp))


if __name__ == "__main__":
    main()
