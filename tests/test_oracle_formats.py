"""Format pins.  The only byte formats the reference tree itself fixes on this path are
(1) the Fr encoding inside the e2e R1CS fixtures and (2) the 64-byte raw hash files
(SURVEY.md §8c).  Everything else (point flags etc.) is the published ark-serialize 0.4 format,
checked here for self-consistency and edge cases."""
import os
import struct

import pytest

from oracle import serialize as ser
from oracle.curves import CURVE_NAMES, get_curve
from oracle.phase1 import blank_hash, calculate_hash


def _parse_matrices(buf, fr_bytes):
    """ark-relations `Matrices` as written by ark-serialize 0.4: six u64 header fields, then three
    Vec<Vec<(Fr, u64)>> with u64 length prefixes."""
    off = 0

    def u64():
        nonlocal off
        v = struct.unpack_from("<Q", buf, off)[0]
        off += 8
        return v

    header = [u64() for _ in range(6)]
    mats = []
    for _ in range(3):
        rows = []
        for _ in range(u64()):
            row = []
            for _ in range(u64()):
                coeff = int.from_bytes(buf[off:off + fr_bytes], "little")
                off += fr_bytes
                row.append((coeff, u64()))
            rows.append(row)
        mats.append(rows)
    return header, mats, off


@pytest.mark.parametrize("name", CURVE_NAMES)
def test_fr_encoding_pinned_by_reference_fixture(name, golden_dir):
    c = get_curve(name)
    buf = open(os.path.join(golden_dir, "circuit_%s.bin" % name), "rb").read()
    header, mats, used = _parse_matrices(buf, c.Fr.nbytes)
    assert used == len(buf)                       # parses exactly only with the right |Fr|
    assert header == [2, 1, 6, 6, 4, 4]
    for m in mats:
        assert len(m) == 6
        for row in m:
            for coeff, col in row:
                assert coeff == 1 and col < 3     # canonical little-endian 1, not Montgomery
                assert ser.field_to_bytes(c.Fr, coeff) == (1).to_bytes(c.Fr.nbytes, "little")
    assert len(buf) == 48 + 168 + 14 * (c.Fr.nbytes + 8)


def test_hash_conventions():
    # src/utils.rs:618-623: Blake2b-512; new_setup.rs:200-201: the zero hash literal is 128 hex chars
    assert len(calculate_hash(b"abc")) == 64
    assert calculate_hash(b"abc").hex().startswith("ba80a53f981c4d0d6a2797b69f12f6e9")
    assert blank_hash().hex().startswith("786a02f742015903c6c6fd852552d272")


@pytest.mark.parametrize("name", CURVE_NAMES)
def test_point_roundtrip_and_flags(name):
    c = get_curve(name)
    for G in (c.g1, c.g2):
        P = G.mul(G.gen, 0xDEADBEEF)
        for Q in (P, G.neg(P), None):
            for comp in (True, False):
                b = ser.point_to_bytes(G, Q, comp)
                assert len(b) == ser.point_size(G, comp)
                assert G.eq(ser.point_from_bytes(G, b, comp), Q)
        # exactly one of P, -P carries the "negative" flag; infinity carries bit 6 and x = 0
        f1 = ser.point_to_bytes(G, P, True)[-1] & 0x80
        f2 = ser.point_to_bytes(G, G.neg(P), True)[-1] & 0x80
        assert f1 != f2
        inf = ser.point_to_bytes(G, None, True)
        assert inf[-1] == 0x40 and not any(inf[:-1])
        # uncompressed form carries the same flags on y
        assert ser.point_to_bytes(G, P, False)[-1] & 0x80 == f1
        with pytest.raises(ser.FormatError):
            bad = bytearray(ser.point_to_bytes(G, P, True)); bad[-1] |= 0xC0
            ser.point_from_bytes(G, bytes(bad), True)
        with pytest.raises(ser.FormatError):
            nb = G.F.base.nbytes if G.F.deg > 1 else G.F.nbytes
            bad = bytearray(ser.point_to_bytes(G, P, False)); bad[:nb] = (G.F.p).to_bytes(nb, "little")
            ser.point_from_bytes(G, bytes(bad), False)
