"""Canonical byte formats of field elements and curve points.

Test infrastructure (see oracle/__init__.py).  Restates ark-serialize 0.4.2 /
ark-ec 0.4.2 `CanonicalSerialize` for `Fp`, `QuadExtField`, `CubicExtField` and
short-Weierstrass `Affine`:

* Fp: canonical integer, little-endian, ceil(bits/8) bytes (pinned for Fr by the
  e2e/circuit_* fixtures of the reference, tests/test_oracle_formats.py).
* extensions: coefficients in order c0 || c1 (|| c2).
* `SWFlags` live in the two top bits of the LAST byte of the field element they are
  attached to: bit 7 = y is "negative" (y > -y), bit 6 = point at infinity.
* compressed point  : x with flags.
* uncompressed point: x (no flags) || y with flags.
"""
from __future__ import annotations

from .curves import Group

FLAG_NEG = 0x80
FLAG_INF = 0x40


class FormatError(ValueError):
    pass


def field_to_bytes(F, a, flags: int = 0) -> bytes:
    nb = (F.base.nbytes if F.deg > 1 else F.nbytes)
    out = bytearray()
    for c in F.coeffs(a):
        out += int(c).to_bytes(nb, "little")
    out[-1] |= flags
    return bytes(out)


def field_from_bytes(F, buf: bytes, with_flags: bool = False):
    """-> (element, flags).  Raises FormatError on a non-canonical coefficient."""
    base = F.base if F.deg > 1 else F
    nb = base.nbytes
    if len(buf) != nb * F.deg:
        raise FormatError("bad length")
    buf = bytearray(buf)
    flags = 0
    if with_flags:
        flags = buf[-1] & (FLAG_NEG | FLAG_INF)
        buf[-1] &= 0x3F
    cs = []
    for i in range(F.deg):
        c = int.from_bytes(buf[i * nb:(i + 1) * nb], "little")
        if c >= base.p:
            raise FormatError("non-canonical field element")
        cs.append(c)
    return F.from_coeffs(cs), flags


def point_size(G: Group, compressed: bool) -> int:
    return G.F.nbytes * (1 if compressed else 2)


def _y_flag(G: Group, y) -> int:
    F = G.F
    return FLAG_NEG if F.gt(y, F.neg(y)) else 0


def point_to_bytes(G: Group, P, compressed: bool) -> bytes:
    F = G.F
    if P is None:
        if compressed:
            return field_to_bytes(F, F.zero, FLAG_INF)
        return field_to_bytes(F, F.zero) + field_to_bytes(F, F.zero, FLAG_INF)
    x, y = P
    fl = _y_flag(G, y)
    if compressed:
        return field_to_bytes(F, x, fl)
    return field_to_bytes(F, x) + field_to_bytes(F, y, fl)


def point_from_bytes(G: Group, buf: bytes, compressed: bool):
    """No curve / subgroup validation here (ark `Validate::No`); see phase1.check_point."""
    F = G.F
    n = F.nbytes
    if compressed:
        x, fl = field_from_bytes(F, buf, True)
        if fl == (FLAG_NEG | FLAG_INF):
            raise FormatError("invalid flags")
        if fl & FLAG_INF:
            return None
        P = G.point_from_x(x, bool(fl & FLAG_NEG))
        if P is None:
            raise FormatError("x is not on the curve")
        return P
    x, _ = field_from_bytes(F, buf[:n], False)
    y, fl = field_from_bytes(F, buf[n:], True)
    if fl == (FLAG_NEG | FLAG_INF):
        raise FormatError("invalid flags")
    if fl & FLAG_INF:
        return None
    return (x, y)


def points_to_bytes(G: Group, pts, compressed: bool) -> bytes:
    return b"".join(point_to_bytes(G, P, compressed) for P in pts)


def points_from_bytes(G: Group, buf: bytes, compressed: bool):
    sz = point_size(G, compressed)
    assert len(buf) % sz == 0
    return [point_from_bytes(G, buf[i:i + sz], compressed) for i in range(0, len(buf), sz)]
