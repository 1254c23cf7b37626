"""ctypes front end of oracle/c/oracle.cpp — the multi-threaded C++ restatement of the reference
algorithm (per-index `pow`, per-point double-and-add, batch normalisation, arkworks byte
formats).  Test infrastructure and timed CPU baseline only (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess

from .curves import get_curve
from .params import HASH_SIZE, Phase1Params
from .phase1 import PrivateKey, calculate_hash

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "c"), "-s"])


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "c", "oracle.cpp")
        if not os.path.exists(_SO) or os.path.getmtime(src) > os.path.getmtime(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def hw_threads() -> int:
    return lib().orc_hw_threads()


def _sb(curve, k):
    return None if k is None else int(k).to_bytes(curve.Fr.nbytes, "little")


class OracleStatus(Exception):
    NAMES = {1: "non-canonical field element", 2: "invalid flags", 3: "point not on curve", 4: "point at infinity",
             5: "point not in the prime-order subgroup"}

    def __init__(self, code, index):
        super().__init__("%s (element %d)" % (self.NAMES.get(code, "?"), index))
        self.code, self.index = code, index


def batch_exp(curve, group: int, data: bytes, n: int, first_index: int, tau: int, coeff, mode: int = 0, in_compressed=False,
              out_compressed=True, check: int = 0, threads: int = 0, raise_on_status=True) -> bytes:
    c = get_curve(curve) if isinstance(curve, str) else curve
    G = (c.g1, c.g2)[group]
    osz = G.F.nbytes * (1 if out_compressed else 2)
    out = ctypes.create_string_buffer(max(1, n * osz))
    st = (ctypes.c_uint64 * 2)()
    rc = lib().orc_batch_exp(c.cid, group, data, int(in_compressed), ctypes.c_uint64(n), ctypes.c_uint64(first_index), _sb(c, tau),
                             _sb(c, coeff), mode, out, int(out_compressed), check, threads or hw_threads(), st)
    assert rc == 0
    if st[0] and raise_on_status:
        raise OracleStatus(st[0], st[1])
    return out.raw[:n * osz]


def reencode(curve, group: int, data: bytes, n: int, in_compressed=True, out_compressed=False, check: int = 2, subgroup=True,
             threads: int = 0) -> bytes:
    c = get_curve(curve) if isinstance(curve, str) else curve
    G = (c.g1, c.g2)[group]
    osz = G.F.nbytes * (1 if out_compressed else 2)
    out = ctypes.create_string_buffer(max(1, n * osz))
    st = (ctypes.c_uint64 * 2)()
    rc = lib().orc_reencode(c.cid, group, data, int(in_compressed), ctypes.c_uint64(n), out, int(out_compressed), check,
                            int(subgroup), threads or hw_threads(), st)
    assert rc == 0
    if st[0]:
        raise OracleStatus(st[0], st[1])
    return out.raw[:n * osz]


def contribute_with_key(params: Phase1Params, challenge: bytes, key: PrivateKey, pubkey_bytes: bytes, threads: int = 0,
                        check: int = 0) -> bytes:
    """Same contract as oracle.phase1.contribute_with_key, computed by the C++ port."""
    assert len(challenge) == params.accumulator_size
    c = params.curve
    ou = params.offsets(False)
    s, g1n, on = params.start, params.g1_count, params.other_count
    parts = [calculate_hash(challenge)]
    parts.append(batch_exp(c, 0, challenge[ou[0]:ou[1]], g1n, s, key.tau, None, check=check, threads=threads))
    parts.append(batch_exp(c, 1, challenge[ou[1]:ou[2]], on, s, key.tau, None, check=check, threads=threads))
    parts.append(batch_exp(c, 0, challenge[ou[2]:ou[3]], on, s, key.tau, key.alpha, check=check, threads=threads))
    parts.append(batch_exp(c, 0, challenge[ou[3]:ou[4]], on, s, key.tau, key.beta, check=check, threads=threads))
    parts.append(batch_exp(c, 1, challenge[ou[4]:ou[5]], 1, 0, 1, key.beta, mode=1, check=check, threads=1))
    out = b"".join(parts) + pubkey_bytes
    assert len(out) == params.contribution_size
    return out


def decompress_response(params: Phase1Params, response: bytes, check: int = 2, subgroup: bool = True, threads: int = 0) -> bytes:
    oc = params.offsets(True)
    counts = (params.g1_count, params.other_count, params.other_count, params.other_count, 1)
    groups = (0, 1, 0, 0, 1)
    parts = [calculate_hash(response)]
    for i in range(5):
        parts.append(reencode(params.curve, groups[i], response[oc[i]:oc[i + 1]], counts[i], check=check, subgroup=subgroup,
                              threads=threads))
    return b"".join(parts)


def field_mul(field: int, a: bytes, b: bytes, n: int) -> bytes:
    out = ctypes.create_string_buffer(len(a))
    assert lib().orc_field_mul(field, a, b, out, ctypes.c_uint64(n)) == 0
    return out.raw


def new_challenge(params: Phase1Params) -> bytes:
    """oracle.phase1.new_challenge without the per-element Python loop (every element is the same generator)."""
    from . import serialize as ser
    c = params.curve
    g1 = ser.point_to_bytes(c.g1, c.g1.gen, False)
    g2 = ser.point_to_bytes(c.g2, c.g2.gen, False)
    out = calculate_hash(b"") + g1 * params.g1_count + g2 * params.other_count + g1 * (2 * params.other_count) + g2
    assert len(out) == params.accumulator_size
    return out


def decompress_vectors(params: Phase1Params, response: bytes, threads: int = 0):
    """the five vectors of a response, uncompressed, no checks (the `decompress` hook of oracle.phase1.combine)"""
    oc = params.offsets(True)
    counts = (params.g1_count, params.other_count, params.other_count, params.other_count, 1)
    groups = (0, 1, 0, 0, 1)
    return [reencode(params.curve, groups[i], response[oc[i]:oc[i + 1]], counts[i], check=0, subgroup=False, threads=threads)
            for i in range(5)]


def combine(params: Phase1Params, responses, threads: int = 0) -> bytes:
    from . import phase1
    return phase1.combine(params, responses, decompress=lambda cp, r: decompress_vectors(cp, r, threads))
