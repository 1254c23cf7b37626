"""Validates the C++ restatement (oracle/c/oracle.cpp — the timed CPU baseline) against the Python big-int
oracle and the committed golden chunks, so that both legs of the oracle agree before either is trusted."""
import os
import random

import pytest

from oracle import cport, phase1, serialize as ser, synth
from oracle.curves import CURVE_NAMES, get_curve
from oracle.params import Phase1Params


@pytest.mark.parametrize("name", CURVE_NAMES)
@pytest.mark.parametrize("k", [0, 3])
def test_golden_chunks(name, k, golden_dir):
    p = Phase1Params.new_chunk(name, k, 4, 3, 4)
    ch = open(os.path.join(golden_dir, "p1_%s_c%d.challenge.bin" % (name, k)), "rb").read()
    want = open(os.path.join(golden_dir, "p1_%s_c%d.response.bin" % (name, k)), "rb").read()
    key = synth.contributor_key(p.curve)
    got = cport.contribute_with_key(p, ch, key, bytes(p.public_key_size), threads=4)
    assert got == want
    assert cport.decompress_response(p, got, threads=4) == phase1.decompress_response(p, got)


@pytest.mark.parametrize("fid,name,which", [(0, "bls12_377", "Fr"), (1, "bls12_377", "Fq"), (2, "bw6_761", "Fq"),
                                             (3, "mnt4_753", "Fq"), (4, "mnt6_753", "Fq")])
def test_field_mul(fid, name, which):
    F = getattr(get_curve(name), which)
    rnd = random.Random(fid)
    a = [rnd.randrange(F.p) for _ in range(64)] + [0, 1, F.p - 1]
    b = [rnd.randrange(F.p) for _ in range(64)] + [F.p - 1, F.p - 1, F.p - 1]
    A = b"".join(ser.field_to_bytes(F, x) for x in a)
    B = b"".join(ser.field_to_bytes(F, x) for x in b)
    assert cport.field_mul(fid, A, B, len(a)) == b"".join(ser.field_to_bytes(F, x * y % F.p) for x, y in zip(a, b))


def test_status_reporting():
    c = get_curve("bls12_377")
    from oracle.curves import _some_point
    rogue = ser.point_to_bytes(c.g1, _some_point(c.g1, 11), True)
    with pytest.raises(cport.OracleStatus) as e:
        cport.reencode(c, 0, rogue, 1)
    assert e.value.code == 5
    assert cport.reencode(c, 0, rogue, 1, subgroup=False) == ser.point_to_bytes(c.g1, _some_point(c.g1, 11), False)
    with pytest.raises(cport.OracleStatus) as e:
        cport.batch_exp(c, 0, ser.point_to_bytes(c.g1, None, False), 1, 0, 5, None, check=1)
    assert e.value.code == 4
