"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol that
include/sso_b200.h declares, does its host-only arithmetic, and refuses to compute without a
GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, has_gpu

import snark_setup_operator_b200 as sso
from oracle.params import Phase1Params
from oracle.phase1 import calculate_hash

LIB = sso.library_path()
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="libsso_b200.so not built (run __graft_entry__.build())")


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "sso_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sso_[a-z0-9_]+)\s*\(", hdr)))


def test_exports_every_declared_symbol():
    L = ctypes.CDLL(LIB)
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(L, n), "missing export " + n


@pytest.mark.parametrize("curve,power,clog,k", [("bw6_761", 12, 10, 0), ("bw6_761", 12, 10, 7), ("bls12_377", 20, 16, 0),
                                               ("mnt4_753", 20, 16, 31), ("mnt6_753", 20, 16, 0), ("bw6_761", 26, 20, 127)])
def test_sizes_agree_with_oracle(curve, power, clog, k):
    got = sso.Phase1Parameters.new_chunk(curve, k, 1 << clog, power, 1 << clog).sizes()
    o = Phase1Params.new_chunk(curve, k, 1 << clog, power, 1 << clog)
    assert got["g1_count"] == o.g1_count and got["other_count"] == o.other_count
    assert got["accumulator_size"] == o.accumulator_size and got["contribution_size"] == o.contribution_size
    assert got["public_key_size"] == o.public_key_size and got["num_chunks"] == o.num_chunks


def test_full_mode_sizes():
    got = sso.Phase1Parameters.new_full("bw6_761", 26, 1 << 20).sizes()
    assert got["accumulator_size"] == 64424509504 and got["contribution_size"] == 32212256512 + 0


def test_bad_parameters_are_errors():
    with pytest.raises(sso.SsoError) as e:
        sso.Phase1Parameters("bls12_377", 10, 0, 0, 0, 0, 0).sizes()        # chunked with chunk_size 0
    assert e.value.code == -1
    with pytest.raises(sso.SsoError):
        sso.Phase1Parameters("bls12_377", 10, 0, 4, 4, 0, 1).sizes()        # Marlin


def test_blake2b_matches_hashlib():
    L = sso.lib()
    for n in (0, 1, 127, 128, 129, 1000, 65537):
        data = bytes((i * 7 + 3) & 0xFF for i in range(n))
        out = ctypes.create_string_buffer(64)
        assert L.sso_blake2b_512(data, n, out) == 0
        assert out.raw == calculate_hash(data)


@pytest.mark.skipif(has_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    assert sso.lib().sso_device_count() == 0
    p = sso.Phase1Parameters.new_chunk("bls12_377", 0, 4, 3, 4)
    resp = bytearray(p.contribution_size)
    with pytest.raises(sso.SsoError) as e:
        sso.contribute_buf(p, bytes(p.accumulator_size), resp, 1, 2, 3)
    assert e.value.code == -2 and "no CPU fallback" in e.value.message
