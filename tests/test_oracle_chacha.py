"""Known-answer tests for the ChaCha20 stream behind `derive_rng_from_seed`
(rand_chacha 0.3.1; reference src/bin/contribute.rs:789) and the arkworks samplers."""
from oracle import chacha
from oracle.curves import get_curve


def test_rfc7539_block():
    key = [0x03020100, 0x07060504, 0x0b0a0908, 0x0f0e0d0c, 0x13121110, 0x17161514, 0x1b1a1918, 0x1f1e1d1c]
    # RFC 7539 §2.3.2: counter 1, nonce 00:00:00:09 00:00:00:4a 00:00:00:00 mapped onto (ctr_hi, stream)
    blk = chacha.chacha20_block(key, 1 | (0x09000000 << 32), 0x4a000000)
    assert blk[:4] == [0xe4e7f110, 0x15593bd1, 0x1fdd0f50, 0xc47120a3]
    assert blk[12:] == [0xd19c12b5, 0xb94e16de, 0xe883d0cb, 0x4e3c50a2]


def test_zero_key_stream_and_word_order():
    r = chacha.ChaChaRng(bytes(32))
    assert r.fill_bytes(16).hex() == "76b8e0ada0f13d90405d6ae55386bd28"
    r = chacha.ChaChaRng(bytes(32))
    assert r.next_u64() == 0x903df1a0ade0b876          # low word first
    r = chacha.ChaChaRng(bytes(32))
    first = [r.next_u32() for _ in range(17)]
    assert first[16] == chacha.chacha20_block([0] * 8, 1)[0]  # 64-bit block counter increments


def test_fr_rand_is_montgomery_interpretation():
    c = get_curve("bls12_377")
    rng = chacha.ChaChaRng(bytes(range(32)))
    raw = chacha.ChaChaRng(bytes(range(32)))
    limbs = [raw.next_u64() for _ in range(4)]
    limbs[3] &= (1 << 61) - 1                            # 256 - 253 = 3 bits shaved
    v = sum(l << (64 * i) for i, l in enumerate(limbs))
    got = chacha.fp_rand(c.Fr, rng)
    if v < c.Fr.p:                                       # no rejection on this seed
        assert got * (1 << 256) % c.Fr.p == v
    assert 0 <= got < c.Fr.p


def test_group_rand_lands_in_subgroup():
    c = get_curve("bls12_377")
    rng = chacha.ChaChaRng(b"\x07" * 32)
    for G in (c.g1, c.g2):
        P = chacha.group_rand(G, rng)
        assert P is not None and G.on_curve(P) and G.mul(P, G.r) is None
