"""Pins the oracle's curve constants: primality, curve membership and
order of the generators, twist parameters, field encodings sizes."""
import pytest
import sympy

from oracle.curves import CURVE_NAMES, get_curve


@pytest.mark.parametrize("name", CURVE_NAMES)
def test_primes_and_generators(name):
    c = get_curve(name)
    assert sympy.isprime(c.Fq.p) and sympy.isprime(c.Fr.p)
    for G in (c.g1, c.g2):
        assert G.r == c.Fr.p
        assert G.on_curve(G.gen)
        assert G.mul(G.gen, G.r) is None
        assert G.mul(G.gen, G.r - 1) == G.neg(G.gen)


def test_sizes_table():
    # (Fq bits, Fr bits, G1 compressed, G2 compressed) — SURVEY.md §8a
    want = {"bls12_377": (377, 253, 48, 96), "bw6_761": (761, 377, 96, 96), "mnt4_753": (753, 753, 95, 190),
            "mnt6_753": (753, 753, 95, 285)}
    for name, (qb, rb, g1c, g2c) in want.items():
        c = get_curve(name)
        assert (c.Fq.bits, c.Fr.bits, c.g1.F.nbytes, c.g2.F.nbytes) == (qb, rb, g1c, g2c)


def test_cycle_and_chain_relations():
    m4, m6, bls, bw6 = (get_curve(n) for n in ("mnt4_753", "mnt6_753", "bls12_377", "bw6_761"))
    assert m4.Fq.p == m6.Fr.p and m4.Fr.p == m6.Fq.p          # MNT4/6 cycle
    assert bw6.Fr.p == bls.Fq.p                               # BW6-761 is built over BLS12-377's base field
    x = 0x8508c00000000001
    assert bls.Fr.p == x ** 4 - x ** 2 + 1
    assert bls.Fq.p == (x - 1) ** 2 * bls.Fr.p // 3 + x


@pytest.mark.parametrize("name", ["bls12_377", "mnt4_753", "mnt6_753"])
def test_cofactor_clears(name):
    c = get_curve(name)
    for G in (c.g1, c.g2):
        # a point built without reference to the generator lands in the r-torsion after clearing
        from oracle.curves import _some_point
        P = G.mul(_some_point(G, 7), G.cofactor)
        assert G.mul(P, G.r) is None


def test_twist_nonresidues():
    bls, m4, m6 = (get_curve(n) for n in ("bls12_377", "mnt4_753", "mnt6_753"))
    assert bls.Fq.legendre(bls.Fq.from_int(-5)) == -1
    assert m4.Fq.legendre(13) == -1
    q = m6.Fq.p
    assert pow(11, (q - 1) // 3, q) != 1                      # 11 is a cubic non-residue
