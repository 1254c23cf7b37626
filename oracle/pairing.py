"""Reduced Tate pairing and the `same_ratio` check, restated generically.

Test infrastructure (see oracle/__init__.py).  `setup_utils::same_ratio((a, b), (c, d))` is
e(a, d) == e(b, c) with a, b in G1 and c, d in G2.  arkworks uses
optimal ate pairings per curve family; any non-degenerate bilinear pairing on G1 x G2 gives
the same verdict, so the oracle uses the textbook reduced Tate pairing

    t(P, Q) = f_{r,P}(psi(Q)) ^ ((q^k - 1) / r)

over the full extension written as a binomial ring Fq[w] / (w^k - nu), with psi the untwisting
map (x', y') -> (x' w^(2s), y' w^(3s)), s = +1 for the D-type twist of BLS12-377 and s = -1 for
the other three curves.  Vertical lines are dropped (k even, x-coordinates of psi(Q) lie in
Fq^(k/2)): they die in the final exponentiation.
"""
from __future__ import annotations

from dataclasses import dataclass

from .curves import Curve, get_curve


class PolyExt:
    """Fq[w] / (w^k - nu); elements are length-k tuples of ints (coefficient of w^i at index i)."""

    def __init__(self, p: int, k: int, nu: int):
        self.p, self.k, self.nu = p, k, nu % p
        self.one = (1,) + (0,) * (k - 1)
        self.zero = (0,) * k

    def mul(self, a, b):
        p, k, nu = self.p, self.k, self.nu
        acc = [0] * (2 * k - 1)
        for i, ai in enumerate(a):
            if ai:
                for j, bj in enumerate(b):
                    if bj:
                        acc[i + j] += ai * bj
        out = [(acc[i] + (nu * acc[i + k] if i + k < 2 * k - 1 else 0)) % p for i in range(k)]
        return tuple(out)

    def sub(self, a, b):
        return tuple((x - y) % self.p for x, y in zip(a, b))

    def scale(self, a, s: int):
        return tuple(x * s % self.p for x in a)

    def pow(self, a, e: int):
        r = self.one
        for bit in bin(e)[2:]:
            r = self.mul(r, r)
            if bit == "1":
                r = self.mul(r, a)
        return r

    def w_pow(self, n: int):
        """w^n for any integer n (w^-1 = w^(k-1) / nu)."""
        k = self.k
        q, r = divmod(n, k)                       # w^n = nu^q * w^r (q may be negative)
        coeff = pow(self.nu, q, self.p)
        out = [0] * k
        out[r] = coeff
        return tuple(out)


@dataclass
class PairingParams:
    k: int          # embedding degree
    nu: int         # w^k = nu
    twist_sign: int  # +1: psi multiplies by w^2, w^3 ; -1: by w^-2, w^-3


PAIRING = {
    "bls12_377": PairingParams(12, -5, +1),     # Fq2 = Fq[u]/(u^2+5), w^6 = u
    "bw6_761": PairingParams(6, -4, -1),        # G2 over Fq, M-type twist b' = b * (-4)
    "mnt4_753": PairingParams(4, 13, -1),       # Fq2 = Fq[u]/(u^2-13), w^2 = u, quadratic twist by u
    "mnt6_753": PairingParams(6, 11, -1),       # Fq3 = Fq[u]/(u^3-11), w^2 = u
}


def embed(curve: Curve, K: PolyExt, x):
    """twist-field element (Fq, Fq2 or Fq3; u = w^(k/deg)) -> Fq^k"""
    F = curve.g2.F
    cs = F.coeffs(x) if F.deg > 1 else (x,)
    step = K.k // F.deg if F.deg > 1 else 0
    out = [0] * K.k
    for i, c in enumerate(cs):
        out[i * step] = c % K.p
    return tuple(out)


def untwist(curve: Curve, K: PolyExt, Q):
    pp = PAIRING[curve.name]
    x = K.mul(embed(curve, K, Q[0]), K.w_pow(2 * pp.twist_sign))
    y = K.mul(embed(curve, K, Q[1]), K.w_pow(3 * pp.twist_sign))
    return x, y


def miller(curve: Curve, K: PolyExt, P, Qk):
    """f_{r,P}(Q) with vertical lines dropped; P affine in E(Fq), Qk = psi(Q) in E(Fq^k)."""
    G = curve.g1
    p = curve.Fq.p
    xq, yq = Qk
    f = K.one
    T = P
    xP, yP = P

    def line(T, lam):
        # (yq - yT) - lam (xq - xT), as an Fq^k element
        xT, yT = T
        c = list(K.sub(yq, K.scale(xq, lam)))
        c[0] = (c[0] - yT + lam * xT) % p
        return tuple(c)

    for bit in bin(G.r)[3:]:
        xT, yT = T
        lam = (3 * xT * xT + G.a) * pow(2 * yT, -1, p) % p
        f = K.mul(K.mul(f, f), line(T, lam))
        T = G.add(T, T)
        if bit == "1":
            if T is None or (T[0] - xP) % p == 0:
                # T = -P (only at the very end): the line is vertical, dropped
                T = G.add(T, P)
                continue
            xT, yT = T
            lam = (yT - yP) * pow(xT - xP, -1, p) % p
            f = K.mul(f, line(T, lam))
            T = G.add(T, P)
    assert T is None
    return f


def tate(curve, P, Q):
    """Reduced Tate pairing of P in G1 (affine, not infinity) and Q in G2 (on the twist)."""
    c = get_curve(curve) if isinstance(curve, str) else curve
    pp = PAIRING[c.name]
    K = PolyExt(c.Fq.p, pp.k, pp.nu)
    if P is None or Q is None:
        return K.one
    f = miller(c, K, P, untwist(c, K, Q))
    return K.pow(f, (c.Fq.p ** pp.k - 1) // c.Fr.p)


def same_ratio(curve, g1_pair, g2_pair) -> bool:
    """setup_utils::same_ratio: e(a, d) == e(b, c)."""
    a, b = g1_pair
    c_, d = g2_pair
    return tate(curve, a, d) == tate(curve, b, c_)
