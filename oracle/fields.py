"""Prime fields and their quadratic / cubic extensions on Python integers.

Test infrastructure (see oracle/__init__.py).  Restates ark-ff 0.4.2
(Cargo.lock:222-223 of the reference): `Fp` with canonical representatives,
`QuadExtField` c0 + c1*u with u^2 = nr, `CubicExtField` c0 + c1*u + c2*u^2 with
u^3 = nr.  Elements are plain ints (Fp) or tuples of ints (extensions); the
field object carries the operations so the curve code is generic.

Ordering (`gt`) follows ark-ff's `Ord`: extension elements compare the highest
coefficient first (c1 then c0; c2, c1, c0) — this decides the "y is negative"
serialisation flag.
"""
from __future__ import annotations


class Fp:
    """GF(p), canonical representatives 0..p-1."""

    deg = 1

    def __init__(self, p: int):
        self.p = p
        self.bits = p.bit_length()
        self.nbytes = (self.bits + 7) // 8          # ark-serialize: ceil(bits/8)
        self.limbs64 = (self.bits + 63) // 64        # ark-ff BigInt<N>
        self.R = 1 << (64 * self.limbs64)            # Montgomery radix used by ark-ff
        self.zero = 0
        self.one = 1
        # 2-adicity for Tonelli-Shanks
        s, t = 0, p - 1
        while t % 2 == 0:
            s, t = s + 1, t // 2
        self.two_adicity, self.t_odd = s, t
        self._qnr = None

    # -- arithmetic ---------------------------------------------------------
    def add(self, a, b): return (a + b) % self.p
    def sub(self, a, b): return (a - b) % self.p
    def neg(self, a): return (-a) % self.p
    def mul(self, a, b): return (a * b) % self.p
    def sqr(self, a): return (a * a) % self.p
    def muli(self, a, k: int): return (a * k) % self.p     # multiply by small integer
    def inv(self, a):
        if a % self.p == 0:
            raise ZeroDivisionError("inverse of zero")
        return pow(a, -1, self.p)
    def pow(self, a, e: int): return pow(a, e, self.p)
    def is_zero(self, a): return a % self.p == 0
    def eq(self, a, b): return (a - b) % self.p == 0
    def from_int(self, k: int): return k % self.p
    def coeffs(self, a): return (a,)
    def from_coeffs(self, cs): return cs[0] % self.p

    def gt(self, a, b) -> bool:
        return a > b

    def legendre(self, a) -> int:
        if a % self.p == 0:
            return 0
        return 1 if pow(a, (self.p - 1) // 2, self.p) == 1 else -1

    def _nonresidue(self):
        if self._qnr is None:
            g = 2
            while self.legendre(g) != -1:
                g += 1
            self._qnr = g
        return self._qnr

    def sqrt(self, a):
        """Some square root of a, or None.  Callers pick the sign themselves."""
        a %= self.p
        if a == 0:
            return 0
        if self.legendre(a) != 1:
            return None
        p = self.p
        if p % 4 == 3:
            return pow(a, (p + 1) // 4, p)
        # Tonelli-Shanks
        s, t = self.two_adicity, self.t_odd
        z = pow(self._nonresidue(), t, p)
        m, c = s, z
        x = pow(a, (t + 1) // 2, p)
        b = pow(a, t, p)
        while b != 1:
            i, b2 = 0, b
            while b2 != 1:
                b2 = b2 * b2 % p
                i += 1
            w = pow(c, 1 << (m - i - 1), p)
            x = x * w % p
            c = w * w % p
            b = b * c % p
            m = i
        return x


class ExtField:
    """Fp[u]/(u^deg - nr), deg in {2, 3}; elements are tuples of ints."""

    def __init__(self, base: Fp, deg: int, nr: int):
        assert deg in (2, 3)
        self.base, self.deg, self.nr = base, deg, nr % base.p
        self.p = base.p
        self.nbytes = base.nbytes * deg
        self.zero = (0,) * deg
        self.one = (1,) + (0,) * (deg - 1)
        self.order = base.p ** deg
        s, t = 0, self.order - 1
        while t % 2 == 0:
            s, t = s + 1, t // 2
        self.two_adicity, self.t_odd = s, t
        self._qnr = None

    def add(self, a, b): p = self.p; return tuple((x + y) % p for x, y in zip(a, b))
    def sub(self, a, b): p = self.p; return tuple((x - y) % p for x, y in zip(a, b))
    def neg(self, a): p = self.p; return tuple((-x) % p for x in a)
    def muli(self, a, k: int): p = self.p; return tuple((x * k) % p for x in a)
    def is_zero(self, a): return all(x % self.p == 0 for x in a)
    def eq(self, a, b): return all((x - y) % self.p == 0 for x, y in zip(a, b))
    def from_int(self, k: int): return (k % self.p,) + (0,) * (self.deg - 1)
    def coeffs(self, a): return tuple(a)
    def from_coeffs(self, cs): return tuple(c % self.p for c in cs)

    def mul(self, a, b):
        p, nr = self.p, self.nr
        if self.deg == 2:
            a0, a1 = a; b0, b1 = b
            return ((a0 * b0 + nr * a1 * b1) % p, (a0 * b1 + a1 * b0) % p)
        a0, a1, a2 = a; b0, b1, b2 = b
        return ((a0 * b0 + nr * (a1 * b2 + a2 * b1)) % p,
                (a0 * b1 + a1 * b0 + nr * a2 * b2) % p,
                (a0 * b2 + a1 * b1 + a2 * b0) % p)

    def sqr(self, a): return self.mul(a, a)

    def inv(self, a):
        p, nr = self.p, self.nr
        if self.is_zero(a):
            raise ZeroDivisionError("inverse of zero")
        if self.deg == 2:
            a0, a1 = a
            n = pow((a0 * a0 - nr * a1 * a1) % p, -1, p)
            return (a0 * n % p, (-a1 * n) % p)
        a0, a1, a2 = a
        # adjugate of the multiplication matrix
        t0 = (a0 * a0 - nr * a1 * a2) % p
        t1 = (nr * a2 * a2 - a0 * a1) % p
        t2 = (a1 * a1 - a0 * a2) % p
        n = pow((a0 * t0 + nr * (a2 * t1 + a1 * t2)) % p, -1, p)
        return (t0 * n % p, t1 * n % p, t2 * n % p)

    def pow(self, a, e: int):
        r = self.one
        if e < 0:
            a, e = self.inv(a), -e
        while e:
            if e & 1:
                r = self.mul(r, a)
            a = self.mul(a, a)
            e >>= 1
        return r

    def gt(self, a, b) -> bool:
        # ark-ff Ord for Quad/CubicExtField: highest coefficient first
        for x, y in zip(reversed(a), reversed(b)):
            if x != y:
                return x > y
        return False

    def legendre(self, a) -> int:
        if self.is_zero(a):
            return 0
        return 1 if self.pow(a, (self.order - 1) // 2) == self.one else -1

    def _nonresidue(self):
        if self._qnr is None:
            # deterministic search over small elements c0 + u
            k = 0
            while True:
                cand = (k,) + (1,) + (0,) * (self.deg - 2)
                if self.legendre(cand) == -1:
                    self._qnr = cand
                    break
                k += 1
        return self._qnr

    def sqrt(self, a):
        if self.is_zero(a):
            return self.zero
        if self.legendre(a) != 1:
            return None
        s, t = self.two_adicity, self.t_odd
        z = self.pow(self._nonresidue(), t)
        m, c = s, z
        x = self.pow(a, (t + 1) // 2)
        b = self.pow(a, t)
        one = self.one
        while b != one:
            i, b2 = 0, b
            while b2 != one:
                b2 = self.mul(b2, b2)
                i += 1
            w = c
            for _ in range(m - i - 1):
                w = self.mul(w, w)
            x = self.mul(x, w)
            c = self.mul(w, w)
            b = self.mul(b, c)
            m = i
        return x
