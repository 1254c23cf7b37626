// CPU restatement of the reference's algorithm for the phase-1 / phase-2 hot path — TEST
// INFRASTRUCTURE and timed CPU baseline only (see oracle/__init__.py; PARITY UNPINNED: the
// reference holds no golden vectors for this path).
//
// What is restated (the arithmetic lives in un-vendored crates: nimiq/snark-setup rev bd530da over
// arkworks 0.4.2 — Cargo.lock:150-368, 2603-2694, 3477-3500 of the reference):
//   ark-ff  Fp<MontBackend<_, N>>      64-bit limbs, Montgomery CIOS, R = 2^(64 N)
//   ark-ff  QuadExtField/CubicExtField  schoolbook / Karatsuba towers
//   ark-ec  short_weierstrass::Projective  Jacobian add/double, `mul_bigint` = MSB-first double-and-add
//   setup_utils::generate_powers_of_tau    one independent `pow` per index
//   setup_utils::batch_exp / batch_mul     per-point scalar multiplication, rayon over points
//   BatchSerializer / BatchDeserializer    ark-serialize 0.4 canonical formats (SW flags)
// Call sites in the reference: src/bin/contribute.rs:809-840 (contribute), :966-1009 (verify).
//
// Validated against the Python big-int oracle on the golden vectors (tests/test_oracle_c.py).
// Threads over points stand in for rayon's par_iter.
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

typedef unsigned __int128 u128;

struct CurveParams {
  const char *q, *r;
  const char *g1_a, *g1_b, *g1_x, *g1_y;
  int g2_deg;
  const char* g2_nr;
  const char* g2_a[3];
  const char* g2_b[3];
  const char* g2_x[3];
  const char* g2_y[3];
  int f1_s; const char *f1_tm1h, *f1_t; const char* f1_qnr[3];
  int f2_s; const char *f2_tm1h, *f2_t; const char* f2_qnr[3];
};
#include "params.inc"

// ------------------------------------------------------------------------------------------
// big integers as little-endian 64-bit limb vectors (exponents, init-time only)
// ------------------------------------------------------------------------------------------
static std::vector<uint64_t> hex_to_limbs(const char* hex, size_t min_limbs = 0) {
  size_t len = strlen(hex);
  std::vector<uint64_t> out((len + 15) / 16, 0);
  for (size_t i = 0; i < len; i++) {
    char ch = hex[len - 1 - i];
    uint64_t d = ch <= '9' ? ch - '0' : (ch | 32) - 'a' + 10;
    out[i / 16] |= d << (4 * (i % 16));
  }
  if (out.size() < min_limbs) out.resize(min_limbs, 0);
  return out;
}

// ------------------------------------------------------------------------------------------
// prime field, Montgomery form.  TAG separates fields of equal width.
// ------------------------------------------------------------------------------------------
template <int N, int TAG> struct Fp {
  uint64_t v[N];
  static uint64_t P[N], R1[N], R2[N], HALF[N], PM2[N];
  static uint64_t INV;
  static int BITS, NBYTES;
  static const int DEG = 1;
  static const int LIMBS = N;

  static int cmp_raw(const uint64_t* a, const uint64_t* b) {
    for (int i = N - 1; i >= 0; i--) if (a[i] != b[i]) return a[i] > b[i] ? 1 : -1;
    return 0;
  }
  static uint64_t add_raw(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    u128 c = 0;
    for (int i = 0; i < N; i++) { c += (u128)a[i] + b[i]; r[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
  }
  static uint64_t sub_raw(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    uint64_t borrow = 0;
    for (int i = 0; i < N; i++) {
      u128 d = (u128)a[i] - b[i] - borrow;
      r[i] = (uint64_t)d;
      borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
  }
  static void init(const char* hex_p) {
    std::vector<uint64_t> p = hex_to_limbs(hex_p, N);
    for (int i = 0; i < N; i++) P[i] = p[i];
    BITS = 0;
    for (int i = N * 64 - 1; i >= 0; i--) if ((P[i / 64] >> (i % 64)) & 1) { BITS = i + 1; break; }
    NBYTES = (BITS + 7) / 8;
    uint64_t x = 1;                               // Newton iteration for p^-1 mod 2^64
    for (int i = 0; i < 6; i++) x *= 2 - P[0] * x;
    INV = (uint64_t)0 - x;
    // R mod p by 64 N doublings of 1, R^2 by 64 N more
    uint64_t t[N] = {1};
    for (int i = 1; i < N; i++) t[i] = 0;
    for (int k = 0; k < 2 * 64 * N; k++) {
      uint64_t carry = add_raw(t, t, t);
      uint64_t u[N];
      if (carry || cmp_raw(t, P) >= 0) { sub_raw(u, t, P); memcpy(t, u, sizeof u); }
      if (k == 64 * N - 1) memcpy(R1, t, sizeof t);
    }
    memcpy(R2, t, sizeof t);
    uint64_t one[N] = {1}, two[N] = {2};
    for (int i = 1; i < N; i++) one[i] = two[i] = 0;
    sub_raw(HALF, P, one);
    for (int i = 0; i < N; i++) HALF[i] = (HALF[i] >> 1) | (i + 1 < N ? HALF[i + 1] << 63 : 0);
    sub_raw(PM2, P, two);
  }

  static Fp zero() { Fp r; memset(r.v, 0, sizeof r.v); return r; }
  static Fp one() { Fp r; memcpy(r.v, R1, sizeof r.v); return r; }
  bool is_zero() const { for (int i = 0; i < N; i++) if (v[i]) return false; return true; }
  bool operator==(const Fp& o) const { return memcmp(v, o.v, sizeof v) == 0; }
  bool operator!=(const Fp& o) const { return !(*this == o); }

  Fp operator+(const Fp& o) const {
    Fp r, t;
    uint64_t c = add_raw(r.v, v, o.v);
    if (c || cmp_raw(r.v, P) >= 0) { sub_raw(t.v, r.v, P); return t; }
    return r;
  }
  Fp operator-(const Fp& o) const {
    Fp r, t;
    if (sub_raw(r.v, v, o.v)) { add_raw(t.v, r.v, P); return t; }
    return r;
  }
  Fp neg() const { if (is_zero()) return *this; Fp r; sub_raw(r.v, P, v); return r; }
  Fp dbl() const { return *this + *this; }
  // CIOS Montgomery multiplication (ark-ff `mul_assign` without the asm / no-carry shortcuts)
  Fp operator*(const Fp& o) const {
    uint64_t t[N + 2];
    memset(t, 0, sizeof t);
    for (int i = 0; i < N; i++) {
      u128 c = 0;
      for (int j = 0; j < N; j++) { c += (u128)v[j] * o.v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
      c += t[N]; t[N] = (uint64_t)c; t[N + 1] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * INV;
      c = ((u128)m * P[0] + t[0]) >> 64;
      for (int j = 1; j < N; j++) { c += (u128)m * P[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
      c += t[N]; t[N - 1] = (uint64_t)c;
      t[N] = t[N + 1] + (uint64_t)(c >> 64);
    }
    Fp r, u;
    memcpy(r.v, t, sizeof r.v);
    if (t[N] || cmp_raw(r.v, P) >= 0) { sub_raw(u.v, r.v, P); return u; }
    return r;
  }
  Fp sqr() const { return *this * *this; }
  Fp mul_small(unsigned k) const { Fp acc = zero(), b = *this; while (k) { if (k & 1) acc = acc + b; b = b.dbl(); k >>= 1; } return acc; }
  static Fp from_canonical(const uint64_t* limbs) { Fp a, r2; memcpy(a.v, limbs, sizeof a.v); memcpy(r2.v, R2, sizeof r2.v); return a * r2; }
  void to_canonical(uint64_t* out) const { Fp o = zero(); o.v[0] = 1; Fp c = *this * o; memcpy(out, c.v, sizeof c.v); }
  static Fp from_hex(const char* hex) { std::vector<uint64_t> l = hex_to_limbs(hex, N); return from_canonical(l.data()); }
  static Fp from_u64(uint64_t k) { uint64_t l[N] = {k}; for (int i = 1; i < N; i++) l[i] = 0; return from_canonical(l); }

  Fp pow(const uint64_t* e, int nlimbs) const {     // MSB-first square-and-multiply (ark-ff `pow`)
    Fp r = one();
    bool started = false;
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
      if (started) r = r.sqr();
      if ((e[i / 64] >> (i % 64)) & 1) { r = started ? r * *this : *this; started = true; }
    }
    return r;
  }
  Fp inv() const { return pow(PM2, N); }
  // y > -y in canonical order
  bool lex_neg() const { uint64_t c[N]; to_canonical(c); return cmp_raw(c, HALF) > 0; }
  int lex_sign() const { if (is_zero()) return 0; return lex_neg() ? 1 : -1; }

  // ark-serialize: canonical little-endian, NBYTES bytes, flags in the top bits of the last byte
  void to_bytes(uint8_t* dst, uint8_t flags) const {
    uint64_t c[N];
    to_canonical(c);
    for (int i = 0; i < NBYTES; i++) dst[i] = (uint8_t)(c[i / 8] >> (8 * (i % 8)));
    dst[NBYTES - 1] |= flags;
  }
  static bool from_bytes(const uint8_t* src, bool with_flags, uint8_t& flags, Fp& out) {
    uint64_t c[N];
    memset(c, 0, sizeof c);
    for (int i = 0; i < NBYTES; i++) {
      uint8_t b = src[i];
      if (i == NBYTES - 1) { flags = with_flags ? (b & 0xC0) : 0; if (with_flags) b &= 0x3F; }
      c[i / 8] |= (uint64_t)b << (8 * (i % 8));
    }
    if (cmp_raw(c, P) >= 0) return false;
    out = from_canonical(c);
    return true;
  }
};
template <int N, int TAG> uint64_t Fp<N, TAG>::P[N];
template <int N, int TAG> uint64_t Fp<N, TAG>::R1[N];
template <int N, int TAG> uint64_t Fp<N, TAG>::R2[N];
template <int N, int TAG> uint64_t Fp<N, TAG>::HALF[N];
template <int N, int TAG> uint64_t Fp<N, TAG>::PM2[N];
template <int N, int TAG> uint64_t Fp<N, TAG>::INV;
template <int N, int TAG> int Fp<N, TAG>::BITS;
template <int N, int TAG> int Fp<N, TAG>::NBYTES;

// ------------------------------------------------------------------------------------------
// extensions: B[u]/(u^2 - NR), B[u]/(u^3 - NR); NR set at init (TAG separates instances)
// ------------------------------------------------------------------------------------------
template <class B, int TAG> struct Fp2 {
  B c0, c1;
  static B NR;
  static const int DEG = 2;
  static int nbytes() { return 2 * B::NBYTES; }
  static Fp2 zero() { return {B::zero(), B::zero()}; }
  static Fp2 one() { return {B::one(), B::zero()}; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
  bool operator!=(const Fp2& o) const { return !(*this == o); }
  Fp2 operator+(const Fp2& o) const { return {c0 + o.c0, c1 + o.c1}; }
  Fp2 operator-(const Fp2& o) const { return {c0 - o.c0, c1 - o.c1}; }
  Fp2 neg() const { return {c0.neg(), c1.neg()}; }
  Fp2 dbl() const { return {c0.dbl(), c1.dbl()}; }
  Fp2 operator*(const Fp2& o) const {              // Karatsuba, as ark-ff QuadExtField::mul_assign
    B v0 = c0 * o.c0, v1 = c1 * o.c1;
    return {v0 + NR * v1, (c0 + c1) * (o.c0 + o.c1) - v0 - v1};
  }
  Fp2 sqr() const { return *this * *this; }
  Fp2 mul_small(unsigned k) const { return {c0.mul_small(k), c1.mul_small(k)}; }
  Fp2 inv() const { B n = (c0.sqr() - NR * c1.sqr()).inv(); return {c0 * n, (c1 * n).neg()}; }
  bool lex_neg() const { int s = c1.lex_sign(); if (s) return s > 0; return c0.lex_sign() > 0; }
  Fp2 pow(const uint64_t* e, int nlimbs) const {
    Fp2 r = one(); bool started = false;
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
      if (started) r = r.sqr();
      if ((e[i / 64] >> (i % 64)) & 1) { r = started ? r * *this : *this; started = true; }
    }
    return r;
  }
  void to_bytes(uint8_t* dst, uint8_t flags) const { c0.to_bytes(dst, 0); c1.to_bytes(dst + B::NBYTES, flags); }
  static bool from_bytes(const uint8_t* src, bool with_flags, uint8_t& flags, Fp2& out) {
    uint8_t f0;
    bool a = B::from_bytes(src, false, f0, out.c0);
    bool b = B::from_bytes(src + B::NBYTES, with_flags, flags, out.c1);
    return a && b;
  }
  static Fp2 from_hex3(const char* const* h) { return {B::from_hex(h[0]), B::from_hex(h[1])}; }
};
template <class B, int TAG> B Fp2<B, TAG>::NR;

template <class B, int TAG> struct Fp3 {
  B c0, c1, c2;
  static B NR;
  static const int DEG = 3;
  static int nbytes() { return 3 * B::NBYTES; }
  static Fp3 zero() { return {B::zero(), B::zero(), B::zero()}; }
  static Fp3 one() { return {B::one(), B::zero(), B::zero()}; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero() && c2.is_zero(); }
  bool operator==(const Fp3& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
  bool operator!=(const Fp3& o) const { return !(*this == o); }
  Fp3 operator+(const Fp3& o) const { return {c0 + o.c0, c1 + o.c1, c2 + o.c2}; }
  Fp3 operator-(const Fp3& o) const { return {c0 - o.c0, c1 - o.c1, c2 - o.c2}; }
  Fp3 neg() const { return {c0.neg(), c1.neg(), c2.neg()}; }
  Fp3 dbl() const { return {c0.dbl(), c1.dbl(), c2.dbl()}; }
  Fp3 operator*(const Fp3& o) const {              // schoolbook: 9 base multiplications
    B r0 = c0 * o.c0 + NR * (c1 * o.c2 + c2 * o.c1);
    B r1 = c0 * o.c1 + c1 * o.c0 + NR * (c2 * o.c2);
    B r2 = c0 * o.c2 + c1 * o.c1 + c2 * o.c0;
    return {r0, r1, r2};
  }
  Fp3 sqr() const { return *this * *this; }
  Fp3 mul_small(unsigned k) const { return {c0.mul_small(k), c1.mul_small(k), c2.mul_small(k)}; }
  Fp3 inv() const {
    B t0 = c0.sqr() - NR * (c1 * c2), t1 = NR * c2.sqr() - c0 * c1, t2 = c1.sqr() - c0 * c2;
    B n = (c0 * t0 + NR * (c2 * t1 + c1 * t2)).inv();
    return {t0 * n, t1 * n, t2 * n};
  }
  bool lex_neg() const {
    int s = c2.lex_sign(); if (s) return s > 0;
    s = c1.lex_sign(); if (s) return s > 0;
    return c0.lex_sign() > 0;
  }
  Fp3 pow(const uint64_t* e, int nlimbs) const {
    Fp3 r = one(); bool started = false;
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
      if (started) r = r.sqr();
      if ((e[i / 64] >> (i % 64)) & 1) { r = started ? r * *this : *this; started = true; }
    }
    return r;
  }
  void to_bytes(uint8_t* dst, uint8_t flags) const { c0.to_bytes(dst, 0); c1.to_bytes(dst + B::NBYTES, 0); c2.to_bytes(dst + 2 * B::NBYTES, flags); }
  static bool from_bytes(const uint8_t* src, bool with_flags, uint8_t& flags, Fp3& out) {
    uint8_t f0;
    bool a = B::from_bytes(src, false, f0, out.c0);
    bool b = B::from_bytes(src + B::NBYTES, false, f0, out.c1);
    bool c = B::from_bytes(src + 2 * B::NBYTES, with_flags, flags, out.c2);
    return a && b && c;
  }
  static Fp3 from_hex3(const char* const* h) { return {B::from_hex(h[0]), B::from_hex(h[1]), B::from_hex(h[2])}; }
};
template <class B, int TAG> B Fp3<B, TAG>::NR;

template <class F> static int f_nbytes() { if constexpr (F::DEG == 1) return F::NBYTES; else return F::nbytes(); }

// ------------------------------------------------------------------------------------------
// short Weierstrass group, Jacobian coordinates (ark-ec formulas: dbl-2009-l when a = 0,
// dbl-2007-bl otherwise; add-2007-bl; madd-2007-bl)
// ------------------------------------------------------------------------------------------
template <class F, int TAG> struct Curve {
  static F A, B;
  static bool A_ZERO;
  static std::vector<uint64_t> ORDER;            // r
  static int TS_S;                                // Tonelli-Shanks data of the coordinate field
  static std::vector<uint64_t> TS_TM1H;
  static F TS_Z;

  struct Aff { F x, y; bool inf; };
  struct Jac { F X, Y, Z; };
  static Jac identity() { return {F::one(), F::one(), F::zero()}; }

  static F rhs(const F& x) { F r = x.sqr() * x + B; if (!A_ZERO) r = r + A * x; return r; }
  static bool on_curve(const Aff& p) { return p.inf || p.y.sqr() == rhs(p.x); }

  static Jac dbl(const Jac& p) {
    if (p.Z.is_zero()) return p;
    Jac r;
    if (A_ZERO) {
      F a = p.X.sqr(), b = p.Y.sqr(), c = b.sqr();
      F d = ((p.X + b).sqr() - a - c).dbl();
      F e = a.dbl() + a, f = e.sqr();
      r.Z = (p.Y * p.Z).dbl();
      r.X = f - d.dbl();
      r.Y = e * (d - r.X) - c.dbl().dbl().dbl();
    } else {
      F xx = p.X.sqr(), yy = p.Y.sqr(), yyyy = yy.sqr(), zz = p.Z.sqr();
      F s = ((p.X + yy).sqr() - xx - yyyy).dbl();
      F m = xx.dbl() + xx + A * zz.sqr();
      r.X = m.sqr() - s.dbl();
      r.Z = (p.Y + p.Z).sqr() - yy - zz;
      r.Y = m * (s - r.X) - yyyy.dbl().dbl().dbl();
    }
    return r;
  }
  static Jac madd(const Jac& p, const Aff& q) {
    if (q.inf) return p;
    if (p.Z.is_zero()) return {q.x, q.y, F::one()};
    F z1z1 = p.Z.sqr(), u2 = q.x * z1z1, s2 = q.y * p.Z * z1z1;
    if (p.X == u2) {
      if (p.Y == s2) return dbl(p);
      return identity();
    }
    F h = u2 - p.X, hh = h.sqr(), i = hh.dbl().dbl(), j = h * i, rr = (s2 - p.Y).dbl(), v = p.X * i;
    Jac r;
    r.X = rr.sqr() - j - v.dbl();
    r.Y = rr * (v - r.X) - (p.Y * j).dbl();
    r.Z = (p.Z + h).sqr() - z1z1 - hh;
    return r;
  }
  static Jac add(const Jac& p, const Jac& q) {
    if (p.Z.is_zero()) return q;
    if (q.Z.is_zero()) return p;
    F z1z1 = p.Z.sqr(), z2z2 = q.Z.sqr();
    F u1 = p.X * z2z2, u2 = q.X * z1z1, s1 = p.Y * q.Z * z2z2, s2 = q.Y * p.Z * z1z1;
    if (u1 == u2) {
      if (s1 == s2) return dbl(p);
      return identity();
    }
    F h = u2 - u1, i = h.dbl().sqr(), j = h * i, rr = (s2 - s1).dbl(), v = u1 * i;
    Jac r;
    r.X = rr.sqr() - j - v.dbl();
    r.Y = rr * (v - r.X) - (s1 * j).dbl();
    r.Z = ((p.Z + q.Z).sqr() - z1z1 - z2z2) * h;
    return r;
  }
  // ark-ec `mul_bigint`: MSB-first double-and-add over the scalar's bits
  static Jac mul(const Aff& base, const uint64_t* k, int nlimbs) {
    Jac acc = identity();
    bool started = false;
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
      if (started) acc = dbl(acc);
      if ((k[i / 64] >> (i % 64)) & 1) { acc = madd(acc, base); started = true; }
    }
    return acc;
  }
  // ark-ec `batch_normalization` (Montgomery's trick) over a contiguous range
  static void normalize(const Jac* in, Aff* out, size_t n) {
    std::vector<F> pre(n);
    F acc = F::one();
    for (size_t i = 0; i < n; i++) { pre[i] = acc; if (!in[i].Z.is_zero()) acc = acc * in[i].Z; }
    F inv = acc.inv();
    for (size_t k = n; k-- > 0;) {
      if (in[k].Z.is_zero()) { out[k] = {F::zero(), F::zero(), true}; continue; }
      F zi = inv * pre[k];
      inv = inv * in[k].Z;
      F zi2 = zi.sqr();
      out[k] = {in[k].X * zi2, in[k].Y * zi2 * zi, false};
    }
  }
  static bool sqrt(const F& a, F& out) {            // Tonelli-Shanks
    if (a.is_zero()) { out = a; return true; }
    F w = a.pow(TS_TM1H.data(), (int)TS_TM1H.size());
    F x = a * w, b = x * w, z = TS_Z;
    int m = TS_S;
    while (b != F::one()) {
      int k = 0;
      F b2 = b;
      while (b2 != F::one()) { b2 = b2.sqr(); if (++k >= m) return false; }
      F wz = z;
      for (int j = 0; j < m - k - 1; j++) wz = wz.sqr();
      z = wz.sqr(); b = b * z; x = x * wz; m = k;
    }
    out = x;
    return true;
  }

  static int size_c() { return f_nbytes<F>(); }
  static int size_u() { return 2 * f_nbytes<F>(); }
  // 0 ok, 1 non-canonical, 2 bad flags, 3 not on curve (compressed: x has no y)
  static int read(const uint8_t* src, bool compressed, Aff& out) {
    uint8_t fl = 0, f0 = 0;
    out.inf = false;
    if (compressed) {
      bool ok = F::from_bytes(src, true, fl, out.x);
      if (fl == 0xC0) return 2;
      if (!ok) return 1;
      if (fl & 0x40) { out = {F::zero(), F::zero(), true}; return 0; }
      F y;
      if (!sqrt(rhs(out.x), y)) return 3;
      if (y.lex_neg() != ((fl & 0x80) != 0)) y = y.neg();
      out.y = y;
      return 0;
    }
    bool okx = F::from_bytes(src, false, f0, out.x);
    bool oky = F::from_bytes(src + f_nbytes<F>(), true, fl, out.y);
    if (fl == 0xC0) return 2;
    if (!okx || !oky) return 1;
    if (fl & 0x40) out = {F::zero(), F::zero(), true};
    return 0;
  }
  static void write(uint8_t* dst, bool compressed, const Aff& p) {
    int nb = f_nbytes<F>();
    if (p.inf) {
      if (compressed) { F::zero().to_bytes(dst, 0x40); return; }
      F::zero().to_bytes(dst, 0); F::zero().to_bytes(dst + nb, 0x40); return;
    }
    uint8_t fl = p.y.lex_neg() ? 0x80 : 0;
    if (compressed) { p.x.to_bytes(dst, fl); return; }
    p.x.to_bytes(dst, 0); p.y.to_bytes(dst + nb, fl);
  }
};
template <class F, int TAG> F Curve<F, TAG>::A;
template <class F, int TAG> F Curve<F, TAG>::B;
template <class F, int TAG> bool Curve<F, TAG>::A_ZERO;
template <class F, int TAG> std::vector<uint64_t> Curve<F, TAG>::ORDER;
template <class F, int TAG> int Curve<F, TAG>::TS_S;
template <class F, int TAG> std::vector<uint64_t> Curve<F, TAG>::TS_TM1H;
template <class F, int TAG> F Curve<F, TAG>::TS_Z;

// ------------------------------------------------------------------------------------------
// the four curves
// ------------------------------------------------------------------------------------------
using Fr253 = Fp<4, 0>;
using Fq377 = Fp<6, 1>;
using Fq761 = Fp<12, 2>;
using Fq4 = Fp<12, 3>;
using Fq6 = Fp<12, 4>;
using Fq377x2 = Fp2<Fq377, 0>;
using Fq4x2 = Fp2<Fq4, 1>;
using Fq6x3 = Fp3<Fq6, 0>;

struct Bls { using Fr = Fr253; using G1 = Curve<Fq377, 0>; using G2 = Curve<Fq377x2, 1>; };
struct Bw6 { using Fr = Fq377; using G1 = Curve<Fq761, 2>; using G2 = Curve<Fq761, 3>; };
struct Mnt4 { using Fr = Fq6; using G1 = Curve<Fq4, 4>; using G2 = Curve<Fq4x2, 5>; };
struct Mnt6 { using Fr = Fq4; using G1 = Curve<Fq6, 6>; using G2 = Curve<Fq6x3, 7>; };

template <class F> static F elem_from_hex3(const char* const* h) {
  if constexpr (F::DEG == 1) return F::from_hex(h[0]); else return F::from_hex3(h);
}
template <class C, class F> static void init_group(const F& a, const F& b, const char* r, int s, const char* tm1h, const char* t,
                                                   const char* const* qnr) {
  C::A = a; C::B = b; C::A_ZERO = a.is_zero();
  C::ORDER = hex_to_limbs(r);
  C::TS_S = s;
  C::TS_TM1H = hex_to_limbs(tm1h);
  std::vector<uint64_t> tt = hex_to_limbs(t);
  C::TS_Z = elem_from_hex3<F>(qnr).pow(tt.data(), (int)tt.size());
}
template <class CV, class F1, class F2> static void init_curve(const CurveParams& P) {
  const char* a1[3] = {P.g1_a, "0", "0"};
  const char* b1[3] = {P.g1_b, "0", "0"};
  init_group<typename CV::G1, F1>(elem_from_hex3<F1>(a1), elem_from_hex3<F1>(b1), P.r, P.f1_s, P.f1_tm1h, P.f1_t, P.f1_qnr);
  init_group<typename CV::G2, F2>(elem_from_hex3<F2>(P.g2_a), elem_from_hex3<F2>(P.g2_b), P.r, P.f2_s, P.f2_tm1h, P.f2_t, P.f2_qnr);
}

static bool g_inited = false;
static void init_all() {
  if (g_inited) return;
  Fr253::init(PARAMS_BLS12_377.r);
  Fq377::init(PARAMS_BLS12_377.q);
  Fq761::init(PARAMS_BW6_761.q);
  Fq4::init(PARAMS_MNT4_753.q);
  Fq6::init(PARAMS_MNT6_753.q);
  Fq377x2::NR = Fq377::from_hex(PARAMS_BLS12_377.g2_nr);
  Fq4x2::NR = Fq4::from_hex(PARAMS_MNT4_753.g2_nr);
  Fq6x3::NR = Fq6::from_hex(PARAMS_MNT6_753.g2_nr);
  init_curve<Bls, Fq377, Fq377x2>(PARAMS_BLS12_377);
  init_curve<Bw6, Fq761, Fq761>(PARAMS_BW6_761);
  init_curve<Mnt4, Fq4, Fq4x2>(PARAMS_MNT4_753);
  init_curve<Mnt6, Fq6, Fq6x3>(PARAMS_MNT6_753);
  g_inited = true;
}

// ------------------------------------------------------------------------------------------
// batch operations (threads over points = rayon par_iter)
// ------------------------------------------------------------------------------------------
template <class Fn> static void parallel_ranges(size_t n, int threads, Fn fn) {
  if (threads < 1) threads = 1;
  if ((size_t)threads > n) threads = n ? (int)n : 1;
  std::vector<std::thread> th;
  size_t per = (n + threads - 1) / threads;
  for (int t = 0; t < threads; t++) {
    size_t lo = t * per, hi = lo + per < n ? lo + per : n;
    if (lo >= hi) break;
    th.emplace_back([=] { fn(lo, hi); });
  }
  for (auto& t : th) t.join();
}

// status: first failing (code, index) — 1 non-canonical, 2 bad flags, 3 not on curve, 4 infinity, 5 not in subgroup
struct Status { int code = 0; uint64_t index = 0; };

// out[j] = (coeff * tau^(first + j)) * in[j]   (mode 0)   |   out[j] = coeff * in[j]   (mode 1)
template <class C, class Fr>
static Status batch_exp(const uint8_t* in, bool in_c, size_t n, uint64_t first, const uint8_t* tau_b, const uint8_t* coeff_b, int mode,
                        uint8_t* out, bool out_c, int check, int threads) {
  std::vector<Status> sts(threads > 0 ? threads : 1);
  uint8_t fl;
  Fr tau = Fr::one(), coeff = Fr::one();
  if (tau_b) Fr::from_bytes(tau_b, false, fl, tau);
  if (coeff_b) Fr::from_bytes(coeff_b, false, fl, coeff);
  int isz = in_c ? C::size_c() : C::size_u(), osz = out_c ? C::size_c() : C::size_u();
  size_t per = (n + (threads > 0 ? threads : 1) - 1) / (threads > 0 ? threads : 1);
  parallel_ranges(n, threads, [&](size_t lo, size_t hi) {
    Status& st = sts[per ? lo / per : 0];
    std::vector<typename C::Jac> jac(hi - lo);
    for (size_t j = lo; j < hi; j++) {
      typename C::Aff p;
      int rc = C::read(in + j * isz, in_c, p);
      if (rc && !st.code) st = {rc, j};
      if (rc) p.inf = true;
      if (!rc && check >= 1 && p.inf && !st.code) st = {4, j};
      if (!rc && check >= 2 && !in_c && !C::on_curve(p)) { if (!st.code) st = {3, j}; p.inf = true; }
      Fr s;
      if (mode == 0) {
        uint64_t e = first + j;
        s = tau.pow(&e, 1);                          // generate_powers_of_tau: independent pow per index
        if (coeff_b) s = s * coeff;
      } else {
        s = coeff;
      }
      uint64_t k[Fr::LIMBS];
      s.to_canonical(k);
      jac[j - lo] = p.inf ? C::identity() : C::mul(p, k, Fr::LIMBS);
    }
    std::vector<typename C::Aff> aff(hi - lo);
    C::normalize(jac.data(), aff.data(), hi - lo);
    for (size_t j = lo; j < hi; j++) C::write(out + j * osz, out_c, aff[j - lo]);
  });
  for (auto& s : sts) if (s.code) return s;
  return Status();
}

template <class C>
static Status reencode(const uint8_t* in, bool in_c, size_t n, uint8_t* out, bool out_c, int check, bool subgroup, int threads) {
  std::vector<Status> sts(threads > 0 ? threads : 1);
  int isz = in_c ? C::size_c() : C::size_u(), osz = out_c ? C::size_c() : C::size_u();
  size_t per = (n + (threads > 0 ? threads : 1) - 1) / (threads > 0 ? threads : 1);
  parallel_ranges(n, threads, [&](size_t lo, size_t hi) {
    Status& st = sts[per ? lo / per : 0];
    for (size_t j = lo; j < hi; j++) {
      typename C::Aff p;
      int rc = C::read(in + j * isz, in_c, p);
      if (rc) { if (!st.code) st = {rc, j}; p = {p.x, p.y, true}; }
      else if (check >= 1) {
        if (p.inf) { if (!st.code) st = {4, j}; }
        else if (check >= 2) {
          if (!in_c && !C::on_curve(p)) { if (!st.code) st = {3, j}; }
          else if (subgroup) {
            typename C::Jac q = C::mul(p, C::ORDER.data(), (int)C::ORDER.size());
            if (!q.Z.is_zero() && !st.code) st = {5, j};
          }
        }
      }
      if (out) C::write(out + j * osz, out_c, p);
    }
  });
  for (auto& s : sts) if (s.code) return s;
  return Status();
}

template <class Fn> static int dispatch(uint32_t curve, uint32_t group, Fn fn) {
  switch (curve * 2 + group) {
    case 0: fn((Bls::G1*)0, (Bls::Fr*)0); return 0;
    case 1: fn((Bls::G2*)0, (Bls::Fr*)0); return 0;
    case 2: fn((Bw6::G1*)0, (Bw6::Fr*)0); return 0;
    case 3: fn((Bw6::G2*)0, (Bw6::Fr*)0); return 0;
    case 4: fn((Mnt4::G1*)0, (Mnt4::Fr*)0); return 0;
    case 5: fn((Mnt4::G2*)0, (Mnt4::Fr*)0); return 0;
    case 6: fn((Mnt6::G1*)0, (Mnt6::Fr*)0); return 0;
    case 7: fn((Mnt6::G2*)0, (Mnt6::Fr*)0); return 0;
  }
  return -1;
}

extern "C" {

// returns 0 ok; status[0] = first failing code, status[1] = index
int orc_batch_exp(uint32_t curve, uint32_t group, const uint8_t* in, uint32_t in_compressed, uint64_t n, uint64_t first_index,
                  const uint8_t* tau, const uint8_t* coeff, uint32_t mode, uint8_t* out, uint32_t out_compressed, uint32_t check,
                  int threads, uint64_t* status) {
  init_all();
  Status st;
  int rc = dispatch(curve, group, [&](auto* c, auto* fr) {
    using C = typename std::remove_pointer<decltype(c)>::type;
    using Fr = typename std::remove_pointer<decltype(fr)>::type;
    st = batch_exp<C, Fr>(in, in_compressed != 0, n, first_index, tau, coeff, (int)mode, out, out_compressed != 0, (int)check, threads);
  });
  status[0] = st.code; status[1] = st.index;
  return rc;
}

int orc_reencode(uint32_t curve, uint32_t group, const uint8_t* in, uint32_t in_compressed, uint64_t n, uint8_t* out,
                 uint32_t out_compressed, uint32_t check, uint32_t subgroup, int threads, uint64_t* status) {
  init_all();
  Status st;
  int rc = dispatch(curve, group, [&](auto* c, auto*) {
    using C = typename std::remove_pointer<decltype(c)>::type;
    st = reencode<C>(in, in_compressed != 0, n, out, out_compressed != 0, (int)check, subgroup != 0, threads);
  });
  status[0] = st.code; status[1] = st.index;
  return rc;
}

// element-wise field multiplication on canonical byte strings (field ids as in include/sso_b200.h test hook)
int orc_field_mul(uint32_t field, const uint8_t* a, const uint8_t* b, uint8_t* out, uint64_t n) {
  init_all();
  auto run = [&](auto* f) {
    using F = typename std::remove_pointer<decltype(f)>::type;
    uint8_t fl;
    for (uint64_t i = 0; i < n; i++) {
      F x, y;
      F::from_bytes(a + i * F::NBYTES, false, fl, x);
      F::from_bytes(b + i * F::NBYTES, false, fl, y);
      (x * y).to_bytes(out + i * F::NBYTES, 0);
    }
  };
  switch (field) {
    case 0: run((Fr253*)0); return 0;
    case 1: run((Fq377*)0); return 0;
    case 2: run((Fq761*)0); return 0;
    case 3: run((Fq4*)0); return 0;
    case 4: run((Fq6*)0); return 0;
  }
  return -1;
}

int orc_hw_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
