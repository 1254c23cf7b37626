// Quadratic and cubic extension fields over Fp for the G2 groups.
//
// B200-native counterpart of ark-ff 0.4.2 `QuadExtField` / `CubicExtField` as used by the G2
// groups of BLS12-377 (Fq2, u^2 = -5), MNT4-753 (Fq2, u^2 = 13) and MNT6-753 (Fq3, u^3 = 11)
// (SURVEY.md §8a table, Appendix A.1).  Karatsuba multiplication: Fq2 = 3 Fq-mul, Fq3 = 6 Fq-mul.
// Ordering for the serialisation sign flag compares the highest coefficient first (A.2).
#pragma once
#include "fp.cuh"

namespace sso {

// Non-residue multiplication by a small integer: x -> (NEG ? -K x : K x)
template <class B, int K, bool NEG> struct SmallNR {
  static constexpr int ABS = K;
  static constexpr bool IS_NEG = NEG;
  __device__ __forceinline__ static typename B::T mul(const typename B::T& x) {
    typename B::T t = B::template mul_small<K>(x);
    return NEG ? B::neg(t) : t;
  }
  // c + nr x as an UNREDUCED integer below (K + 1) p: an operand for a multiplication (fp.cuh "lazy operands")
  __device__ __forceinline__ static typename B::T mad_lazy(const typename B::T& c, const typename B::T& x) {
    return B::mad_small_lazy(c, (uint32_t)K, NEG ? B::neg_lazy(x) : x);
  }
  // the same for an x below 4 p (point formulas with unreduced sums, ec.cuh SSO_LAZY_EC): a negative non-residue subtracts x
  // from 4 p; the result stays below c + 4 K p
  __device__ __forceinline__ static typename B::T mad_lazy_wide(const typename B::T& c, const typename B::T& x) {
    if constexpr (!NEG) return B::mad_small_lazy(c, (uint32_t)K, x);
    else {
      typename B::T p4, n;
#pragma unroll
      for (int i = 0; i < B::L; i++) p4.v[i] = (B::P::p()[i] << 2) | (i ? B::P::p()[i - 1] >> 30 : 0u);
      limbs_sub<B::L>(n.v, p4.v, x.v);
      return B::mad_small_lazy(c, (uint32_t)K, n);
    }
  }
  // c - (nr + 1) x, canonical
  __device__ __forceinline__ static typename B::T sub_nr_plus_one(const typename B::T& c, const typename B::T& x) {
    constexpr int M = NEG ? K - 1 : K + 1;                 // |nr + 1|
    if constexpr (M <= 4) {
      typename B::T m = B::template mul_small<M>(x);
      return NEG ? B::add(c, m) : B::sub(c, m);
    } else {
      return B::reduce_small(B::mad_small_lazy(c, (uint32_t)M, NEG ? x : B::neg_lazy(x)));
    }
  }
};

template <class B_, class NR> struct Fp2 {
  using B = B_;
  using Base = B_;
  static constexpr int DEG = 2;
  static constexpr int COOP = 0;
  // The products below only hand their arguments (and unreduced sums of them) to base multiplications, so the point formulas may
  // pass coefficients below A p, B p.  mul needs 4 A B p < R.  sqr forms (a0 + a1)(a0 + nr a1): for u^2 = -5 (BLS12-377,
  // R / p = 152.3) the subtrahend is taken from 4 p, so A <= 3 (the largest ec.cuh produces there: 3 X^2) gives
  // 6 p * (3 + 5 * 4) p = 138 p^2; for u^2 = 13 (MNT4-753, R / p = 37054) A <= 29 (3 X^2 + 26 Z^4) gives 58 p * 406 p = 23548 p^2.
  static constexpr bool LAZY_OK = B_::SPARE_BITS >= 7;
  static constexpr bool SQR_CHEAPER = true;           // two base multiplications against three
  static constexpr int NBYTES = 2 * B::NBYTES;
  static constexpr int WORDS = 2 * B::L;
  struct T { typename B::T c0, c1; };

  __device__ __forceinline__ static T zero() { return T{B::zero(), B::zero()}; }
  __device__ __forceinline__ static T one() { return T{B::one(), B::zero()}; }
  __device__ __forceinline__ static bool is_zero(const T& a) { return B::is_zero(a.c0) && B::is_zero(a.c1); }
  __device__ __forceinline__ static bool eq(const T& a, const T& b) { return B::eq(a.c0, b.c0) && B::eq(a.c1, b.c1); }
  __device__ __forceinline__ static T add(const T& a, const T& b) { return T{B::add(a.c0, b.c0), B::add(a.c1, b.c1)}; }
  __device__ __forceinline__ static T sub(const T& a, const T& b) { return T{B::sub(a.c0, b.c0), B::sub(a.c1, b.c1)}; }
  __device__ __forceinline__ static T dbl(const T& a) { return T{B::dbl(a.c0), B::dbl(a.c1)}; }
  __device__ __forceinline__ static T neg(const T& a) { return T{B::neg(a.c0), B::neg(a.c1)}; }
  template <int K> __device__ __forceinline__ static T mul_small(const T& a) {
    return T{B::template mul_small<K>(a.c0), B::template mul_small<K>(a.c1)};
  }
  __device__ __forceinline__ static T add_lazy(const T& a, const T& b) { return T{B::add_lazy(a.c0, b.c0), B::add_lazy(a.c1, b.c1)}; }
  __device__ __forceinline__ static T sub_lazy(const T& a, const T& b) { return T{B::sub_lazy(a.c0, b.c0), B::sub_lazy(a.c1, b.c1)}; }
  __device__ __forceinline__ static T mad_small_lazy(const T& c, uint32_t k, const T& x) {
    return T{B::mad_small_lazy(c.c0, k, x.c0), B::mad_small_lazy(c.c1, k, x.c1)};
  }
  // out of line, parameters and result by value (registers under the device ABI; see fp.cuh)
#ifndef SSO_FQ2_INLINE_BASE_MUL
#define SSO_FQ2_INLINE_BASE_MUL 0
#endif
  __device__ __forceinline__ static typename B::T bmul(const typename B::T& x, const typename B::T& y) {
#if SSO_FQ2_INLINE_BASE_MUL
    return B::mul_inl(x, y);
#else
    return B::mul(x, y);
#endif
  }
  __device__ __noinline__ static T mul_val(T a, T b) {
    typename B::T v0 = bmul(a.c0, b.c0);
    typename B::T v1 = bmul(a.c1, b.c1);
#ifndef SSO_NO_LAZY_OPERANDS
    typename B::T s = bmul(B::add_lazy(a.c0, a.c1), B::add_lazy(b.c0, b.c1));
#else
    typename B::T s = bmul(B::add(a.c0, a.c1), B::add(b.c0, b.c1));
#endif
    T r;
    r.c0 = B::add(v0, NR::mul(v1));
    r.c1 = B::sub(B::sub(s, v0), v1);
    return r;
  }
  __device__ __noinline__ static T sqr_val(T a) {
    // (a0 + a1)(a0 + nr a1) = a0^2 + nr a1^2 + (nr + 1) a0 a1; the two factors stay unreduced (< 2 p and < (|nr| + 1) p:
    // their product is below p R for every tower here), so forming them costs two carry chains instead of seven modular ops
    typename B::T v = bmul(a.c0, a.c1);
#ifndef SSO_NO_LAZY_OPERANDS
    typename B::T pr = bmul(B::add_lazy(a.c0, a.c1), NR::mad_lazy_wide(a.c0, a.c1));
    T r;
    r.c0 = NR::sub_nr_plus_one(pr, v);
#else
    typename B::T pr = bmul(B::add(a.c0, a.c1), B::add(a.c0, NR::mul(a.c1)));
    T r;
    r.c0 = B::sub(B::sub(pr, v), NR::mul(v));
#endif
    r.c1 = B::dbl(v);
    return r;
  }
  __device__ __forceinline__ static T mul(const T& a, const T& b) { return mul_val(a, b); }
  __device__ __forceinline__ static T sqr(const T& a) { return sqr_val(a); }
  __device__ __forceinline__ static T mul_base(const T& a, const typename B::T& k) { return T{B::mul(a.c0, k), B::mul(a.c1, k)}; }
  __device__ __forceinline__ static T conj(const T& a) { return T{a.c0, B::neg(a.c1)}; }           // Frobenius of Fq2
  __device__ __noinline__ static T inv(const T& a) {
    typename B::T n = B::sub(B::sqr(a.c0), NR::mul(B::sqr(a.c1)));
    typename B::T ni = B::inv(n);
    return T{B::mul(a.c0, ni), B::neg(B::mul(a.c1, ni))};
  }
  __device__ __forceinline__ static bool lex_is_neg(const T& a) {
    int s1 = B::sign_canonical(B::from_mont(a.c1));
    if (s1 != 0) return s1 > 0;
    return B::sign_canonical(B::from_mont(a.c0)) > 0;
  }
  // a^e, e little-endian words
  __device__ __noinline__ static T pow_words(const T& a, const uint32_t* e, int nwords) {
    T r = one();
    bool started = false;
    for (int i = nwords * 32 - 1; i >= 0; i--) {
      if (started) r = sqr(r);
      if ((e[i >> 5] >> (i & 31)) & 1) { r = started ? mul(r, a) : a; started = true; }
    }
    return r;
  }
  // complex-method square root; false if a is a non-square
  __device__ __noinline__ static bool sqrt(const T& a, T& out) {
    if (B::is_zero(a.c1)) {
      typename B::T r;
      if (B::sqrt(a.c0, r)) { out = T{r, B::zero()}; return true; }
      // a0 is a non-residue in Fp: sqrt = sqrt(a0 / nr) * u
      typename B::T nrv = NR::mul(B::one());
      typename B::T q = B::mul(a.c0, B::inv(nrv));
      if (!B::sqrt(q, r)) return false;
      out = T{B::zero(), r};
      return true;
    }
    typename B::T norm = B::sub(B::sqr(a.c0), NR::mul(B::sqr(a.c1)));
    typename B::T s;
    if (!B::sqrt(norm, s)) return false;
    typename B::T two_inv = B::inv(B::dbl(B::one()));
    typename B::T delta = B::mul(B::add(a.c0, s), two_inv);
    typename B::T c0;
    if (!B::sqrt(delta, c0)) {
      delta = B::mul(B::sub(a.c0, s), two_inv);
      if (!B::sqrt(delta, c0)) return false;
    }
    typename B::T c1 = B::mul(a.c1, B::inv(B::dbl(c0)));
    out = T{c0, c1};
    return eq(sqr(out), a);
  }
  __device__ __forceinline__ static bool from_bytes(const uint8_t* src, bool with_flags, uint32_t& flags, T& out) {
    uint32_t f0;
    bool ok0 = B::from_bytes(src, false, f0, out.c0);
    bool ok1 = B::from_bytes(src + B::NBYTES, with_flags, flags, out.c1);
    return ok0 && ok1;
  }
  __device__ __forceinline__ static void to_bytes(uint8_t* dst, const T& a, uint32_t flags) {
    B::to_bytes(dst, a.c0, 0);
    B::to_bytes(dst + B::NBYTES, a.c1, flags);
  }
  __device__ __forceinline__ static T load(const uint32_t* p, size_t stride) {
    return T{B::load(p, stride), B::load(p + B::L * stride, stride)};
  }
  __device__ __forceinline__ static void store(uint32_t* p, size_t stride, const T& a) {
    B::store(p, stride, a.c0);
    B::store(p + B::L * stride, stride, a.c1);
  }
  __device__ __forceinline__ static T from_const(const uint32_t* c) { return T{B::from_const(c), B::from_const(c + B::L)}; }
};

template <class B_, class NR> struct Fp3 {
  using B = B_;
  using Base = B_;
  static constexpr int DEG = 3;
  static constexpr int COOP = 0;
  static constexpr bool LAZY_OK = B_::SPARE_BITS >= 15;  // coefficients below A p, B p: products up to 4 A B p^2 (ec.cuh keeps A, B <= 8)
  static constexpr bool SQR_CHEAPER = false;          // squaring = multiplication
  static constexpr int NBYTES = 3 * B::NBYTES;
  static constexpr int WORDS = 3 * B::L;
  struct T { typename B::T c0, c1, c2; };

  __device__ __forceinline__ static T zero() { return T{B::zero(), B::zero(), B::zero()}; }
  __device__ __forceinline__ static T one() { return T{B::one(), B::zero(), B::zero()}; }
  __device__ __forceinline__ static bool is_zero(const T& a) { return B::is_zero(a.c0) && B::is_zero(a.c1) && B::is_zero(a.c2); }
  __device__ __forceinline__ static bool eq(const T& a, const T& b) { return B::eq(a.c0, b.c0) && B::eq(a.c1, b.c1) && B::eq(a.c2, b.c2); }
  __device__ __forceinline__ static T add(const T& a, const T& b) { return T{B::add(a.c0, b.c0), B::add(a.c1, b.c1), B::add(a.c2, b.c2)}; }
  __device__ __forceinline__ static T sub(const T& a, const T& b) { return T{B::sub(a.c0, b.c0), B::sub(a.c1, b.c1), B::sub(a.c2, b.c2)}; }
  __device__ __forceinline__ static T dbl(const T& a) { return T{B::dbl(a.c0), B::dbl(a.c1), B::dbl(a.c2)}; }
  __device__ __forceinline__ static T neg(const T& a) { return T{B::neg(a.c0), B::neg(a.c1), B::neg(a.c2)}; }
  template <int K> __device__ __forceinline__ static T mul_small(const T& a) {
    return T{B::template mul_small<K>(a.c0), B::template mul_small<K>(a.c1), B::template mul_small<K>(a.c2)};
  }
  __device__ __forceinline__ static T add_lazy(const T& a, const T& b) { return T{B::add_lazy(a.c0, b.c0), B::add_lazy(a.c1, b.c1), B::add_lazy(a.c2, b.c2)}; }
  __device__ __forceinline__ static T sub_lazy(const T& a, const T& b) { return T{B::sub_lazy(a.c0, b.c0), B::sub_lazy(a.c1, b.c1), B::sub_lazy(a.c2, b.c2)}; }
  __device__ __noinline__ static T mul(const T& a, const T& b) {
    typename B::T v0 = B::mul(a.c0, b.c0);
    typename B::T v1 = B::mul(a.c1, b.c1);
    typename B::T v2 = B::mul(a.c2, b.c2);
#ifndef SSO_NO_LAZY_OPERANDS
    typename B::T t12 = B::mul(B::add_lazy(a.c1, a.c2), B::add_lazy(b.c1, b.c2));
    typename B::T t01 = B::mul(B::add_lazy(a.c0, a.c1), B::add_lazy(b.c0, b.c1));
    typename B::T t02 = B::mul(B::add_lazy(a.c0, a.c2), B::add_lazy(b.c0, b.c2));
#else
    typename B::T t12 = B::mul(B::add(a.c1, a.c2), B::add(b.c1, b.c2));
    typename B::T t01 = B::mul(B::add(a.c0, a.c1), B::add(b.c0, b.c1));
    typename B::T t02 = B::mul(B::add(a.c0, a.c2), B::add(b.c0, b.c2));
#endif
    T r;
    r.c0 = B::add(v0, NR::mul(B::sub(B::sub(t12, v1), v2)));
    r.c1 = B::add(B::sub(B::sub(t01, v0), v1), NR::mul(v2));
    r.c2 = B::add(B::sub(B::sub(t02, v0), v2), v1);
    return r;
  }
  __device__ __noinline__ static T sqr(const T& a) { return mul(a, a); }
  __device__ __forceinline__ static T mul_base(const T& a, const typename B::T& k) { return T{B::mul(a.c0, k), B::mul(a.c1, k), B::mul(a.c2, k)}; }
  // (c0, c1 w1, c2 w2): the q-power Frobenius with the base-field constants w1, w2 of the tower
  __device__ __forceinline__ static T frob_w(const T& a, const uint32_t* w1, const uint32_t* w2) {
    return T{a.c0, B::mul(a.c1, B::from_const(w1)), B::mul(a.c2, B::from_const(w2))};
  }
  // x u^2 = (nr c1, nr c2, c0)
  __device__ __forceinline__ static T mul_u2(const T& a) { return T{NR::mul(a.c1), NR::mul(a.c2), a.c0}; }
  __device__ __noinline__ static T inv(const T& a) {
    typename B::T t0 = B::sub(B::sqr(a.c0), NR::mul(B::mul(a.c1, a.c2)));
    typename B::T t1 = B::sub(NR::mul(B::sqr(a.c2)), B::mul(a.c0, a.c1));
    typename B::T t2 = B::sub(B::sqr(a.c1), B::mul(a.c0, a.c2));
    typename B::T n = B::add(B::mul(a.c0, t0), NR::mul(B::add(B::mul(a.c2, t1), B::mul(a.c1, t2))));
    typename B::T ni = B::inv(n);
    return T{B::mul(t0, ni), B::mul(t1, ni), B::mul(t2, ni)};
  }
  __device__ __forceinline__ static bool lex_is_neg(const T& a) {
    int s = B::sign_canonical(B::from_mont(a.c2));
    if (s != 0) return s > 0;
    s = B::sign_canonical(B::from_mont(a.c1));
    if (s != 0) return s > 0;
    return B::sign_canonical(B::from_mont(a.c0)) > 0;
  }
  __device__ __noinline__ static T pow_words(const T& a, const uint32_t* e, int nwords) {
    T r = one();
    bool started = false;
    for (int i = nwords * 32 - 1; i >= 0; i--) {
      if (started) r = sqr(r);
      if ((e[i >> 5] >> (i & 31)) & 1) { r = started ? mul(r, a) : a; started = true; }
    }
    return r;
  }
  // Tonelli-Shanks in the cubic extension (q^3 - 1 has the 2-adicity of q - 1); TS supplies
  // (t-1)/2 and qnr^t for this tower.  false if a is a non-square.
  // The same square root with w = a^((t-1)/2) formed through the norm map: q^3 - 1 = 2^s t with t = m N, m = (q - 1) / 2^s,
  // N = q^2 + q + 1, and (t - 1)/2 = ((m - 1)/2) N + (N - 1)/2 with (N - 1)/2 = q (q + 1)/2, so
  //   w = Norm(a)^((m-1)/2) * Frob(a^((q+1)/2)),   Norm(a) = a a^q a^(q^2) in Fq:
  // one 753-bit exponentiation in Fq3 and one in Fq instead of the 2230-bit one in Fq3 (a third of the multiplications of the
  // decompression of an MNT6-753 G2 point).  Same w, hence the same root, as sqrt_ts.  w1, w2: Frobenius constants of the tower.
  template <class TS>
  __device__ __noinline__ static bool sqrt_ts_norm(const T& a, T& out, const uint32_t* tsz, const uint32_t* w1, const uint32_t* w2) {
    if (is_zero(a)) { out = a; return true; }
    T fa = frob_w(a, w1, w2), ffa = frob_w(fa, w1, w2);
    T aa = mul(a, fa);
    // coefficient 0 of aa * ffa (the others vanish)
    typename B::T n = B::add(B::mul(aa.c0, ffa.c0), NR::mul(B::add(B::mul(aa.c1, ffa.c2), B::mul(aa.c2, ffa.c1))));
    typename B::T u = B::pow_const(n, B::P::tm1h());                  // Norm(a)^((m-1)/2)
    T v = mul(pow_words(a, B::P::half(), B::L), a);                   // a^((q-1)/2) a
    T w = mul_base(frob_w(v, w1, w2), u);
    return ts_finish<TS>(a, w, out, tsz);
  }
  template <class TS>
  __device__ __noinline__ static bool sqrt_ts(const T& a, T& out, const uint32_t* tm1h, const uint32_t* tsz) {
    if (is_zero(a)) { out = a; return true; }
    T w = pow_words(a, tm1h, TS::TM1H_WORDS);
    return ts_finish<TS>(a, w, out, tsz);
  }
  // Tonelli-Shanks from w = a^((t-1)/2)
  template <class TS>
  __device__ __forceinline__ static bool ts_finish(const T& a, const T& w, T& out, const uint32_t* tsz) {
    T x = mul(a, w);
    T b = mul(x, w);
    T z = from_const(tsz);
    int m = TS::TWO_ADICITY;
    T o = one();
    while (!eq(b, o)) {
      int k = 0;
      T b2 = b;
      while (!eq(b2, o)) { b2 = sqr(b2); k++; if (k >= m) return false; }
      T wz = z;
      for (int j = 0; j < m - k - 1; j++) wz = sqr(wz);
      z = sqr(wz);
      b = mul(b, z);
      x = mul(x, wz);
      m = k;
    }
    out = x;
    return true;
  }
  __device__ __forceinline__ static bool from_bytes(const uint8_t* src, bool with_flags, uint32_t& flags, T& out) {
    uint32_t f0;
    bool ok0 = B::from_bytes(src, false, f0, out.c0);
    bool ok1 = B::from_bytes(src + B::NBYTES, false, f0, out.c1);
    bool ok2 = B::from_bytes(src + 2 * B::NBYTES, with_flags, flags, out.c2);
    return ok0 && ok1 && ok2;
  }
  __device__ __forceinline__ static void to_bytes(uint8_t* dst, const T& a, uint32_t flags) {
    B::to_bytes(dst, a.c0, 0);
    B::to_bytes(dst + B::NBYTES, a.c1, 0);
    B::to_bytes(dst + 2 * B::NBYTES, a.c2, flags);
  }
  __device__ __forceinline__ static T load(const uint32_t* p, size_t stride) {
    return T{B::load(p, stride), B::load(p + B::L * stride, stride), B::load(p + 2 * B::L * stride, stride)};
  }
  __device__ __forceinline__ static void store(uint32_t* p, size_t stride, const T& a) {
    B::store(p, stride, a.c0);
    B::store(p + B::L * stride, stride, a.c1);
    B::store(p + 2 * B::L * stride, stride, a.c2);
  }
  __device__ __forceinline__ static T from_const(const uint32_t* c) {
    return T{B::from_const(c), B::from_const(c + B::L), B::from_const(c + 2 * B::L)};
  }
};

}  // namespace sso
