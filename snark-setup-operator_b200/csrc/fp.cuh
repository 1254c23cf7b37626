// Prime-field arithmetic for sm_100a: Montgomery form, 32-bit limbs, one thread per element.
//
// B200-native replacement for the role ark-ff 0.4.2 `Fp<MontBackend<_, N>>` plays under
// setup_utils::batch_exp.  The Montgomery radix is 2^(32 L), which
// equals ark-ff's 2^(64 ceil(bits/64)) for all five primes, so Montgomery residues coincide
// with the reference's in-memory representation (relevant for Fr::rand).
//
// Multiplication is CIOS with the partial products split by limb parity into two carry chains
// ("even" and "odd" accumulators) so that every 32x32->64 product lands in an adjacent register
// pair and ptxas can emit one IMAD.WIDE-class instruction per product with the carry riding the
// chain.  Work per multiplication: 2 L^2 + L multiply-accumulates.
//
// Attribution: the even/odd split of the CIOS rows (helpers mul_n / cmad_n / madc_n_rshift / mad_n_redc below) is the
// well-known public formulation used by Supranational's sppark (ff/mont_t.cuh, Apache-2.0) and several GPU bigint
// libraries after it; it is restated here from the published technique, not taken from /root/reference (which holds
// no arithmetic).  The dedicated squaring, the run-time Montgomery factor and the out-of-line by-value policy are ours.
#pragma once
#include <cstdint>
#include "constants.cuh"

// Field multiplications wider than this many limbs are kept out of line: ONE copy per field, parameters and
// result passed by value (= in registers under the device ABI).  Measured on B200: with the 12-limb
// multiplication inlined the point formulas are ~0.6 MB of SASS and the kernels stall on instruction fetch;
// out of line the hot loop fits the instruction cache: k_batch_exp G1 -20 %, G2 -30 %.
#ifndef SSO_INLINE_MUL_MAX_L
#define SSO_INLINE_MUL_MAX_L 8
#endif

// The point formulas use the dedicated squaring up to this many limbs (see Fp::sqr)
#ifndef SSO_SQR_EC_MAX_L
#define SSO_SQR_EC_MAX_L 12
#endif
// 1: the point formulas trade squarings for multiplications on every prime field (A/B knob; default: only where the field has
// no dedicated squaring)
#ifndef SSO_LAZY_TRADE_ALL
#define SSO_LAZY_TRADE_ALL 0
#endif

namespace sso {

// ---------------------------------------------------------------------------------------------
// PTX carry-chain primitives (CC.CF lives across consecutive asm volatile statements)
// ---------------------------------------------------------------------------------------------
#ifndef SSO_HOST_EMUL
__device__ __forceinline__ uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint64_t mul_wide(uint32_t a, uint32_t b) { uint64_t r; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t shl1_from(uint32_t lo, uint32_t hi) { return __funnelshift_l(lo, hi, 1); }   // (hi:lo << 1) >> 32
#else
// TEST-ONLY instruction-level emulation of the PTX carry flag so that the device algorithms can be
// exercised by g++ in this GPU-less container (tests/emul/).  Never part of the product library.
static thread_local uint32_t g_cf = 0;
static inline uint32_t emu_add(uint32_t a, uint32_t b, uint32_t cin, bool setc) { uint64_t s = (uint64_t)a + b + cin; if (setc) g_cf = (uint32_t)(s >> 32); return (uint32_t)s; }
static inline uint32_t emu_sub(uint32_t a, uint32_t b, uint32_t bin, bool setc) { uint64_t s = (uint64_t)a - b - bin; if (setc) g_cf = (uint32_t)((s >> 32) & 1); return (uint32_t)s; }
static inline uint32_t add_cc(uint32_t a, uint32_t b) { return emu_add(a, b, 0, true); }
static inline uint32_t addc_cc(uint32_t a, uint32_t b) { return emu_add(a, b, g_cf, true); }
static inline uint32_t addc(uint32_t a, uint32_t b) { return emu_add(a, b, g_cf, false); }
static inline uint32_t sub_cc(uint32_t a, uint32_t b) { return emu_sub(a, b, 0, true); }
static inline uint32_t subc_cc(uint32_t a, uint32_t b) { return emu_sub(a, b, g_cf, true); }
static inline uint32_t subc(uint32_t a, uint32_t b) { return emu_sub(a, b, g_cf, false); }
static inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
static inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint64_t mul_wide(uint32_t a, uint32_t b) { return (uint64_t)a * b; }
static inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_lo(a, b), c, 0, true); }
static inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_lo(a, b), c, g_cf, true); }
static inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_hi(a, b), c, g_cf, true); }
static inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_hi(a, b), c, g_cf, false); }
static inline uint32_t shl1_from(uint32_t lo, uint32_t hi) { return (hi << 1) | (lo >> 31); }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
#endif

// ---------------------------------------------------------------------------------------------
// fixed-width helpers on raw limb arrays
// ---------------------------------------------------------------------------------------------
template <int N> struct big { uint32_t v[N]; };

template <int N> __device__ __forceinline__ bool limbs_is_zero(const uint32_t* a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < N; i++) o |= a[i];
  return o == 0;
}
template <int N> __device__ __forceinline__ bool limbs_eq(const uint32_t* a, const uint32_t* b) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < N; i++) o |= a[i] ^ b[i];
  return o == 0;
}
// a > b (unsigned, little-endian limbs)
template <int N> __device__ __forceinline__ bool limbs_gt(const uint32_t* a, const uint32_t* b) {
  // b - a borrows  <=>  a > b
  sub_cc(b[0], a[0]);
#pragma unroll
  for (int i = 1; i < N; i++) subc_cc(b[i], a[i]);
  return subc(0, 0) != 0;
}
// r = a + b, returns carry
template <int N> __device__ __forceinline__ uint32_t limbs_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  r[0] = add_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < N; i++) r[i] = addc_cc(a[i], b[i]);
  return addc(0, 0);
}
// r = a - b, returns borrow (0 or 0xffffffff)
template <int N> __device__ __forceinline__ uint32_t limbs_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  r[0] = sub_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < N; i++) r[i] = subc_cc(a[i], b[i]);
  return subc(0, 0);
}

// ---------------------------------------------------------------------------------------------
// Montgomery multiplication building blocks (n = L limbs, n even)
// ---------------------------------------------------------------------------------------------
// acc[0..n) = sum_{j even} a[j] * bi * 2^(32 j)   (disjoint register pairs, no carries)
template <int n> __device__ __forceinline__ void mul_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
#pragma unroll
  for (int j = 0; j < n; j += 2) {
    uint64_t t = mul_wide(a[j], bi);               // one IMAD.WIDE instead of an IMAD + IMAD.HI pair
    acc[j] = (uint32_t)t;
    acc[j + 1] = (uint32_t)(t >> 32);
  }
}
// acc += sum_{j even} a[j] * bi * 2^(32 j); carry-out left in CC.CF
template <int n> __device__ __forceinline__ void cmad_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
  acc[0] = mad_lo_cc(a[0], bi, acc[0]);
  acc[1] = madc_hi_cc(a[0], bi, acc[1]);
#pragma unroll
  for (int j = 2; j < n; j += 2) {
    acc[j] = madc_lo_cc(a[j], bi, acc[j]);
    acc[j + 1] = madc_hi_cc(a[j], bi, acc[j + 1]);
  }
}
// acc = (acc >> 64) + sum_{j even} a[j] * bi * 2^(32 j), carry-in from CC.CF
template <int n> __device__ __forceinline__ void madc_n_rshift(uint32_t* acc, const uint32_t* a, uint32_t bi) {
#pragma unroll
  for (int j = 0; j < n - 2; j += 2) {
    acc[j] = madc_lo_cc(a[j], bi, acc[j + 2]);
    acc[j + 1] = madc_hi_cc(a[j], bi, acc[j + 3]);
  }
  acc[n - 2] = madc_lo_cc(a[n - 2], bi, 0);
  acc[n - 1] = madc_hi(a[n - 2], bi, 0);
}

// Note on instruction selection (measured, tools/mulbench + SASS): the Montgomery factor m = t * (-p^-1 mod 2^32)
// must be computed with the constant read at RUN TIME (P::inv_rt(), constant bank).  For p = 1 (mod 2^32) — the
// 377-bit and 253-bit primes — the compile-time constant is 0xffffffff; ptxas then rewrites m as a negation and
// emits every m * p product of the reduction as an IMAD.X + IMAD.HI.U32.X pair instead of one IMAD.WIDE.U32.X
// (+120 integer-pipe instructions per 12-limb multiplication).
template <class P> __device__ __forceinline__ void load_modulus(uint32_t* mod) {
#pragma unroll
  for (int i = 0; i < P::L; i++) mod[i] = P::p()[i];
}

// One CIOS step: T <- (T + a * bi + m * p) / 2^32 with T held as E (aligned at limb 0) plus
// O (aligned at limb 1).  After the step the roles of E and O are exchanged (caller swaps).
template <class P> __device__ __forceinline__ void mad_n_redc(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi, bool first,
                                                              const uint32_t* MOD) {
  constexpr int n = P::L;
  if (first) {
    mul_n<n>(O, a + 1, bi);
    mul_n<n>(E, a, bi);
  } else {
    E[0] = add_cc(E[0], O[1]);
    madc_n_rshift<n>(O, a + 1, bi);
    cmad_n<n>(E, a, bi);
    O[n - 1] = addc(O[n - 1], 0);
  }
  uint32_t mi = E[0] * P::inv_rt();
  cmad_n<n>(O, MOD + 1, mi);
  cmad_n<n>(E, MOD, mi);
  O[n - 1] = addc(O[n - 1], 0);
}

// r = a * b * R^-1 mod p, inputs and output fully reduced
template <class P> __device__ __forceinline__ void mont_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int n = P::L;
  uint32_t even[n], odd[n], mod[n];
  load_modulus<P>(mod);
#pragma unroll
  for (int i = 0; i < n; i += 2) {
    mad_n_redc<P>(even, odd, a, b[i], i == 0, mod);
    mad_n_redc<P>(odd, even, a, b[i + 1], false, mod);
  }
  // merge: result[k] = even[k] + odd[k+1]
  even[0] = add_cc(even[0], odd[1]);
#pragma unroll
  for (int k = 1; k < n - 1; k++) even[k] = addc_cc(even[k], odd[k + 1]);
  even[n - 1] = addc(even[n - 1], 0);
  // final subtraction: result < 2p
  uint32_t t[n];
  uint32_t borrow = limbs_sub<n>(t, even, mod);
#pragma unroll
  for (int k = 0; k < n; k++) r[k] = borrow ? even[k] : t[k];
}

// ---------------------------------------------------------------------------------------------
// Dedicated squaring: the n(n-1)/2 off-diagonal products once, doubled by a funnel shift, plus the n squares on
// the diagonal, then a Montgomery reduction of the 2n-limb result: n(n+1)/2 + n^2 + n multiply-accumulates
// instead of 2n^2 + n (234 / 300 for 12 limbs, 900 / 1176 for 24).
// ---------------------------------------------------------------------------------------------
// one row of off-diagonal products a[i] * a[j], j = j0, j0+2, ... accumulated at limb positions i+j (lo), i+j+1 (hi)
// of acc; the pairs of one row are contiguous, so the carry rides the chain and is deposited in the limb above the
// row (which no product of this or an earlier row has reached: it holds at most earlier deposits)
template <int n, int i, int j0> __device__ __forceinline__ void sqr_row(uint32_t* acc, const uint32_t* a) {
  if constexpr (j0 < n) {
    acc[i + j0] = mad_lo_cc(a[i], a[j0], acc[i + j0]);
    acc[i + j0 + 1] = madc_hi_cc(a[i], a[j0], acc[i + j0 + 1]);
    constexpr int last = j0 + 2 * ((n - 1 - j0) / 2);
#pragma unroll
    for (int j = j0 + 2; j < n; j += 2) {
      acc[i + j] = madc_lo_cc(a[i], a[j], acc[i + j]);
      acc[i + j + 1] = madc_hi_cc(a[i], a[j], acc[i + j + 1]);
    }
    acc[i + last + 2] = addc(acc[i + last + 2], 0);
  }
}
template <int n, int i> __device__ __forceinline__ void sqr_rows(uint32_t* ev, uint32_t* od, const uint32_t* a) {
  if constexpr (i < n - 1) {
    sqr_row<n, i, i + 1>(od, a);                    // i + j odd: register pairs (odd, even) of od
    sqr_row<n, i, i + 2>(ev, a);                    // i + j even: register pairs (even, odd) of ev
    sqr_rows<n, i + 1>(ev, od, a);
  }
}
// w[0..2n) = a^2
template <int n> __device__ __forceinline__ void sqr_wide(uint32_t* w, const uint32_t* a) {
  uint32_t ev[2 * n], od[2 * n];
#pragma unroll
  for (int k = 0; k < 2 * n; k++) { ev[k] = 0; od[k] = 0; }
  sqr_rows<n, 0>(ev, od, a);
  // s = ev + od (limb 0 is empty: the lowest off-diagonal product sits at limb 1)
  ev[1] = add_cc(ev[1], od[1]);
#pragma unroll
  for (int k = 2; k < 2 * n - 1; k++) ev[k] = addc_cc(ev[k], od[k]);
  ev[2 * n - 1] = addc(ev[2 * n - 1], od[2 * n - 1]);
  // w = 2 s + sum a[i]^2 2^(64 i)
  w[0] = mad_lo_cc(a[0], a[0], 0);
  w[1] = madc_hi_cc(a[0], a[0], ev[1] << 1);
#pragma unroll
  for (int i = 1; i < n; i++) {
    w[2 * i] = madc_lo_cc(a[i], a[i], shl1_from(ev[2 * i - 1], ev[2 * i]));
    w[2 * i + 1] = i == n - 1 ? madc_hi(a[i], a[i], shl1_from(ev[2 * i], ev[2 * i + 1]))
                              : madc_hi_cc(a[i], a[i], shl1_from(ev[2 * i], ev[2 * i + 1]));
  }
}
// One reduction step: T <- (T + m p) / 2^32 with m = -T p^-1 mod 2^32, T = E + O 2^32 as in mad_n_redc
template <class P> __device__ __forceinline__ void redc_step(uint32_t* E, uint32_t* O, bool first, const uint32_t* MOD) {
  constexpr int n = P::L;
  if (first) {
    uint32_t mi = E[0] * P::inv_rt();
    mul_n<n>(O, MOD + 1, mi);
    cmad_n<n>(E, MOD, mi);
    O[n - 1] = addc(O[n - 1], 0);
  } else {
    uint32_t mi = (E[0] + O[1]) * P::inv_rt();
    E[0] = add_cc(E[0], O[1]);
    madc_n_rshift<n>(O, MOD + 1, mi);
    cmad_n<n>(E, MOD, mi);
    O[n - 1] = addc(O[n - 1], 0);
  }
}
// r = a^2 R^-1 mod p, input and output fully reduced
template <class P> __device__ __forceinline__ void mont_sqr(uint32_t* r, const uint32_t* a) {
  constexpr int n = P::L;
  uint32_t w[2 * n], odd[n], mod[n];
  sqr_wide<n>(w, a);
  load_modulus<P>(mod);
  uint32_t* even = w;                                // the low half is the window the reduction works on
#pragma unroll
  for (int i = 0; i < n; i += 2) {
    redc_step<P>(even, odd, i == 0, mod);
    redc_step<P>(odd, even, false, mod);
  }
  even[0] = add_cc(even[0], odd[1]);
#pragma unroll
  for (int k = 1; k < n - 1; k++) even[k] = addc_cc(even[k], odd[k + 1]);
  even[n - 1] = addc(even[n - 1], 0);
  // + the untouched high half of a^2: (a^2 + M p) / R = REDC(low half) + high half < 2p
  limbs_add<n>(even, even, w + n);
  uint32_t t[n];
  uint32_t borrow = limbs_sub<n>(t, even, mod);
#pragma unroll
  for (int k = 0; k < n; k++) r[k] = borrow ? even[k] : t[k];
}

// ---------------------------------------------------------------------------------------------
// Field wrapper: elements are structs of L limbs in Montgomery form
// ---------------------------------------------------------------------------------------------
template <class P_> struct Fp {
  using P = P_;
  static constexpr int L = P::L;
  static constexpr int DEG = 1;
  static constexpr int COOP = 0;                    // lanes per element (coop.cuh); 0 = one thread per element
  static constexpr int NBYTES = P::NBYTES;          // serialized size
  static constexpr int WORDS = L;                   // uint32 words per element in device arrays
  struct T { uint32_t v[L]; };
  using Base = Fp<P_>;

  __device__ __forceinline__ static T zero() { T r;
#pragma unroll
    for (int i = 0; i < L; i++) r.v[i] = 0;
    return r; }
  __device__ __forceinline__ static T one() { T r;
#pragma unroll
    for (int i = 0; i < L; i++) r.v[i] = P::r1()[i];
    return r; }
  __device__ __forceinline__ static bool is_zero(const T& a) { return limbs_is_zero<L>(a.v); }
  __device__ __forceinline__ static bool eq(const T& a, const T& b) { return limbs_eq<L>(a.v, b.v); }

  __device__ __forceinline__ static T add(const T& a, const T& b) {
    T s, t, r;
    limbs_add<L>(s.v, a.v, b.v);                    // no overflow: p has spare top bits
    uint32_t borrow = limbs_sub<L>(t.v, s.v, P::p());
#pragma unroll
    for (int i = 0; i < L; i++) r.v[i] = borrow ? s.v[i] : t.v[i];
    return r;
  }
  __device__ __forceinline__ static T dbl(const T& a) { return add(a, a); }
  __device__ __forceinline__ static T sub(const T& a, const T& b) {
    T d, t, r;
    uint32_t borrow = limbs_sub<L>(d.v, a.v, b.v);
    limbs_add<L>(t.v, d.v, P::p());
#pragma unroll
    for (int i = 0; i < L; i++) r.v[i] = borrow ? t.v[i] : d.v[i];
    return r;
  }
  __device__ __forceinline__ static T neg(const T& a) {
    T t;
    limbs_sub<L>(t.v, P::p(), a.v);
    bool z = is_zero(a);
#pragma unroll
    for (int i = 0; i < L; i++) t.v[i] = z ? 0u : t.v[i];
    return t;
  }
  // ---- unreduced ("lazy") operands --------------------------------------------------------------------------------
  // mont_mul does not need reduced inputs: for a < A p, b < B p the CIOS accumulator stays below (A + 1) p and the result is
  // < a b / R + p, i.e. < 2 p whenever A B p < R — its final conditional subtraction then returns the canonical residue.
  // The base fields of the curves leave 7 (377 / 761 bits) or 15 (753 bits) spare bits in R = 2^(32 L), so sums and small
  // multiples of reduced elements (A B <= 32) can feed a multiplication WITHOUT the compare-subtract-select of add() /
  // mul_small(): one carry chain instead of up to six.  Only for products of the extension-field formulas (ext.cuh, coop.cuh);
  // every value that leaves them is canonical again.
  static constexpr int SPARE_BITS = 32 * L - P::BITS;
  // May the point formulas (ec.cuh, SSO_LAZY_EC) hand unreduced sums to mul / sqr of this field?  R / p >= 128 is enough for them.
  static constexpr bool LAZY_OK = SPARE_BITS >= 7;
  // in a point formula, is a squaring worth keeping where a multiplication would save modular additions?  (dedicated squaring)
  static constexpr bool SQR_CHEAPER = L <= SSO_SQR_EC_MAX_L && !SSO_LAZY_TRADE_ALL;
  // a + b as integers (< 2 p)
  __device__ __forceinline__ static T add_lazy(const T& a, const T& b) { T r; limbs_add<L>(r.v, a.v, b.v); return r; }
  // p - a as an integer (a < p; p for a = 0: another representative of 0)
  __device__ __forceinline__ static T neg_lazy(const T& a) { T r; limbs_sub<L>(r.v, P::p(), a.v); return r; }
  // a - b + p as an integer (a < A p, b <= p: the result is below (A + 1) p and positive)
  __device__ __forceinline__ static T sub_lazy(const T& a, const T& b) {
    T n, r;
    limbs_sub<L>(n.v, P::p(), b.v);
    limbs_add<L>(r.v, a.v, n.v);
    return r;
  }
  // c + k x as integers; the caller guarantees (k + 1) p < R
  __device__ __forceinline__ static T mad_small_lazy(const T& c, uint32_t k, const T& x) {
    uint32_t ev[L], od[L];
    mul_n<L>(ev, x.v, k);                            // k x[j] for even j: disjoint 64-bit pairs at limbs j, j + 1
    mul_n<L>(od, x.v + 1, k);                        // k x[j] for odd j: pairs at limbs j, j + 1 (od[] starts at limb 1)
    T r;
    r.v[0] = add_cc(c.v[0], ev[0]);
#pragma unroll
    for (int i = 1; i < L - 1; i++) r.v[i] = addc_cc(c.v[i], ev[i]);
    r.v[L - 1] = addc(c.v[L - 1], ev[L - 1]);
    r.v[1] = add_cc(r.v[1], od[0]);
#pragma unroll
    for (int i = 2; i < L - 1; i++) r.v[i] = addc_cc(r.v[i], od[i - 1]);
    r.v[L - 1] = addc(r.v[L - 1], od[L - 2]);         // od[L - 1] = high word of k x[L - 1] = 0 by the no-overflow guarantee
    return r;
  }
  // canonical residue of y < 16 p: quotient estimate from the top limbs (never too large, at most one too small because the
  // top limb of p has >= 17 bits), one multiple of p subtracted, one conditional subtraction
  __device__ __forceinline__ static T reduce_small(const T& y) {
    static_assert(SPARE_BITS >= 7 && SPARE_BITS <= 15, "top limb of p must hold at least 17 bits, and 16 p < R");
    uint32_t q = y.v[L - 1] / (P::p()[L - 1] + 1u);
    uint32_t ev[L], od[L], mod[L];
    load_modulus<P>(mod);
    mul_n<L>(ev, mod, q);
    mul_n<L>(od, mod + 1, q);
    T d;
    d.v[0] = sub_cc(y.v[0], ev[0]);
#pragma unroll
    for (int i = 1; i < L - 1; i++) d.v[i] = subc_cc(y.v[i], ev[i]);
    d.v[L - 1] = subc(y.v[L - 1], ev[L - 1]);
    d.v[1] = sub_cc(d.v[1], od[0]);
#pragma unroll
    for (int i = 2; i < L - 1; i++) d.v[i] = subc_cc(d.v[i], od[i - 1]);
    d.v[L - 1] = subc(d.v[L - 1], od[L - 2]);
    T t, r;
    uint32_t borrow = limbs_sub<L>(t.v, d.v, mod);
#pragma unroll
    for (int i = 0; i < L; i++) r.v[i] = borrow ? d.v[i] : t.v[i];
    return r;
  }

  // 24-limb multiplications are ~2.5k instructions each: keep one out-of-line copy so that point
  // formulas do not overflow the instruction cache; 8/12-limb ones are inlined.
  // by-value parameters: the device ABI passes them in registers (by-reference would round-trip through local memory)
  __device__ __noinline__ static T mul_outlined(T a, T b) { T r; mont_mul<P>(r.v, a.v, b.v); return r; }
  __device__ __forceinline__ static T mul(const T& a, const T& b) {
    if constexpr (L > SSO_INLINE_MUL_MAX_L) { return mul_outlined(a, b); } else { T r; mont_mul<P>(r.v, a.v, b.v); return r; }
  }
  // Dedicated squaring.  Measured on B200 (tools/mulbench, profiles/r1_mulbench.txt): +13 % over mul(a, a) on 12 limbs,
  // +22 % on 24 limbs in isolation; inside the point formulas the 24-limb version does not pay (the out-of-line
  // squaring needs more registers than the multiplication and the callers spill around it: batch_exp -1..3 %), so the
  // point formulas use it up to SSO_SQR_EC_MAX_L limbs and the exponentiation chains (sqrt, pow) always.
#ifndef SSO_SQR_EC_MAX_L
#define SSO_SQR_EC_MAX_L 12
#endif
  __device__ __noinline__ static T sqr_outlined(T a) { T r; mont_sqr<P>(r.v, a.v); return r; }
  __device__ __forceinline__ static T sqr_dedicated(const T& a) {
    if constexpr (L > SSO_INLINE_MUL_MAX_L) { return sqr_outlined(a); } else { T r; mont_sqr<P>(r.v, a.v); return r; }
  }
  __device__ __forceinline__ static T sqr(const T& a) {
    if constexpr (L <= SSO_SQR_EC_MAX_L) return sqr_dedicated(a); else return mul(a, a);
  }
  // always-inlined multiplication, for callers that want independent multiplications interleaved by the scheduler
  __device__ __forceinline__ static T mul_inl(const T& a, const T& b) { T r; mont_mul<P>(r.v, a.v, b.v); return r; }
  // multiply by a small non-negative integer constant
  template <int K> __device__ __forceinline__ static T mul_small(const T& a) {
    static_assert(K >= 0 && K < 64, "small constant");
    T acc = zero();
    T base = a;
    bool started = false;
#pragma unroll
    for (int bit = 0; bit < 6; bit++) {
      if ((K >> bit) & 1) { acc = started ? add(acc, base) : base; started = true; }
      if ((K >> (bit + 1)) != 0) base = dbl(base);
    }
    return acc;
  }

  // Montgomery <-> canonical
  __device__ __forceinline__ static T to_mont(const T& a) { T r2;
#pragma unroll
    for (int i = 0; i < L; i++) r2.v[i] = P::r2()[i];
    return mul(a, r2); }
  __device__ __forceinline__ static T from_mont(const T& a) { T o = zero(); o.v[0] = 1; return mul(a, o); }

  // a^e for a constant-memory exponent of L limbs (not inlined: large, rarely on the hot path)
  __device__ __noinline__ static T pow_const(const T& a, const uint32_t* e) {
    T r = one();
    bool started = false;
    for (int i = L * 32 - 1; i >= 0; i--) {
      if (started) r = sqr_dedicated(r);
      if ((e[i >> 5] >> (i & 31)) & 1) { r = started ? mul(r, a) : a; started = true; }
    }
    return r;
  }
  // Fermat inversion a^(p-2): ~1.5 * bits multiplications (kept for cross-checking the binary algorithm)
  __device__ __forceinline__ static T inv_fermat(const T& a) { return pow_const(a, P::pm2()); }

  // Modular inversion by the binary extended Euclidean algorithm on the raw residue (no multiplications:
  // shifts, subtractions and comparisons only — ~5x fewer instructions than Fermat, which matters where ONE
  // thread inverts for a whole block).  Input/output in Montgomery form; inv(0) = 0.
  //   raw = a R  ->  raw^-1 = a^-1 R^-1 ; two multiplications by R^2 give a^-1 R.
  __device__ __noinline__ static T inv(const T& a) {
    if (is_zero(a)) return a;
    uint32_t u[L], v[L], x1[L], x2[L], t[L];
#pragma unroll
    for (int i = 0; i < L; i++) { u[i] = a.v[i]; v[i] = P::p()[i]; x1[i] = i == 0 ? 1u : 0u; x2[i] = 0u; }
    // halve x modulo p: x even -> x/2, x odd -> (x + p)/2   (x + p < 2^(32 L): p has spare top bits)
    auto halve_mod = [&](uint32_t* x) {
      if (x[0] & 1u) limbs_add<L>(x, x, P::p());
#pragma unroll
      for (int i = 0; i < L - 1; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
      x[L - 1] >>= 1;
    };
    auto shr1 = [&](uint32_t* x) {
#pragma unroll
      for (int i = 0; i < L - 1; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
      x[L - 1] >>= 1;
    };
    auto is_one = [&](const uint32_t* x) {
      uint32_t o = x[0] ^ 1u;
#pragma unroll
      for (int i = 1; i < L; i++) o |= x[i];
      return o == 0;
    };
    while (!is_one(u) && !is_one(v)) {
      while ((u[0] & 1u) == 0) { shr1(u); halve_mod(x1); }
      while ((v[0] & 1u) == 0) { shr1(v); halve_mod(x2); }
      // u, v odd
      if (limbs_sub<L>(t, u, v) == 0) {             // u >= v
#pragma unroll
        for (int i = 0; i < L; i++) u[i] = t[i];
        if (limbs_sub<L>(x1, x1, x2) != 0) limbs_add<L>(x1, x1, P::p());
      } else {
        limbs_sub<L>(v, v, u);
        if (limbs_sub<L>(x2, x2, x1) != 0) limbs_add<L>(x2, x2, P::p());
      }
    }
    T r;
    bool use1 = is_one(u);
#pragma unroll
    for (int i = 0; i < L; i++) r.v[i] = use1 ? x1[i] : x2[i];
    T r2;
#pragma unroll
    for (int i = 0; i < L; i++) r2.v[i] = P::r2()[i];
    return mul(mul(r, r2), r2);
  }

  // canonical-order comparison "a > -a" used by the serialisation sign flag
  __device__ __forceinline__ static bool is_neg_canonical(const T& canon) { return limbs_gt<L>(canon.v, P::half()); }
  // lexicographic helper for extension fields: -1 / 0 / +1 for "a vs -a" on a canonical coefficient
  __device__ __forceinline__ static int sign_canonical(const T& canon) {
    if (is_zero(canon)) return 0;
    return is_neg_canonical(canon) ? 1 : -1;
  }
  // y > -y (in the field's canonical ordering); input in Montgomery form
  __device__ __forceinline__ static bool lex_is_neg(const T& a) { return is_neg_canonical(from_mont(a)); }

  // square root (some root) — returns false if a is a non-residue.  p = 3 mod 4 -> one exponentiation,
  // otherwise Tonelli-Shanks over the 2-Sylow subgroup.
  __device__ __noinline__ static bool sqrt(const T& a, T& out) {
    if (is_zero(a)) { out = a; return true; }
    if (P::P_IS_3_MOD_4) {
      // (p+1)/4 = (p-3)/4 + 1 = tm1h ... use: w = a^((p-3)/4), x = a*w, check x^2 == a
      T w = pow_const(a, P::tm1h());                // t = (p-1)/2 odd, (t-1)/2 = (p-3)/4
      T x = mul(a, w);
      out = x;
      return eq(sqr(x), a);
    }
    // w = a^((t-1)/2); x = a w; b = x w  (= a^t)
    T w = pow_const(a, P::tm1h());
    T x = mul(a, w);
    T b = mul(x, w);
    T z;
#pragma unroll
    for (int i = 0; i < L; i++) z.v[i] = P::tsz()[i];
    int m = P::TWO_ADICITY;
    T o = one();
    while (!eq(b, o)) {
      int k = 0;
      T b2 = b;
      while (!eq(b2, o)) { b2 = sqr(b2); k++; if (k >= m) return false; }   // non-residue
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

  // ---- byte formats (ark-serialize 0.4.2: canonical little-endian, ceil(bits/8) bytes) ----
  // Reads NBYTES bytes; the two top bits of the last byte are returned in `flags` and masked off
  // when with_flags.  Returns false when the integer is >= p.  Output in Montgomery form.
  __device__ __forceinline__ static bool from_bytes(const uint8_t* src, bool with_flags, uint32_t& flags, T& out) {
    T c = zero();
#pragma unroll
    for (int i = 0; i < NBYTES; i++) {
      uint32_t byte = src[i];
      if (i == NBYTES - 1) {
        flags = with_flags ? (byte & 0xC0u) : 0u;
        if (with_flags) byte &= 0x3Fu;
      }
      c.v[i >> 2] |= byte << ((i & 3) * 8);
    }
    T t;
    uint32_t borrow = limbs_sub<L>(t.v, c.v, P::p());
    out = to_mont(c);
    return borrow != 0;                             // c < p
  }
  __device__ __forceinline__ static void to_bytes(uint8_t* dst, const T& a, uint32_t flags) {
    T c = from_mont(a);
#pragma unroll
    for (int i = 0; i < NBYTES; i++) {
      uint32_t byte = (c.v[i >> 2] >> ((i & 3) * 8)) & 0xFFu;
      if (i == NBYTES - 1) byte |= flags;
      dst[i] = (uint8_t)byte;
    }
  }
  // word-array I/O for device-resident SoA/AoS buffers (Montgomery form)
  __device__ __forceinline__ static T load(const uint32_t* p, size_t stride) { T r;
#pragma unroll
    for (int i = 0; i < L; i++) r.v[i] = p[i * stride];
    return r; }
  __device__ __forceinline__ static void store(uint32_t* p, size_t stride, const T& a) {
#pragma unroll
    for (int i = 0; i < L; i++) p[i * stride] = a.v[i];
  }
  __device__ __forceinline__ static T from_const(const uint32_t* c) { T r;
#pragma unroll
    for (int i = 0; i < L; i++) r.v[i] = c[i];
    return r; }
};

}  // namespace sso
