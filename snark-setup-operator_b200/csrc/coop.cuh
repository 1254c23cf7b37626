// Warp-cooperative extension fields: ONE ELEMENT OF Fq^k IS HELD BY k ADJACENT LANES, one base-field coefficient per lane.
//
// The one-thread-per-element G2 bodies of the 753-bit curves keep a Jacobian point of 144 (Fq2) / 216 (Fq3) limbs plus the
// temporaries of the point formulas per thread: 255 registers and 11-20 KB of stack per thread, and the extension-field half
// of a chunk runs at 0.25-0.30 of the multiply-accumulate peak while the single-coefficient G1 bodies reach 0.5
// (DESIGN.md §4).  Here every lane carries exactly what a G1 thread carries — one coefficient of every coordinate — and
// the k lanes of a group run the SAME instruction stream:
//   * additions, subtractions, small multiples, multiplications by base-field constants: lane-local, no exchange;
//   * Fq2 multiplication: each lane forms a_r b_r and a_r b_(1-r) (two base multiplications per lane, four per element
//     instead of Karatsuba's three) and the lanes swap one product;
//   * Fq2 squaring (complex method): lane 0 forms (a0 + a1)(a0 + nr a1), lane 1 forms a0 a1 — ONE multiplication per lane,
//     the same two per element as the single-thread version: the squaring-heavy doublings lose nothing;
//   * Fq3 multiplication (Karatsuba, six base multiplications): two per lane — v_r = a_r b_r and
//     s_r = (a_r + a_(r+1))(b_r + b_(r+1)) — then two rotations of one coefficient each;
//   * predicates (is_zero, eq) are AND-reduced over the group (__all_sync on the group's lanes), so every data-dependent
//     branch of the point formulas is uniform inside a group and the shuffles below stay convergent.
// Exchanges are __shfl_sync over the group's own lane mask: groups of one warp may diverge from each other (different window
// digits) exactly as single threads do in the other kernels.
//
// This is the "warp-cooperative multi-thread-per-element layout" BASELINE.json's north_star names for the wide fields,
// applied where the measurements put the register problem: the extension-field groups.
// No counterpart in the reference (rayon over points, one core per point).
#pragma once
#include "ext.cuh"
#ifdef SSO_HOST_EMUL
#include <atomic>
#include <thread>
#include <functional>
#endif

namespace sso {

#ifndef SSO_HOST_EMUL
template <int DEG> struct Coop {
  static_assert(DEG == 2 || DEG == 3, "two or three lanes per element");
  // lanes per warp that carry points: 32 for pairs, 30 for triples (lanes 30, 31 idle)
  static constexpr uint32_t GROUPS_PER_WARP = 32 / DEG;
  __device__ __forceinline__ static uint32_t lane() { return threadIdx.x & 31u; }
  __device__ __forceinline__ static uint32_t role() { return DEG == 2 ? (threadIdx.x & 1u) : (threadIdx.x & 31u) % 3u; }
  __device__ __forceinline__ static uint32_t base() { return lane() - role(); }
  __device__ __forceinline__ static uint32_t mask() { return ((1u << DEG) - 1u) << base(); }
  __device__ __forceinline__ static bool lane_active() { return lane() < GROUPS_PER_WARP * DEG; }
  // point handled by this lane's group, counted inside the thread block
  __device__ __forceinline__ static uint32_t group_in_block() { return (threadIdx.x >> 5) * GROUPS_PER_WARP + lane() / DEG; }
  // out = the N words `in` of the lane whose role is (role + k) mod DEG
  template <int N> __device__ __forceinline__ static void rot(uint32_t* out, const uint32_t* in, uint32_t k) {
    uint32_t r = role() + k;
    if (r >= DEG) r -= DEG;
    const uint32_t src = base() + r, m = mask();
#pragma unroll
    for (int i = 0; i < N; i++) out[i] = __shfl_sync(m, in[i], src);
  }
  // the word of the lane with role `from`
  __device__ __forceinline__ static uint32_t get(uint32_t v, uint32_t from) { return __shfl_sync(mask(), v, base() + from); }
  __device__ __forceinline__ static bool all(bool c) { return __all_sync(mask(), c) != 0; }
};
#else
// TEST-ONLY emulation: the DEG lanes of a group are DEG host threads in lockstep; an exchange is a pair of barriers around a
// shared slot array (tests/emul).  Never part of the product library.
struct CoopEmuState {
  std::atomic<int> count{0};
  std::atomic<int> sense{0};
  const uint32_t* slot[3] = {nullptr, nullptr, nullptr};
  uint32_t flag[3] = {0, 0, 0};
  int lanes = 1;
};
inline CoopEmuState& coop_emu() { static CoopEmuState s; return s; }
static thread_local int g_coop_role = 0;
static thread_local int g_coop_sense = 0;
inline void coop_emu_barrier() {
  CoopEmuState& s = coop_emu();
  g_coop_sense ^= 1;
  if (s.count.fetch_add(1) + 1 == s.lanes) { s.count.store(0); s.sense.store(g_coop_sense); }
  else while (s.sense.load() != g_coop_sense) std::this_thread::yield();
}
// runs fn(role) on DEG lockstep threads
inline void coop_emu_run(int deg, const std::function<void(int)>& fn) {
  CoopEmuState& s = coop_emu();
  s.lanes = deg; s.count.store(0); s.sense.store(0);
  std::vector<std::thread> th;
  for (int r = 0; r < deg; r++) th.emplace_back([&, r] { g_coop_role = r; g_coop_sense = 0; fn(r); });
  for (auto& t : th) t.join();
}
template <int DEG> struct Coop {
  static uint32_t role() { return (uint32_t)g_coop_role; }
  template <int N> static void rot(uint32_t* out, const uint32_t* in, uint32_t k) {
    CoopEmuState& s = coop_emu();
    uint32_t tmp[N];
    s.slot[g_coop_role] = in;
    coop_emu_barrier();
    const uint32_t* src = s.slot[(g_coop_role + k) % DEG];
    for (int i = 0; i < N; i++) tmp[i] = src[i];
    coop_emu_barrier();
    for (int i = 0; i < N; i++) out[i] = tmp[i];
  }
  static uint32_t get(uint32_t v, uint32_t from) {
    CoopEmuState& s = coop_emu();
    s.flag[g_coop_role] = v;
    coop_emu_barrier();
    uint32_t r = s.flag[from];
    coop_emu_barrier();
    return r;
  }
  static bool all(bool c) {
    CoopEmuState& s = coop_emu();
    s.flag[g_coop_role] = c ? 1u : 0u;
    coop_emu_barrier();
    bool r = true;
    for (int i = 0; i < DEG; i++) r = r && s.flag[i] != 0;
    coop_emu_barrier();
    return r;
  }
};
#endif

// select without divergence
template <int N> __device__ __forceinline__ void sel_words(uint32_t* out, bool c, const uint32_t* a, const uint32_t* b) {
#pragma unroll
  for (int i = 0; i < N; i++) out[i] = c ? a[i] : b[i];
}

// ---------------------------------------------------------------------------------------------------------------------
// Fq2 = Fq[u] / (u^2 - nr) over two lanes: lane r holds c_r
// ---------------------------------------------------------------------------------------------------------------------
template <class B_, int K, bool NEG> struct CFp2 {
  using B = B_;
  using Base = B_;
  using NR = SmallNR<B_, K, NEG>;
  using CO = Coop<2>;
  static constexpr int DEG = 2;
  static constexpr int COOP = 2;
  // mul_val / sqr_val only ever feed their arguments to base multiplications (through unreduced sums themselves), so the point
  // formulas may pass unreduced arguments: with coefficients below A p and B p the largest product is 13 A B p^2 (mul) and
  // 28 A^2 p^2 (sqr) — inside p R for the bounds ec.cuh produces on the 753-bit tower (A <= 29, A B <= 58; R / p = 37054).
  // A negative non-residue would need p - x of an unreduced x: canonical arguments only.
  static constexpr bool LAZY_OK = !NEG && B_::SPARE_BITS >= 15;
  static constexpr bool SQR_CHEAPER = true;            // one multiplication round against two
  static constexpr int L = B::L;
  static constexpr int NBYTES = 2 * B::NBYTES;
  static constexpr int WORDS = 2 * B::L;              // words of a whole element in device arrays
  using T = typename B::T;                             // this lane's coefficient

  __device__ __forceinline__ static T pick(bool c, const T& a, const T& b) { T r; sel_words<L>(r.v, c, a.v, b.v); return r; }
  __device__ __forceinline__ static T other(const T& a) { T o; CO::template rot<L>(o.v, a.v, 1); return o; }

  __device__ __forceinline__ static T zero() { return B::zero(); }
  __device__ __forceinline__ static T one() { return pick(CO::role() == 0, B::one(), B::zero()); }
  __device__ __forceinline__ static bool is_zero(const T& a) { return CO::all(B::is_zero(a)); }
  __device__ __forceinline__ static bool eq(const T& a, const T& b) { return CO::all(B::eq(a, b)); }
  __device__ __forceinline__ static T add(const T& a, const T& b) { return B::add(a, b); }
  __device__ __forceinline__ static T sub(const T& a, const T& b) { return B::sub(a, b); }
  __device__ __forceinline__ static T dbl(const T& a) { return B::dbl(a); }
  __device__ __forceinline__ static T neg(const T& a) { return B::neg(a); }
  template <int M> __device__ __forceinline__ static T mul_small(const T& a) { return B::template mul_small<M>(a); }
  __device__ __forceinline__ static T add_lazy(const T& a, const T& b) { return B::add_lazy(a, b); }
  __device__ __forceinline__ static T sub_lazy(const T& a, const T& b) { return B::sub_lazy(a, b); }
  __device__ __forceinline__ static T mad_small_lazy(const T& c, uint32_t k, const T& x) { return B::mad_small_lazy(c, k, x); }

  // (a0 + a1 u)(b0 + b1 u), c0 = a0 b0 + nr a1 b1, c1 = a0 b1 + a1 b0: lane r forms a_r (k_r b_r) with k_0 = 1, k_1 = nr — the
  // non-residue rides on an UNREDUCED operand (fp.cuh "lazy operands"), so the multiplication itself reduces nr a1 b1 — and
  // a_r b_(1-r); the lanes swap one product and finish with one addition
  __device__ __noinline__ static T mul_val(T a, T b) {
    const bool hi = CO::role() != 0;
    T ob = other(b);
    T x = NEG ? pick(hi, B::neg_lazy(b), b) : b;
    T kb = B::mad_small_lazy(B::zero(), hi ? (uint32_t)K : 1u, x);
    T ps = B::mul(a, kb);                             // lane 0: a0 b0, lane 1: nr a1 b1
    T pc = B::mul(a, ob);                             // a_r b_(1-r)
    T recv = other(pick(hi, ps, pc));                 // lane 0 receives nr a1 b1, lane 1 receives a0 b1
    return B::add(pick(hi, pc, ps), recv);
  }
  // complex squaring, ONE base multiplication per lane: lane 0 forms P = (a0 + a1)(a0 + nr a1), lane 1 forms w = (2 a1) a0;
  // c0 = P - (nr + 1) a0 a1 = P - h w with h = (nr + 1) / 2 (nr is odd), c1 = w.  All four factors are unreduced operands.
  __device__ __noinline__ static T sqr_val(T a) {
    static_assert((K & 1) == 1, "odd non-residue");
    constexpr int H = (NEG ? K - 1 : K + 1) / 2;      // |h|; the sign of -h is + for a negative non-residue
    const bool hi = CO::role() != 0;
    T o = other(a);
    T X = B::add_lazy(a, pick(hi, a, o));              // lane 0: a0 + a1, lane 1: 2 a1
    T Y = B::mad_small_lazy(pick(hi, o, a), hi ? 0u : (uint32_t)K, NEG ? B::neg_lazy(o) : o);   // lane 0: a0 + nr a1, lane 1: a0
    T P = B::mul(X, Y);
    T w = other(P);                                    // lane 0 receives 2 a0 a1
    return B::reduce_small(B::mad_small_lazy(P, hi ? 0u : (uint32_t)H, NEG ? w : B::neg_lazy(w)));
  }
  __device__ __forceinline__ static T mul(const T& a, const T& b) { return mul_val(a, b); }
  __device__ __forceinline__ static T sqr(const T& a) { return sqr_val(a); }
  __device__ __forceinline__ static T mul_base(const T& a, const typename B::T& k) { return B::mul(a, k); }
  __device__ __forceinline__ static T conj(const T& a) { return pick(CO::role() != 0, B::neg(a), a); }
  // 1 / (a0 + a1 u) = (a0 - a1 u) / (a0^2 - nr a1^2): the norm and its inverse are formed by both lanes (same data, same path)
  __device__ __noinline__ static T inv(const T& a) {
    const bool hi = CO::role() != 0;
    T sq = B::sqr(a);
    T osq = other(sq);
    T n = B::sub(pick(hi, osq, sq), NR::mul(pick(hi, sq, osq)));
    T ni = B::inv(n);
    T r = B::mul(a, ni);
    return pick(hi, B::neg(r), r);
  }
  // serialized element: c0 | c1, flag bits in the last byte of c1
  __device__ __forceinline__ static bool from_bytes(const uint8_t* src, bool with_flags, uint32_t& flags, T& out) {
    const uint32_t r = CO::role();
    uint32_t f = 0;
    bool ok = B::from_bytes(src + r * B::NBYTES, with_flags && r == 1, f, out);
    flags = CO::get(f, 1);
    return CO::all(ok);
  }
  __device__ __forceinline__ static T load(const uint32_t* p, size_t stride) { return B::load(p + (size_t)CO::role() * L * stride, stride); }
  __device__ __forceinline__ static void store(uint32_t* p, size_t stride, const T& a) { B::store(p + (size_t)CO::role() * L * stride, stride, a); }
  __device__ __forceinline__ static T from_const(const uint32_t* c) { return B::from_const(c + CO::role() * L); }
};

// ---------------------------------------------------------------------------------------------------------------------
// Fq3 = Fq[u] / (u^3 - nr) over three lanes: lane r holds c_r
// ---------------------------------------------------------------------------------------------------------------------
template <class B_, int K, bool NEG> struct CFp3 {
  using B = B_;
  using Base = B_;
  using NR = SmallNR<B_, K, NEG>;
  using CO = Coop<3>;
  static constexpr int DEG = 3;
  static constexpr int COOP = 3;
  // as for CFp2: arguments below A p and B p give products up to 4 A B p^2 (ec.cuh keeps A, B <= 8 here)
  static constexpr bool LAZY_OK = B_::SPARE_BITS >= 15;
  static constexpr bool SQR_CHEAPER = false;           // squaring = multiplication
  static constexpr int L = B::L;
  static constexpr int NBYTES = 3 * B::NBYTES;
  static constexpr int WORDS = 3 * B::L;
  using T = typename B::T;

  __device__ __forceinline__ static T pick(bool c, const T& a, const T& b) { T r; sel_words<L>(r.v, c, a.v, b.v); return r; }
  __device__ __forceinline__ static T from(const T& a, uint32_t k) { T o; CO::template rot<L>(o.v, a.v, k); return o; }

  __device__ __forceinline__ static T zero() { return B::zero(); }
  __device__ __forceinline__ static T one() { return pick(CO::role() == 0, B::one(), B::zero()); }
  __device__ __forceinline__ static bool is_zero(const T& a) { return CO::all(B::is_zero(a)); }
  __device__ __forceinline__ static bool eq(const T& a, const T& b) { return CO::all(B::eq(a, b)); }
  __device__ __forceinline__ static T add(const T& a, const T& b) { return B::add(a, b); }
  __device__ __forceinline__ static T sub(const T& a, const T& b) { return B::sub(a, b); }
  __device__ __forceinline__ static T dbl(const T& a) { return B::dbl(a); }
  __device__ __forceinline__ static T neg(const T& a) { return B::neg(a); }
  template <int M> __device__ __forceinline__ static T mul_small(const T& a) { return B::template mul_small<M>(a); }
  __device__ __forceinline__ static T add_lazy(const T& a, const T& b) { return B::add_lazy(a, b); }
  __device__ __forceinline__ static T sub_lazy(const T& a, const T& b) { return B::sub_lazy(a, b); }
  __device__ __forceinline__ static T mad_small_lazy(const T& c, uint32_t k, const T& x) { return B::mad_small_lazy(c, k, x); }

  // Karatsuba with v_r = a_r b_r, s_r = (a_r + a_(r+1))(b_r + b_(r+1)), u_r = s_r - v_r (indices mod 3):
  //   c0 = v0 + nr (u1 - v2)      c1 = (u0 - v1) + nr v2      c2 = (u2 - v0) + v1
  __device__ __noinline__ static T mul_val(T a, T b) {
    static_assert(!NEG, "positive non-residue");
    const uint32_t r = CO::role();
    T v = B::mul(a, b);
    T s = B::mul(B::add_lazy(a, from(a, 1)), B::add_lazy(b, from(b, 1)));     // unreduced factors (< 2 p each): fp.cuh "lazy operands"
    T u = B::sub(s, v);
    T r1 = from(pick(r == 1, u, v), 1);                // lane 0: u1, lane 1: v2, lane 2: v0
    T r2 = from(pick(r == 0, u, v), 2);                // lane 0: v2, lane 1: u0, lane 2: v1
    T d = B::sub(pick(r == 0, r1, pick(r == 1, r2, u)), pick(r == 0, r2, pick(r == 1, v, r1)));
    // c0 = v + nr d, c1 = d + nr r1, c2 = d + r2: one multiply-add by a small constant and one reduction for all three lanes
    return B::reduce_small(B::mad_small_lazy(pick(r == 0, v, d), r == 2 ? 1u : (uint32_t)K, pick(r == 0, d, pick(r == 1, r1, r2))));
  }
  __device__ __forceinline__ static T mul(const T& a, const T& b) { return mul_val(a, b); }
  __device__ __forceinline__ static T sqr(const T& a) { return mul_val(a, a); }
  __device__ __forceinline__ static T mul_base(const T& a, const typename B::T& k) { return B::mul(a, k); }
  // 1 / a = (t0, t1, t2) / n with t0 = a0^2 - nr a1 a2, t1 = nr a2^2 - a0 a1, t2 = a1^2 - a0 a2, n = a0 t0 + nr (a2 t1 + a1 t2);
  // lane r forms t_r and its term of n, the norm and its inverse are formed by all three lanes (same data, same path).
  // Once per thread block (the root of the inversion tree): not tuned.
  __device__ __noinline__ static T inv(const T& a) {
    const uint32_t r = CO::role();
    T an = from(a, 1), ap = from(a, 2);                // a_(r+1), a_(r+2)
    T sq = B::sqr(pick(r == 0, a, pick(r == 1, an, ap)));                           // a0^2 | a2^2 | a1^2
    T pr = B::mul(pick(r == 1, ap, an), pick(r == 0, ap, a));                       // a1 a2 | a0 a1 | a0 a2
    T nrs = NR::mul(pick(r == 0, pr, sq));
    T t = B::sub(pick(r == 1, nrs, sq), pick(r == 0, nrs, pr));
    T m = B::mul(pick(r == 0, a, pick(r == 1, an, ap)), t);                         // a0 t0 | a2 t1 | a1 t2
    T mn = from(m, 1), mp = from(m, 2);
    T cand = B::add(m, NR::mul(B::add(mn, mp)));      // the norm, on lane 0
    T n = from(cand, r == 0 ? 0u : 3u - r);
    return B::mul(t, B::inv(n));
  }
  // (c0, c1 w1, c2 w2): the q-power Frobenius with the base-field constants w1, w2 of the tower
  __device__ __forceinline__ static T frob_w(const T& a, const uint32_t* w1, const uint32_t* w2) {
    const uint32_t r = CO::role();
    T m = B::mul(a, pick(r == 1, B::from_const(w1), B::from_const(w2)));
    return pick(r == 0, a, m);
  }
  // x -> x u^j (j = 1, 2) is a rotation of the coefficients with nr on the wrapped ones:
  //   x u = (nr c2, c0, c1)      x u^2 = (nr c1, nr c2, c0)
  __device__ __forceinline__ static T mul_u2(const T& a) {
    const uint32_t r = CO::role();
    T n = from(a, 1);                                  // lane 0: c1, lane 1: c2, lane 2: c0
    return pick(r == 2, n, NR::mul(n));
  }
  __device__ __forceinline__ static bool from_bytes(const uint8_t* src, bool with_flags, uint32_t& flags, T& out) {
    const uint32_t r = CO::role();
    uint32_t f = 0;
    bool ok = B::from_bytes(src + r * B::NBYTES, with_flags && r == 2, f, out);
    flags = CO::get(f, 2);
    return CO::all(ok);
  }
  __device__ __forceinline__ static T load(const uint32_t* p, size_t stride) { return B::load(p + (size_t)CO::role() * L * stride, stride); }
  __device__ __forceinline__ static void store(uint32_t* p, size_t stride, const T& a) { B::store(p + (size_t)CO::role() * L * stride, stride, a); }
  __device__ __forceinline__ static T from_const(const uint32_t* c) { return B::from_const(c + CO::role() * L); }
};

}  // namespace sso
