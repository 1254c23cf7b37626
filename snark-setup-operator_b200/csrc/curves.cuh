// The eight groups of the ceremony: G1 / G2 of BLS12-377, BW6-761, MNT4-753, MNT6-753.
//
// Mirrors the `curveKind` values the operator accepts (reference src/data_structs.rs:123-131,
// src/bin/new_setup.rs:53-54) and the curve crates ark-bls12-377 / ark-bw6-761 / ark-mnt4-753 /
// ark-mnt6-753 0.4.0 (Cargo.lock:150-151,173-174,282-283,293-294); constants from SURVEY.md A.1
// via tools/gen_constants.py.
#pragma once
#include "ec.cuh"
#include "coop.cuh"

namespace sso {

using Fr253 = Fp<P_r253>;
using Fq377 = Fp<P_q377>;
using Fq761 = Fp<P_q761>;
using Fq4 = Fp<P_q4>;      // MNT4-753 base field = MNT6-753 scalar field
using Fq6 = Fp<P_q6>;      // MNT6-753 base field = MNT4-753 scalar field

using Fq377x2 = Fp2<Fq377, SmallNR<Fq377, 5, true>>;     // u^2 = -5
using Fq4x2 = Fp2<Fq4, SmallNR<Fq4, 13, false>>;          // u^2 = 13
using Fq6x3 = Fp3<Fq6, SmallNR<Fq6, 11, false>>;          // u^3 = 11

// cofactor 1: the curve group itself has prime order r (subgroup membership = being on the curve)
static constexpr bool PRIME_ORDER_bls12_377_g1 = false, PRIME_ORDER_bls12_377_g2 = false, PRIME_ORDER_bw6_761_g1 = false,
                      PRIME_ORDER_bw6_761_g2 = false, PRIME_ORDER_mnt4_753_g1 = true, PRIME_ORDER_mnt4_753_g2 = false,
                      PRIME_ORDER_mnt6_753_g1 = true, PRIME_ORDER_mnt6_753_g2 = false;
#define SSO_GROUP_COMMON(NAME, FIELD, SCALAR)                                                      \
  using F = FIELD;                                                                                  \
  using Fr = SCALAR;                                                                                \
  __device__ __forceinline__ static typename F::T coeff_b() { return F::from_const(c_##NAME##_b); } \
  __device__ __forceinline__ static const uint32_t* order() { return c_##NAME##_order; }           \
  __device__ __forceinline__ static typename F::T gen_x() { return F::from_const(c_##NAME##_gx); }  \
  __device__ __forceinline__ static typename F::T gen_y() { return F::from_const(c_##NAME##_gy); }  \
  __device__ __forceinline__ static const uint32_t* cofactor() { return c_##NAME##_cofactor; }      \
  static constexpr int COFACTOR_WORDS = COFACTOR_WORDS_##NAME;                                     \
  static constexpr bool PRIME_ORDER = PRIME_ORDER_##NAME;

struct Bls12_377_G1 {
  static constexpr bool AFFINE_TABLE = true;
  static constexpr bool HAS_GLV = true;
  using Glv = GLV_bls12_377_g1;
  static constexpr int ENDO_SUBGROUP_TEST = 1;          // phi(P) = [-x^2]P  (ec.cuh::in_subgroup)
  static constexpr bool HAS_GLS4 = false;
  static constexpr bool HAS_GLS2 = false;
  using Endo = ENDO_bls12_377;
  static constexpr uint32_t GROUP = 0;
  SSO_GROUP_COMMON(bls12_377_g1, Fq377, Fr253)
  static constexpr bool A_IS_ZERO = true;
  __device__ __forceinline__ static F::T mul_a(const F::T&) { return F::zero(); }
  __device__ __forceinline__ static bool field_sqrt(const F::T& a, F::T& o) { return F::sqrt(a, o); }
};
struct Bls12_377_G2 {
  static constexpr float VERIFY_WEIGHT = 1.6f;            // curve_ops.cuh::CurveOps::verify_g2_weight
  static constexpr bool AFFINE_TABLE = true;
  static constexpr bool HAS_GLV = true;
  using Glv = GLV_bls12_377_g2;
  static constexpr int ENDO_SUBGROUP_TEST = 2;          // psi(P) = [x]P
  static constexpr bool HAS_GLS4 = true;                // batch_exp: k = k0 + k1 x + k2 x^2 + k3 x^3 with psi (ec.cuh)
  static constexpr bool HAS_GLS2 = false;
  using Endo = ENDO_bls12_377;
  static constexpr uint32_t GROUP = 1;
  SSO_GROUP_COMMON(bls12_377_g2, Fq377x2, Fr253)
  static constexpr bool A_IS_ZERO = true;
  __device__ __forceinline__ static F::T mul_a(const F::T&) { return F::zero(); }
  __device__ __forceinline__ static bool field_sqrt(const F::T& a, F::T& o) { return F::sqrt(a, o); }
};
struct Bw6_761_G1 {
  static constexpr bool AFFINE_TABLE = true;
  static constexpr bool HAS_GLV = true;
  using Glv = GLV_bw6_761_g1;
  static constexpr int ENDO_SUBGROUP_TEST = 3;          // [x + 1]P + [x^3 - x^2 + 1]phi(P) = O  (ec.cuh::in_subgroup)
  using Endo = ENDO_bw6_761;
  static constexpr bool HAS_GLS4 = false;
  static constexpr bool HAS_GLS2 = false;
  static constexpr uint32_t GROUP = 0;
  SSO_GROUP_COMMON(bw6_761_g1, Fq761, Fq377)
  static constexpr bool A_IS_ZERO = true;
  __device__ __forceinline__ static F::T mul_a(const F::T&) { return F::zero(); }
  __device__ __forceinline__ static bool field_sqrt(const F::T& a, F::T& o) { return F::sqrt(a, o); }
};
struct Bw6_761_G2 {
  static constexpr float VERIFY_WEIGHT = 1.5f;            // curve_ops.cuh::CurveOps::verify_g2_weight
  static constexpr bool AFFINE_TABLE = true;
  static constexpr bool HAS_GLV = true;
  using Glv = GLV_bw6_761_g2;
  static constexpr int ENDO_SUBGROUP_TEST = 0;
  static constexpr bool HAS_GLS4 = false;
  static constexpr bool HAS_GLS2 = false;
  static constexpr uint32_t GROUP = 1;
  SSO_GROUP_COMMON(bw6_761_g2, Fq761, Fq377)
  static constexpr bool A_IS_ZERO = true;
  __device__ __forceinline__ static F::T mul_a(const F::T&) { return F::zero(); }
  __device__ __forceinline__ static bool field_sqrt(const F::T& a, F::T& o) { return F::sqrt(a, o); }
};
struct Mnt4_753_G1 {
  static constexpr bool AFFINE_TABLE = true;
  static constexpr bool HAS_GLV = false;
  static constexpr int ENDO_SUBGROUP_TEST = 0;
  static constexpr bool HAS_GLS4 = false;
  static constexpr bool HAS_GLS2 = false;
  static constexpr uint32_t GROUP = 0;                                         // a = 2
  SSO_GROUP_COMMON(mnt4_753_g1, Fq4, Fq6)
  static constexpr bool A_IS_ZERO = false;
  __device__ __forceinline__ static F::T mul_a(const F::T& x) { return F::dbl(x); }
  __device__ __forceinline__ static F::T mad_a_lazy(const F::T& c, const F::T& x) { return F::mad_small_lazy(c, 2u, x); }   // c + a x, unreduced
  __device__ __forceinline__ static bool field_sqrt(const F::T& a, F::T& o) { return F::sqrt(a, o); }
};
struct Mnt4_753_G2 {
  static constexpr float VERIFY_WEIGHT = 10.f;            // curve_ops.cuh::CurveOps::verify_g2_weight
  static constexpr bool AFFINE_TABLE = true;
  static constexpr bool HAS_GLV = false;
  static constexpr int ENDO_SUBGROUP_TEST = 4;          // psi(P) = [t - 1]P  (ec.cuh::in_subgroup)
  using Endo = ENDO_mnt4_753;
  static constexpr bool HAS_GLS4 = false;
  static constexpr bool HAS_GLS2 = true;                 // batch_exp: k = k0 + k1 |t - 1| with the Frobenius endomorphism psi (ec.cuh)
  static constexpr uint32_t GROUP = 1;                                         // a' = (26, 0)
  SSO_GROUP_COMMON(mnt4_753_g2, Fq4x2, Fq6)
  static constexpr bool A_IS_ZERO = false;
  __device__ __forceinline__ static F::T mul_a(const F::T& x) { return F::mul_small<26>(x); }
  __device__ __forceinline__ static F::T mad_a_lazy(const F::T& c, const F::T& x) { return F::mad_small_lazy(c, 26u, x); }
  __device__ __forceinline__ static bool field_sqrt(const F::T& a, F::T& o) { return F::sqrt(a, o); }
};
struct Mnt6_753_G1 {
  static constexpr bool AFFINE_TABLE = true;
  static constexpr bool HAS_GLV = false;
  static constexpr int ENDO_SUBGROUP_TEST = 0;
  static constexpr bool HAS_GLS4 = false;
  static constexpr bool HAS_GLS2 = false;
  static constexpr uint32_t GROUP = 0;                                         // a = 11
  SSO_GROUP_COMMON(mnt6_753_g1, Fq6, Fq4)
  static constexpr bool A_IS_ZERO = false;
  __device__ __forceinline__ static F::T mul_a(const F::T& x) { return F::mul_small<11>(x); }
  __device__ __forceinline__ static F::T mad_a_lazy(const F::T& c, const F::T& x) { return F::mad_small_lazy(c, 11u, x); }
  __device__ __forceinline__ static bool field_sqrt(const F::T& a, F::T& o) { return F::sqrt(a, o); }
};
struct Mnt6_753_G2 {
  static constexpr float VERIFY_WEIGHT = 20.f;            // curve_ops.cuh::CurveOps::verify_g2_weight
  static constexpr bool AFFINE_TABLE = false;   // measured: the 72 KB Fq3 inversion tree costs more than mixed additions save
  static constexpr bool HAS_GLV = false;
  static constexpr int ENDO_SUBGROUP_TEST = 4;          // psi(P) = [t - 1]P  (ec.cuh::in_subgroup)
  using Endo = ENDO_mnt6_753;
  static constexpr bool HAS_GLS4 = false;
  static constexpr bool HAS_GLS2 = true;                 // batch_exp: k = k0 + k1 |t - 1| with the Frobenius endomorphism psi (ec.cuh)
  static constexpr uint32_t GROUP = 1;                                         // a' = (0, 0, 11) = 11 u^2, u^3 = 11
  SSO_GROUP_COMMON(mnt6_753_g2, Fq6x3, Fq4)
  static constexpr bool A_IS_ZERO = false;
  __device__ __forceinline__ static F::T mul_a(const F::T& x) {
    // x * u^2 = (11 c1, 11 c2, c0), then times 11
    return F::mul_small<11>(F::mul_u2(x));
  }
  __device__ __forceinline__ static F::T mad_a_lazy(const F::T& c, const F::T& x) { return F::add_lazy(c, mul_a(x)); }   // a x stays canonical
  __device__ __forceinline__ static bool field_sqrt(const F::T& a, F::T& o) {
#ifdef SSO_FQ3_SQRT_PLAIN
    return F::sqrt_ts<TS_q6x3>(a, o, c_q6x3_tm1h, c_q6x3_tsz);
#else
    return F::sqrt_ts_norm<TS_q6x3>(a, o, c_q6x3_tsz, ENDO_mnt6_753::w1(), ENDO_mnt6_753::w2());
#endif
  }
};

// ---- warp-cooperative variants of the extension-field groups (coop.cuh): the same curve, the same constants, one
// coefficient of every coordinate per lane.  Used by the batch_exp kernels on uncompressed input; everything else (point
// decompression, normalisation, MSM, pairing) keeps the one-thread-per-element types above.
#define SSO_GROUP_COOP(NAME, PLAIN, ...)                                                            \
  using Plain = PLAIN;                                                                              \
  using F = __VA_ARGS__;                                                                            \
  __device__ __forceinline__ static typename F::T coeff_b() { return F::from_const(c_##NAME##_b); } \
  __device__ __forceinline__ static bool field_sqrt(const typename F::T&, typename F::T&) { return false; }
struct Bls12_377_G2C : Bls12_377_G2 {
  static constexpr bool COOP_DEFAULT = false;          // measured: 28.9 ms per chunk against 27.5 ms (profiles/r2_coop_ab.txt)
  SSO_GROUP_COOP(bls12_377_g2, Bls12_377_G2, CFp2<Fq377, 5, true>)
  __device__ __forceinline__ static F::T mul_a(const F::T&) { return F::zero(); }
};
struct Mnt4_753_G2C : Mnt4_753_G2 {
  static constexpr bool COOP_DEFAULT = true;           // 462 ms per chunk against 515 ms
  SSO_GROUP_COOP(mnt4_753_g2, Mnt4_753_G2, CFp2<Fq4, 13, false>)
  __device__ __forceinline__ static F::T mul_a(const F::T& x) { return F::mul_small<26>(x); }
  __device__ __forceinline__ static F::T mad_a_lazy(const F::T& c, const F::T& x) { return F::mad_small_lazy(c, 26u, x); }
};
#ifndef SSO_MNT6_COOP_AFFINE
#define SSO_MNT6_COOP_AFFINE 1
#endif
struct Mnt6_753_G2C : Mnt6_753_G2 {
  // three lanes per point: the inversion tree of a block is 64 leaves of one coefficient per lane (37 KB instead of the 72 KB
  // that made the affine table a loss for the one-thread-per-element body), so mixed additions pay here
  static constexpr bool AFFINE_TABLE = SSO_MNT6_COOP_AFFINE != 0;
  static constexpr bool COOP_DEFAULT = true;           // 759 ms per chunk against 910 ms
  SSO_GROUP_COOP(mnt6_753_g2, Mnt6_753_G2, CFp3<Fq6, 11, false>)
  __device__ __forceinline__ static F::T mul_a(const F::T& x) { return F::mul_small<11>(F::mul_u2(x)); }
  // the Fq3 products tolerate arguments below 8 p only: a x stays canonical
  __device__ __forceinline__ static F::T mad_a_lazy(const F::T& c, const F::T& x) { return F::add_lazy(c, mul_a(x)); }
};
// the cooperative variant of a group, or void
template <class G> struct CoopOf { using type = void; };
template <> struct CoopOf<Bls12_377_G2> { using type = Bls12_377_G2C; };
template <> struct CoopOf<Mnt4_753_G2> { using type = Mnt4_753_G2C; };
template <> struct CoopOf<Mnt6_753_G2> { using type = Mnt6_753_G2C; };

// ids on the C ABI (include/sso_b200.h)
enum : uint32_t { CURVE_BLS12_377 = 0, CURVE_BW6_761 = 1, CURVE_MNT4_753 = 2, CURVE_MNT6_753 = 3 };
enum : uint32_t { GROUP_G1 = 0, GROUP_G2 = 1 };

// Dispatch a generic lambda on (curve, group): f(GroupCfg{})
template <class Fn> inline int dispatch_group(uint32_t curve, uint32_t group, Fn&& f) {
  switch (curve * 2 + group) {
    case 0: f(Bls12_377_G1{}); return 0;
    case 1: f(Bls12_377_G2{}); return 0;
    case 2: f(Bw6_761_G1{}); return 0;
    case 3: f(Bw6_761_G2{}); return 0;
    case 4: f(Mnt4_753_G1{}); return 0;
    case 5: f(Mnt4_753_G2{}); return 0;
    case 6: f(Mnt6_753_G1{}); return 0;
    case 7: f(Mnt6_753_G2{}); return 0;
  }
  return -1;
}

}  // namespace sso
