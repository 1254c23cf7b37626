// Key generation and hash-to-G2 on the device (single-thread kernels; O(1) per chunk).
//
// B200-native counterpart of Phase1::key_generation, setup_utils::compute_g2_s / hash_to_g2 and the
// arkworks samplers they use (Appendix A.3, A.4) [UP]:
//   derive_rng_from_seed      ChaCha20 (rand_chacha 0.3.1), 64-bit block counter from 0, words little-endian
//   Fp::rand                  ceil(bits/64) u64 limbs, top bits shaved, limbs ARE the Montgomery residue,
//                             rejected if >= p
//   Projective::rand          loop { x <- F::rand; greatest <- bool; y = sqrt(x^3 + ax + b) } then * cofactor
// The 32-bit Montgomery limbs of this core coincide with ark-ff's 64-bit ones (same radix), so the
// sampled words can be used as residues directly.
#pragma once
#include "msm.cuh"

namespace sso {

struct ChaChaStream {
  uint32_t key[8];
  uint64_t counter;
  uint32_t buf[16];
  int idx;
  __device__ __forceinline__ void init(const uint32_t* k) {
#pragma unroll
    for (int i = 0; i < 8; i++) key[i] = k[i];
    counter = 0; idx = 16;
  }
  __device__ __forceinline__ uint32_t next_u32() {
    if (idx >= 16) { chacha20_block(key, counter, buf); counter++; idx = 0; }
    return buf[idx++];
  }
};

// ark-ff Fp::rand
template <class Fp_> __device__ __forceinline__ typename Fp_::T fp_rand(ChaChaStream& rng) {
  using P = typename Fp_::P;
  constexpr int L = P::L;
  constexpr int SHAVE = 32 * L - P::BITS;          // < 32 for all five primes
  typename Fp_::T v, t;
  for (;;) {
    for (int i = 0; i < L; i++) v.v[i] = rng.next_u32();
    if (SHAVE > 0) v.v[L - 1] &= (0xffffffffu >> SHAVE);
    if (limbs_sub<L>(t.v, v.v, P::p()) != 0) return v;   // borrow: v < p
  }
}
template <class F> __device__ __forceinline__ typename F::T field_rand(ChaChaStream& rng) {
  if constexpr (F::DEG == 1) { return fp_rand<F>(rng); }
  else if constexpr (F::DEG == 2) { typename F::T r; r.c0 = fp_rand<typename F::Base>(rng); r.c1 = fp_rand<typename F::Base>(rng); return r; }
  else { typename F::T r; r.c0 = fp_rand<typename F::Base>(rng); r.c1 = fp_rand<typename F::Base>(rng); r.c2 = fp_rand<typename F::Base>(rng); return r; }
}

// ark-ec Projective::rand, first half: the curve point before the cofactor is cleared (the acceptance test — is
// x^3 + ax + b a square — decides how many words the RNG hands out, so this part is sequential in the RNG)
template <class G> __device__ __noinline__ typename SW<G>::Affine group_rand_point(ChaChaStream& rng) {
  using C = SW<G>;
  using F = typename G::F;
  for (;;) {
    typename F::T x = field_rand<F>(rng);
    bool greatest = (rng.next_u32() >> 31) != 0;
    typename F::T y;
    if (!G::field_sqrt(C::rhs(x), y)) continue;
    if (F::lex_is_neg(y) != greatest) y = F::neg(y);
    return typename C::Affine{x, y, false};
  }
}
template <class G> __device__ __forceinline__ typename SW<G>::Jac group_rand(ChaChaStream& rng) {
  typename SW<G>::Affine p = group_rand_point<G>(rng);
  return SW<G>::mul_const(p, G::cofactor(), G::COFACTOR_WORDS);
}

template <class G> __device__ __forceinline__ typename SW<G>::Affine jac_to_affine(const typename SW<G>::Jac& j) {
  using C = SW<G>;
  using F = typename G::F;
  typename C::Affine a;
  if (C::is_identity(j)) { a.inf = true; a.x = F::zero(); a.y = F::zero(); }
  else a = C::to_affine_with(j, F::inv(j.Z));
  return a;
}

// Key generation, G1 half: tau, alpha, beta <- Fr::rand; for each scalar x: g1_s <- G1::rand, g1_s_x = x g1_s.
//   scalars_out : 3 canonical scalars, Fr::L words each
//   g1_out      : g1_s(tau) | g1_s_x(tau) | g1_s(alpha) | ... uncompressed (public-key order)
// nscalars = 3 for phase 1 (tau, alpha, beta), 1 for phase 2 (delta).
// Two steps so that the kernel can hand the long group operations of the three proofs to three warps:
//   sample (one thread, sequential in the RNG): the scalars and the three curve points before cofactor clearing
//   finish (independent per scalar)           : cofactor clearing, multiplication by the scalar, serialisation
template <class G1>
__device__ __forceinline__ void keygen_g1_sample(const uint32_t* seed, uint32_t nscalars, uint32_t* scalars_out,
                                                 typename SW<G1>::Affine* raw) {
  using Fr = typename G1::Fr;
  ChaChaStream rng;
  rng.init(seed);
  for (uint32_t i = 0; i < nscalars; i++) {
    typename Fr::T c = Fr::from_mont(fp_rand<Fr>(rng));
    for (int w = 0; w < Fr::L; w++) scalars_out[i * Fr::L + w] = c.v[w];
  }
  for (uint32_t i = 0; i < nscalars; i++) raw[i] = group_rand_point<G1>(rng);
}
template <class G1>
__device__ __forceinline__ void keygen_g1_finish(uint32_t i, const typename SW<G1>::Affine& raw, const uint32_t* scalars, uint8_t* g1_out) {
  using C = SW<G1>;
  using Fr = typename G1::Fr;
  typename C::Affine s = jac_to_affine<G1>(C::mul_const(raw, G1::cofactor(), G1::COFACTOR_WORDS));
  uint32_t k[Fr::L];
  for (int w = 0; w < Fr::L; w++) k[w] = scalars[i * Fr::L + w];
  typename C::Affine sx = jac_to_affine<G1>(C::template scalar_mul<Fr::L, Fr::P::BITS>(s, k));
  C::write_uncompressed(g1_out + (size_t)(2 * i) * C::SIZE_U, s);
  C::write_uncompressed(g1_out + (size_t)(2 * i + 1) * C::SIZE_U, sx);
}
// single-thread form (emulation harness)
template <class G1>
__device__ __forceinline__ void body_keygen_g1(const uint32_t* seed, uint32_t nscalars, uint32_t* scalars_out, uint8_t* g1_out) {
  typename SW<G1>::Affine raw[3];
  keygen_g1_sample<G1>(seed, nscalars, scalars_out, raw);
  for (uint32_t i = 0; i < nscalars; i++) keygen_g1_finish<G1>(i, raw[i], scalars_out, g1_out);
}

// Thread i: g2_s = G2::rand(ChaCha20(seeds[i])); optionally g2_s_x = scalars[i] * g2_s.
//   g2_s_out / g2_sx_out : uncompressed points (g2_sx_out may be null)
template <class G2>
__device__ __forceinline__ void body_hash_to_g2(uint32_t tid, uint32_t n, const uint32_t* seeds, const uint32_t* scalars,
                                                uint8_t* g2_s_out, uint8_t* g2_sx_out) {
  using C = SW<G2>;
  using Fr = typename G2::Fr;
  if (tid >= n) return;
  ChaChaStream rng;
  rng.init(seeds + 8 * tid);
  typename C::Affine s = jac_to_affine<G2>(group_rand<G2>(rng));
  C::write_uncompressed(g2_s_out + (size_t)tid * C::SIZE_U, s);
  if (g2_sx_out) {
    uint32_t k[Fr::L];
    for (int w = 0; w < Fr::L; w++) k[w] = scalars[tid * Fr::L + w];
    typename C::Affine sx = jac_to_affine<G2>(C::template scalar_mul<Fr::L, Fr::P::BITS>(s, k));
    C::write_uncompressed(g2_sx_out + (size_t)tid * C::SIZE_U, sx);
  }
}

}  // namespace sso
