// Random-linear-combination multi-scalar multiplication for the same-ratio checks.
//
// B200-native counterpart of setup_utils::{merge_pairs, power_pairs} -> dense_multiexp /
// VariableBaseMSM: given two point vectors v1, v2 of equal length
// and fresh random scalars r_i, compute (sum r_i v1_i, sum r_i v2_i).  power_pairs(v) is the case
// v1 = v[..n-1], v2 = v[1..].  The reference draws r_i from thread_rng (not seeded), so only the
// verdict downstream is comparable — any unpredictable scalars do; ours come from ChaCha20 keyed with
// fresh host entropy.
//
// Pippenger's bucket method: scalars are cut into c-bit windows; (window, digit) keys are radix-sorted
// with CUB so every bucket owns a contiguous run of point indices; one thread accumulates one bucket
// (for both vectors at once — they share the scalars), then buckets are folded per window with a
// segmented running sum and the windows are combined by Horner's rule.
#pragma once
#include "curves.cuh"

namespace sso {

// ---- ChaCha20 block function (RFC 7539 quarter rounds), used as a counter-mode scalar generator ----
__device__ __forceinline__ uint32_t rotl32(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
#define SSO_QR(a, b, c, d)                          \
  a += b; d ^= a; d = rotl32(d, 16);                \
  c += d; b ^= c; b = rotl32(b, 12);                \
  a += b; d ^= a; d = rotl32(d, 8);                 \
  c += d; b ^= c; b = rotl32(b, 7);
__device__ __forceinline__ void chacha20_block(const uint32_t* key, uint64_t counter, uint32_t* out) {
  uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3],
                    key[4], key[5], key[6], key[7], (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
  uint32_t x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = s[i];
#pragma unroll 1
  for (int r = 0; r < 10; r++) {
    SSO_QR(x[0], x[4], x[8], x[12]) SSO_QR(x[1], x[5], x[9], x[13]) SSO_QR(x[2], x[6], x[10], x[14]) SSO_QR(x[3], x[7], x[11], x[15])
    SSO_QR(x[0], x[5], x[10], x[15]) SSO_QR(x[1], x[6], x[11], x[12]) SSO_QR(x[2], x[7], x[8], x[13]) SSO_QR(x[3], x[4], x[9], x[14])
  }
#pragma unroll
  for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

// Width of the random-linear-combination scalars.  The reference draws full-size Fr::rand scalars from thread_rng; what
// the same-ratio test needs from them is unpredictability — a contribution with a wrong element passes with probability
// 2^-120 when the r_i are uniform 120-bit integers (batch verification commonly works with 2^-128 or less) — so this core draws
// 120-bit scalars: half (BLS12-377), a third (BW6-761) or a sixth (MNT4/6-753) of the windows, bucket additions and
// final doublings of an MSM with full-size scalars.  120 rather than 128 because it is divisible by every window width
// the MSM uses (4, 5, 6, 8, 10, 12): no window is left partial — a top window holding only a few bits has few, very long
// buckets, and one thread sums one bucket (measured: 2x on the bucket phase with 128 bits and 10-bit windows).
// Verdicts, not scalars, are what is comparable with the reference.
static constexpr int RLC_BITS = 120;
static constexpr int RLC_WORDS = 4;

// scalar i = first SBITS bits of the ChaCha20 keystream blocks (2i, 2i+1): uniform in [0, 2^SBITS)
// with SBITS = RLC_BITS < bits(r).  Words: [i][KL].
template <int KL, int SBITS>
__device__ __forceinline__ void body_random_scalars(uint32_t tid, uint32_t n, const uint32_t* key, uint32_t* scalars) {
  if (tid >= n) return;
  uint32_t blk[32];
  chacha20_block(key, 2ull * tid, blk);
  if (KL > 16) chacha20_block(key, 2ull * tid + 1, blk + 16);
#pragma unroll
  for (int i = 0; i < KL; i++) {
    uint32_t w = blk[i];
    int lo = 32 * i;
    if (lo + 32 > SBITS) w = lo >= SBITS ? 0u : (w & ((1u << (SBITS - lo)) - 1u));
    scalars[(size_t)tid * KL + i] = w;
  }
}

// c-bit window `w` of a little-endian scalar of KL words
template <int KL>
__device__ __forceinline__ uint32_t window_digit(const uint32_t* k, uint32_t w, uint32_t c) {
  uint32_t bit = w * c, word = bit >> 5, off = bit & 31;
  if (word >= (uint32_t)KL) return 0;
  uint64_t v = k[word];
  if (word + 1 < (uint32_t)KL) v |= (uint64_t)k[word + 1] << 32;
  return (uint32_t)(v >> off) & ((1u << c) - 1u);
}

// keys[w * n + i] = (w << c) | digit_w(scalar_i), vals[...] = i     (digit 0 is kept: its bucket is skipped later)
template <int KL>
__device__ __forceinline__ void body_msm_keys(uint32_t tid, uint32_t n, uint32_t nwin, uint32_t c, const uint32_t* scalars,
                                              uint32_t* keys, uint32_t* vals) {
  if (tid >= n) return;
  uint32_t k[KL];
#pragma unroll
  for (int i = 0; i < KL; i++) k[i] = scalars[(size_t)tid * KL + i];
  for (uint32_t w = 0; w < nwin; w++) {
    keys[(size_t)w * n + tid] = (w << c) | window_digit<KL>(k, w, c);
    vals[(size_t)w * n + tid] = tid;
  }
}

__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* a, uint32_t n, uint32_t key) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// affine point [x|y] words, all-zero = infinity
template <class G>
__device__ __forceinline__ typename SW<G>::Affine load_affine(const uint32_t* aff, size_t idx) {
  using F = typename G::F;
  typename SW<G>::Affine p;
  const uint32_t* a = aff + idx * 2 * F::WORDS;
  p.x = F::load(a, 1);
  p.y = F::load(a + F::WORDS, 1);
  p.inf = F::is_zero(p.x) && F::is_zero(p.y);
  return p;
}
template <class G>
__device__ __forceinline__ void store_jac(uint32_t* dst, const typename SW<G>::Jac& p) {
  using F = typename G::F;
  F::store(dst, 1, p.X);
  F::store(dst + F::WORDS, 1, p.Y);
  F::store(dst + 2 * F::WORDS, 1, p.Z);
}
template <class G>
__device__ __forceinline__ typename SW<G>::Jac load_jac(const uint32_t* src) {
  using F = typename G::F;
  return typename SW<G>::Jac{F::load(src, 1), F::load(src + F::WORDS, 1), F::load(src + 2 * F::WORDS, 1)};
}

// One thread per bucket (window w, digit d >= 1): sums its run of points, for both vectors.
// buckets: [vector][w * nb + d][3 * WORDS]
template <class G>
__device__ __forceinline__ void body_msm_buckets(uint32_t tid, uint32_t n, uint32_t nwin, uint32_t c, const uint32_t* sorted_keys,
                                                 const uint32_t* sorted_vals, const uint32_t* aff_a, const uint32_t* aff_b,
                                                 uint32_t* buckets) {
  using C = SW<G>;
  using F = typename G::F;
  uint32_t nb = 1u << c;
  if (tid >= nwin * nb) return;
  size_t total = (size_t)n * nwin;
  typename C::Jac sa = C::identity(), sb = C::identity();
  if ((tid & (nb - 1)) != 0) {
    uint32_t lo = lower_bound_u32(sorted_keys, (uint32_t)total, tid);
    for (uint32_t i = lo; i < total && sorted_keys[i] == tid; i++) {
      uint32_t idx = sorted_vals[i];
      sa = C::madd(sa, load_affine<G>(aff_a, idx));
      sb = C::madd(sb, load_affine<G>(aff_b, idx));
    }
  }
  store_jac<G>(buckets + (size_t)tid * 3 * F::WORDS, sa);
  store_jac<G>(buckets + ((size_t)nwin * nb + tid) * 3 * F::WORDS, sb);
}

// Segmented bucket fold.  Thread (vector v, window w, segment s) covers digits [s*SEG, (s+1)*SEG) and leaves
//   S = sum B_d,  T = sum (d - s*SEG) B_d   in seg_out[(v*nwin + w)*nseg + s][2][3*WORDS]
static constexpr uint32_t MSM_SEG = 64;
template <class G>
__device__ __forceinline__ void body_msm_fold(uint32_t tid, uint32_t nwin, uint32_t c, const uint32_t* buckets, uint32_t* seg_out) {
  using C = SW<G>;
  using F = typename G::F;
  uint32_t nb = 1u << c;
  uint32_t seg = nb < MSM_SEG ? nb : MSM_SEG;
  uint32_t nseg = nb / seg;
  if (tid >= 2 * nwin * nseg) return;
  uint32_t s = tid % nseg, vw = tid / nseg;                 // vw = v * nwin + w
  const uint32_t* base = buckets + ((size_t)vw * nb + (size_t)s * seg) * 3 * F::WORDS;
  typename C::Jac run = C::identity(), acc = C::identity();
  for (int d = (int)seg - 1; d >= 1; d--) {                 // running sum: acc = sum_{d>=1} d * B_d
    run = C::add(run, load_jac<G>(base + (size_t)d * 3 * F::WORDS));
    acc = C::add(acc, run);
  }
  run = C::add(run, load_jac<G>(base));                     // digit offset 0 contributes to S only
  store_jac<G>(seg_out + (size_t)tid * 6 * F::WORDS, run);
  store_jac<G>(seg_out + (size_t)tid * 6 * F::WORDS + 3 * F::WORDS, acc);
}

// One thread per (vector, window): window sum = sum_s (T_s + s*SEG * S_s)
template <class G>
__device__ __forceinline__ void body_msm_window(uint32_t tid, uint32_t nwin, uint32_t c, const uint32_t* seg_in, uint32_t* win_out) {
  using C = SW<G>;
  using F = typename G::F;
  uint32_t nb = 1u << c;
  uint32_t seg = nb < MSM_SEG ? nb : MSM_SEG;
  uint32_t nseg = nb / seg;
  if (tid >= 2 * nwin) return;
  // window sum = sum_s T_s + seg * sum_{s >= 1} s * S_s : the second term by the running-sum trick over the segments
  // (two additions per segment) and log2(seg) doublings at the end, instead of one 16-bit scalar multiplication per segment
  typename C::Jac tsum = C::identity(), run = C::identity(), acc = C::identity();
  for (int s = (int)nseg - 1; s >= 0; s--) {
    const uint32_t* p = seg_in + ((size_t)tid * nseg + s) * 6 * F::WORDS;
    tsum = C::add(tsum, load_jac<G>(p + 3 * F::WORDS));
    if (s >= 1) {
      run = C::add(run, load_jac<G>(p));
      acc = C::add(acc, run);
    }
  }
  for (uint32_t b = 1; b < seg; b <<= 1) acc = C::dbl(acc);
  typename C::Jac total = C::add(tsum, acc);
  store_jac<G>(win_out + (size_t)tid * 3 * F::WORDS, total);
}

// One thread per vector: Horner over the windows, normalise, write the point uncompressed.
template <class G>
__device__ __forceinline__ void body_msm_final(uint32_t tid, uint32_t nwin, uint32_t c, const uint32_t* win_in, uint8_t* out) {
  using C = SW<G>;
  using F = typename G::F;
  if (tid >= 2) return;
  typename C::Jac acc = C::identity();
  for (int w = (int)nwin - 1; w >= 0; w--) {
    for (uint32_t i = 0; i < c; i++) acc = C::dbl(acc);
    acc = C::add(acc, load_jac<G>(win_in + ((size_t)tid * nwin + w) * 3 * F::WORDS));
  }
  typename C::Affine a;
  if (C::is_identity(acc)) { a.inf = true; a.x = F::zero(); a.y = F::zero(); }
  else a = C::to_affine_with(acc, F::inv(acc.Z));
  C::write_uncompressed(out + (size_t)tid * C::SIZE_U, a);
}

}  // namespace sso
