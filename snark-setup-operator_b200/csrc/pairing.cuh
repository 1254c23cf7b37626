// same_ratio on the device: reduced Tate pairing, one warp per pairing.
//
// B200-native counterpart of setup_utils::same_ratio / check_same_ratio (SURVEY.md §2.1 K8, §8a
// row a6): e(a, d) == e(b, c) for a, b in G1 and c, d in G2.  arkworks evaluates optimal ate pairings
// with per-family towers; the verdict only needs SOME non-degenerate bilinear pairing, so this core
// uses the reduced Tate pairing t(P, Q) = f_{r,P}(psi(Q))^((q^k-1)/r) over the binomial ring
// Fq[w]/(w^k - nu) for all four curves (k = 12, 6, 4, 6) — the same construction as oracle/pairing.py.
//
// Parallelisation: an Fq^k product is k independent coefficient sums, so lane l of the warp owns
// coefficient l (k Fq multiplications per lane instead of k^2 per thread); the Miller-loop point T lives
// redundantly in every lane's registers; operands live in shared memory.  O(1) pairings per chunk:
// a handful of warps, ~tens of milliseconds — not on the roofline path.
#pragma once
#include "curves.cuh"

namespace sso {

template <class F> __device__ __forceinline__ typename F::Base::T ext_coeff(const typename F::T& x, int j) {
  if constexpr (F::DEG == 1) { return x; }
  else if constexpr (F::DEG == 2) { return j == 0 ? x.c0 : x.c1; }
  else { return j == 0 ? x.c0 : (j == 1 ? x.c1 : x.c2); }
}

template <class G1, class G2, class PP> struct Pairing {
  using Fq = typename G1::F;
  using FT = typename Fq::T;
  using C1 = SW<G1>;
  using C2 = SW<G2>;
  static constexpr int K = PP::K;
  struct Ws { FT f[K], g[K], xq[K], yq[K], ln[K], u[K], v[K], part[3 * K]; };

  __device__ __forceinline__ static FT mul_nu(const FT& x) {
    FT t = Fq::template mul_small<PP::NU_ABS>(x);
    return PP::NU_NEG ? Fq::neg(t) : t;
  }

  // two-phase lane-parallel update: every coefficient is computed from the OLD contents, then stored
  template <class Fn> __device__ __forceinline__ static void coop(int lane, FT* out, Fn fn) {
#ifdef SSO_HOST_EMUL
    (void)lane;
    FT tmp[K];
    for (int l = 0; l < K; l++) tmp[l] = fn(l);
    for (int l = 0; l < K; l++) out[l] = tmp[l];
#else
    FT val = Fq::zero();
    if (lane < K) val = fn(lane);
    __syncwarp();
    if (lane < K) out[lane] = val;
    __syncwarp();
#endif
  }

  // out = a * b in Fq[w]/(w^K - nu).  The K^2 limb products are spread over all lanes: lane l = part * K + c
  // (part < PARTS = 32 / K) sums the products a[i] * b[c - i] for i = part (mod PARTS); the partial sums meet in
  // ws_part (shared memory) and lane c (part 0) adds them.  K = 12: 6 multiplications per lane instead of 12.
  static constexpr int PARTS = (32 / K) < 1 ? 1 : (32 / K > 4 ? 4 : 32 / K);
  __device__ __noinline__ static void kmul(int lane, FT* out, const FT* a, const FT* b, FT* part_buf) {
#ifdef SSO_HOST_EMUL
    (void)lane; (void)part_buf;
    FT tmp[K];
    for (int l = 0; l < K; l++) {
      FT lo = Fq::zero(), hi = Fq::zero();
      for (int i = 0; i < K; i++) {
        int j = l - i;
        bool wrap = j < 0;
        if (wrap) j += K;
        FT pr = Fq::mul(a[i], b[j]);
        if (wrap) hi = Fq::add(hi, pr); else lo = Fq::add(lo, pr);
      }
      tmp[l] = Fq::add(lo, mul_nu(hi));
    }
    for (int l = 0; l < K; l++) out[l] = tmp[l];
#else
    int part = lane / K, c = lane - part * K;
    FT val = Fq::zero();
    if (part < PARTS) {
      FT lo = Fq::zero(), hi = Fq::zero();
      for (int i = part; i < K; i += PARTS) {
        int j = c - i;
        bool wrap = j < 0;
        if (wrap) j += K;
        FT pr = Fq::mul(a[i], b[j]);
        if (wrap) hi = Fq::add(hi, pr); else lo = Fq::add(lo, pr);
      }
      val = Fq::add(lo, mul_nu(hi));
      if (part > 0) part_buf[(part - 1) * K + c] = val;
    }
    __syncwarp();
    if (part == 0) {
      for (int p = 1; p < PARTS; p++) val = Fq::add(val, part_buf[(p - 1) * K + c]);
    }
    __syncwarp();
    if (part == 0) out[c] = val;
    __syncwarp();
#endif
  }
  __device__ __forceinline__ static void kmul(int lane, Ws& ws, FT* out, const FT* a, const FT* b) { kmul(lane, out, a, b, ws.part); }
  // conjugation at `stride`: negate coefficients at odd multiples of stride
  __device__ __forceinline__ static void kconj(int lane, FT* out, const FT* a, int stride) {
    coop(lane, out, [&](int l) { return ((l / stride) & 1) ? Fq::neg(a[l]) : a[l]; });
  }
  __device__ __forceinline__ static void kcopy(int lane, FT* out, const FT* a) {
    coop(lane, out, [&](int l) { return a[l]; });
  }
  __device__ __forceinline__ static void kone(int lane, FT* out) {
    coop(lane, out, [&](int l) { return l == 0 ? Fq::one() : Fq::zero(); });
  }

  // out = a^-1.  a^-1 = conj(a) * (a conj(a))^-1 descends through the even-degree subfields until the
  // subfield has degree 1 (Fq inverse) or 3 (cubic formula).  Uses ws.u, ws.v, ws.ln as scratch.
  __device__ __noinline__ static void kinv(int lane, Ws& ws, FT* out, const FT* a) {
    FT* cur = ws.u;
    FT* acc = ws.v;
    FT* tmp = ws.ln;
    kcopy(lane, cur, a);
    kone(lane, acc);
    int stride = 1;
    while (((K / stride) & 1) == 0) {
      kconj(lane, tmp, cur, stride);
      kmul(lane, ws, acc, acc, tmp);
      kmul(lane, ws, cur, cur, tmp);
      stride *= 2;
    }
    int n = K / stride;
    if (n == 1) {
      FT inv0 = Fq::inv(cur[0]);
      coop(lane, cur, [&](int l) { return l == 0 ? inv0 : Fq::zero(); });
    } else {                                         // n == 3: z = w^stride, z^3 = nu
      FT a0 = cur[0], a1 = cur[stride], a2 = cur[2 * stride];
      FT t0 = Fq::sub(Fq::sqr(a0), mul_nu(Fq::mul(a1, a2)));
      FT t1 = Fq::sub(mul_nu(Fq::sqr(a2)), Fq::mul(a0, a1));
      FT t2 = Fq::sub(Fq::sqr(a1), Fq::mul(a0, a2));
      FT nn = Fq::add(Fq::mul(a0, t0), mul_nu(Fq::add(Fq::mul(a2, t1), Fq::mul(a1, t2))));
      FT ni = Fq::inv(nn);
      coop(lane, cur, [&](int l) {
        if (l == 0) return Fq::mul(t0, ni);
        if (l == stride) return Fq::mul(t1, ni);
        if (l == 2 * stride) return Fq::mul(t2, ni);
        return Fq::zero();
      });
    }
    kmul(lane, ws, out, acc, cur);
  }

  // psi(Q): xq = embed(x') w^(2s), yq = embed(y') w^(3s)
  __device__ __forceinline__ static void untwist_coord(int lane, FT* out, const typename G2::F::T& v, int shift) {
    constexpr int DEG = PP::G2_DEG;
    constexpr int STEP = DEG > 1 ? K / DEG : 0;
    FT nuinv = Fq::from_const(PP::nuinv());
    coop(lane, out, [&](int l) {
      FT r = Fq::zero();
      for (int j = 0; j < DEG; j++) {
        int pos = j * STEP + shift;
        FT c = ext_coeff<typename G2::F>(v, j);
        if (pos < 0) { pos += K; c = Fq::mul(c, nuinv); }
        else if (pos >= K) { pos -= K; c = mul_nu(c); }
        if (pos == l) r = c;
      }
      return r;
    });
  }

  // ws.f = f_{r,P}(psi(Q)) (Miller function, vertical lines dropped); P, Q affine (Montgomery); identity inputs give 1
  __device__ __noinline__ static void miller(int lane, Ws& ws, const typename C1::Affine& P, const typename C2::Affine& Q) {
    kone(lane, ws.f);
    if (P.inf || Q.inf) return;
    untwist_coord(lane, ws.xq, Q.x, 2 * PP::TWIST_SIGN);
    untwist_coord(lane, ws.yq, Q.y, 3 * PP::TWIST_SIGN);
    typename C1::Jac T{P.x, P.y, Fq::one()};
    const uint32_t* r = G1::order();
    constexpr int RBITS = G1::Fr::P::BITS;
    for (int bit = RBITS - 2; bit >= 0; bit--) {
      // tangent at T, scaled by Fq factors: A yq + B xq + C
      {
        FT ZZ = Fq::sqr(T.Z);
        FT t = Fq::add(Fq::dbl(Fq::sqr(T.X)), Fq::sqr(T.X));
        if (!G1::A_IS_ZERO) t = Fq::add(t, G1::mul_a(Fq::sqr(ZZ)));
        FT A = Fq::dbl(Fq::mul(Fq::mul(T.Y, T.Z), ZZ));
        FT B = Fq::neg(Fq::mul(t, ZZ));
        FT Cc = Fq::sub(Fq::mul(t, T.X), Fq::dbl(Fq::sqr(T.Y)));
        coop(lane, ws.ln, [&](int l) {
          FT v = Fq::add(Fq::mul(A, ws.yq[l]), Fq::mul(B, ws.xq[l]));
          return l == 0 ? Fq::add(v, Cc) : v;
        });
      }
      kmul(lane, ws, ws.f, ws.f, ws.f);
      kmul(lane, ws, ws.f, ws.f, ws.ln);
      T = C1::dbl(T);
      if ((r[bit >> 5] >> (bit & 31)) & 1) {
        FT ZZ = Fq::sqr(T.Z);
        FT ZZZ = Fq::mul(ZZ, T.Z);
        FT N = Fq::sub(Fq::mul(T.X, T.Z), Fq::mul(P.x, ZZZ));
        FT M = Fq::sub(T.Y, Fq::mul(P.y, ZZZ));
        if (!Fq::is_zero(N) && !C1::is_identity(T)) {
          FT B = Fq::neg(M);
          FT Cc = Fq::sub(Fq::mul(M, P.x), Fq::mul(P.y, N));
          coop(lane, ws.ln, [&](int l) {
            FT v = Fq::add(Fq::mul(N, ws.yq[l]), Fq::mul(B, ws.xq[l]));
            return l == 0 ? Fq::add(v, Cc) : v;
          });
          kmul(lane, ws, ws.f, ws.f, ws.ln);
        }
        T = C1::madd(T, P);
      }
    }
  }

  // Frobenius on Fq[w]/(w^K - nu): coefficient l is multiplied by c^l, c = nu^((q-1)/K) (PP::frob(), BLS12 only)
  __device__ __forceinline__ static void kfrob(int lane, FT* out, const FT* a) {
    coop(lane, out, [&](int l) { return Fq::mul(a[l], Fq::from_const(PP::frob() + (size_t)l * Fq::L)); });
  }
  // out = a^x for the 64-bit curve parameter x (square-and-multiply, x sparse); out must not alias a; tmp is scratch
  __device__ __noinline__ static void kexp_x(int lane, Ws& ws, FT* out, const FT* a) {
    constexpr unsigned long long X = PP::X;
    kcopy(lane, out, a);
    bool started = false;
    for (int i = 63; i >= 0; i--) {
      bool bit = ((X >> i) & 1ull) != 0;
      if (!started) { started = bit; continue; }
      kmul(lane, ws, out, out, out);
      if (bit) kmul(lane, ws, out, out, a);
    }
  }
  // BLS12: is f^((q^12 - 1) / r) == 1 ?  Easy part g = f^((q^6 - 1)(q^2 + 1)); then, with
  //   3 (q^4 - q^2 + 1) / r = (x - 1)^2 (x + q) (x^2 + q^2 - 1) + 3      (asserted in tools/gen_constants.py)
  // and 3 not dividing r, the test is g^((x-1)^2 (x+q) (x^2+q^2-1)) g^3 == 1: five exponentiations by the 64-bit x (about 350
  // Fq12 multiplications) instead of square-and-multiply over the 2009-bit (q^6 + 1) / r (about 3000).  After the easy part g
  // is unitary: its inverse is the conjugate.  Leaves the value in ws.f.
  __device__ __noinline__ static void final_exp_bls12(int lane, Ws& ws) {
    kinv(lane, ws, ws.g, ws.f);                           // g = f^-1
    kconj(lane, ws.u, ws.f, 1);                           // u = f^(q^6)
    kmul(lane, ws, ws.g, ws.g, ws.u);                     // g = f^(q^6 - 1)
    kfrob(lane, ws.u, ws.g);
    kfrob(lane, ws.u, ws.u);                              // g^(q^2)
    kmul(lane, ws, ws.g, ws.g, ws.u);                     // g = f^((q^6 - 1)(q^2 + 1))
    FT* t0 = ws.xq; FT* t1 = ws.yq; FT* t2 = ws.ln;
    kexp_x(lane, ws, t0, ws.g);
    kconj(lane, ws.u, ws.g, 1);
    kmul(lane, ws, t0, t0, ws.u);                         // t0 = g^(x - 1)
    kexp_x(lane, ws, t1, t0);
    kconj(lane, ws.u, t0, 1);
    kmul(lane, ws, t1, t1, ws.u);                         // t1 = g^((x - 1)^2)
    kexp_x(lane, ws, t2, t1);
    kfrob(lane, ws.u, t1);
    kmul(lane, ws, t2, t2, ws.u);                         // t2 = t1^(x + q)
    kexp_x(lane, ws, t0, t2);
    kexp_x(lane, ws, t1, t0);                             // t1 = t2^(x^2)
    kfrob(lane, ws.u, t2);
    kfrob(lane, ws.u, ws.u);                              // t2^(q^2)
    kmul(lane, ws, t1, t1, ws.u);
    kconj(lane, ws.u, t2, 1);
    kmul(lane, ws, t1, t1, ws.u);                         // t1 = t2^(x^2 + q^2 - 1)
    kmul(lane, ws, ws.f, ws.g, ws.g);
    kmul(lane, ws, ws.f, ws.f, ws.g);                     // g^3
    kmul(lane, ws, ws.f, ws.f, t1);
  }

  // ws.f <- ws.f ^ ((q^k - 1) / r): (q^(k/2) - 1) by conjugation and one inversion, then (q^(k/2) + 1) / r.
  // (BLS12: a value that is 1 exactly when that power is 1 — final_exp_bls12.)
  __device__ __noinline__ static void final_exp(int lane, Ws& ws) {
    if constexpr (PP::BLS12_FINAL_EXP) { final_exp_bls12(lane, ws); return; }
    kinv(lane, ws, ws.g, ws.f);                       // g = f^-1
    kconj(lane, ws.u, ws.f, 1);                       // u = f^(q^(k/2))
    kmul(lane, ws, ws.g, ws.g, ws.u);                     // g = f^(q^(k/2) - 1)
    kone(lane, ws.f);
    const uint32_t* e = PP::hard();
    bool started = false;
    for (int i = PP::HARD_WORDS * 32 - 1; i >= 0; i--) {
      if (started) kmul(lane, ws, ws.f, ws.f, ws.f);
      if ((e[i >> 5] >> (i & 31)) & 1) {
        if (started) kmul(lane, ws, ws.f, ws.f, ws.g); else kcopy(lane, ws.f, ws.g);
        started = true;
      }
    }
  }

  // reduced Tate pairing t(P, Q)
  __device__ __forceinline__ static void tate(int lane, Ws& ws, const typename C1::Affine& P, const typename C2::Affine& Q) {
    miller(lane, ws, P, Q);
    final_exp(lane, ws);
  }

  // bytes of one check: a | b (G1 uncompressed) | c | d (G2 uncompressed)
  static constexpr int CHECK_BYTES = 2 * C1::SIZE_U + 2 * C2::SIZE_U;

  // same_ratio as ONE final exponentiation: e(a, d) == e(b, c)  <=>  (f_a(d) * f_{-b}(c)) ^ ((q^k-1)/r) == 1.
  // side 0 leaves the Miller value f_{r,a}(d) in ws.f, side 1 leaves f_{r,-b}(c).  Returns a deserialisation
  // status (0 ok).
  __device__ __forceinline__ static uint32_t run_side(int lane, Ws& ws, const uint8_t* check, int side) {
    typename C1::Affine P;
    typename C2::Affine Q;
    uint32_t s1 = C1::read_uncompressed(check + (side == 0 ? 0 : C1::SIZE_U), P);
    uint32_t s2 = C2::read_uncompressed(check + 2 * C1::SIZE_U + (side == 0 ? C2::SIZE_U : 0), Q);
    if (s1 != 0 || s2 != 0) return s1 ? s1 : s2;
    if (!C1::on_curve(P) || !C2::on_curve(Q)) return 3;
    if (side == 1 && !P.inf) P.y = Fq::neg(P.y);
    miller(lane, ws, P, Q);
    return 0;
  }
  // after both sides: ws0.f <- (ws0.f * ws1.f)^((q^k-1)/r); true iff the result is 1
  __device__ __forceinline__ static bool combine_and_check(int lane, Ws& ws0, const Ws& ws1) {
    kmul(lane, ws0, ws0.f, ws0.f, ws1.f);
    final_exp(lane, ws0);
    bool ok = true;
#ifdef SSO_HOST_EMUL
    for (int l = 0; l < K; l++) ok = ok && Fq::eq(ws0.f[l], l == 0 ? Fq::one() : Fq::zero());
#else
    if (lane < K) ok = Fq::eq(ws0.f[lane], lane == 0 ? Fq::one() : Fq::zero());
    ok = __all_sync(0xffffffffu, ok);
#endif
    return ok;
  }
};

}  // namespace sso
