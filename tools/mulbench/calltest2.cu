#include "ext.cuh"
#include "curves.cuh"
using namespace sso;
using F2 = Fq377x2;
__device__ __noinline__ F2::T mul2_byval(F2::T a, F2::T b) {
  using B = Fq377;
  B::T v0 = B::mul(a.c0, b.c0), v1 = B::mul(a.c1, b.c1), s = B::mul(B::add(a.c0, a.c1), B::add(b.c0, b.c1));
  F2::T r; r.c0 = B::sub(v0, B::mul_small<5>(v1)); r.c1 = B::sub(B::sub(s, v0), v1); return r;
}
__global__ void k1(uint32_t* out, uint32_t n) {
  F2::T x, y;
  for (int i = 0; i < 12; i++) { x.c0.v[i] = out[i] + threadIdx.x; x.c1.v[i] = out[i+5]; y.c0.v[i] = out[12 + i]; y.c1.v[i] = out[30+i]; }
  for (uint32_t i = 0; i < n; i++) { x = mul2_byval(x, y); y = mul2_byval(y, x); }
  for (int i = 0; i < 12; i++) out[threadIdx.x * 12 + i] = x.c0.v[i] ^ y.c1.v[i] ^ x.c1.v[i] ^ y.c0.v[i];
}
using F24 = Fp<P_q4>;
__global__ void k3(uint32_t* out, uint32_t n) {
  F24::T x, y;
  for (int i = 0; i < 24; i++) { x.v[i] = out[i] + threadIdx.x; y.v[i] = out[24 + i]; }
  for (uint32_t i = 0; i < n; i++) { x = F24::mul(x, y); y = F24::mul(y, x); }
  for (int i = 0; i < 24; i++) out[threadIdx.x * 24 + i] = x.v[i] ^ y.v[i];
}
