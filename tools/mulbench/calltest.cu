#include "fp.cuh"
using namespace sso;
using F = Fp<P_q377>;
struct V { uint32_t v[12]; };
__device__ __noinline__ V mul_byval(V a, V b) { V r; mont_mul<P_q377>(r.v, a.v, b.v); return r; }
__device__ __noinline__ void mul_byref(V& r, const V& a, const V& b) { mont_mul<P_q377>(r.v, a.v, b.v); }
__global__ void k1(uint32_t* out, uint32_t n) {
  V x, y;
  for (int i = 0; i < 12; i++) { x.v[i] = out[i] + threadIdx.x; y.v[i] = out[12 + i]; }
  for (uint32_t i = 0; i < n; i++) { x = mul_byval(x, y); y = mul_byval(y, x); }
  for (int i = 0; i < 12; i++) out[threadIdx.x * 12 + i] = x.v[i] ^ y.v[i];
}
__global__ void k2(uint32_t* out, uint32_t n) {
  V x, y;
  for (int i = 0; i < 12; i++) { x.v[i] = out[i] + threadIdx.x; y.v[i] = out[12 + i]; }
  for (uint32_t i = 0; i < n; i++) { mul_byref(x, x, y); mul_byref(y, y, x); }
  for (int i = 0; i < 12; i++) out[threadIdx.x * 12 + i] = x.v[i] ^ y.v[i];
}
