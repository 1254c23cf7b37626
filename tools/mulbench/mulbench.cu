// Field-multiplication microbenchmark: compares formulations of the 377-bit Montgomery multiplication
// in isolation (dependent chains per thread, ILP independent chains), to guide kernel design.
//   variant 0: current 12 x 32-bit CIOS with even/odd carry chains (csrc/fp.cuh)
//   variant 1: reduced radix 14 x 28-bit limbs, 64-bit column accumulators, no carry chains
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../snark-setup-operator_b200/csrc -o mulbench mulbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "fp.cuh"
using namespace sso;

// ---------------- variant 1: reduced radix ----------------
static constexpr int NL = 14, W = 28;
static constexpr uint32_t MASK = (1u << W) - 1;
__device__ __constant__ uint32_t c_p28[NL];
__device__ __constant__ uint32_t c_inv28;

__device__ __forceinline__ void rr_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint64_t c[2 * NL];
#pragma unroll
  for (int k = 0; k < 2 * NL; k++) c[k] = 0;
#pragma unroll
  for (int i = 0; i < NL; i++)
#pragma unroll
    for (int j = 0; j < NL; j++) c[i + j] += (uint64_t)a[i] * b[j];
#pragma unroll
  for (int i = 0; i < NL; i++) {
    uint32_t m = ((uint32_t)c[i] * c_inv28) & MASK;
#pragma unroll
    for (int j = 0; j < NL; j++) c[i + j] += (uint64_t)m * c_p28[j];
    c[i + 1] += c[i] >> W;
  }
  // two weak carry passes over the upper half
  uint64_t t[NL];
#pragma unroll
  for (int k = 0; k < NL; k++) t[k] = c[NL + k];
  uint32_t lo[NL];
  uint64_t carry = 0;
#pragma unroll
  for (int k = 0; k < NL; k++) { uint64_t v = (t[k] & MASK) + (k ? (t[k - 1] >> W) : 0); lo[k] = 0; t[k] = v; (void)carry; }
  // second pass (values now < 2^28 + 2^36 -> carry < 2^9)
#pragma unroll
  for (int k = 0; k < NL; k++) {
    uint32_t v = (uint32_t)(t[k] & MASK) + (k ? (uint32_t)(t[k - 1] >> W) : 0u);
    r[k] = v;
  }
  // top carry of limb NL-1 is kept inside the limb (value < 2p fits)
  r[NL - 1] += (uint32_t)((c[2 * NL - 1] >> W) << W) * 0u;
  (void)lo;
}

template <int VARIANT, int ILP>
__global__ void __launch_bounds__(128) k_mulbench(uint32_t iters, uint32_t seed, uint32_t* out) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (VARIANT == 0 || VARIANT == 3) {
    using F = Fp<P_q377>;
    typename F::T x[ILP], y[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++)
#pragma unroll
      for (int i = 0; i < 12; i++) { x[k].v[i] = seed * (i + 1) + tid + k; y[k].v[i] = seed * 7 + i * tid + k; x[k].v[11] &= 0xffffff; y[k].v[11] &= 0xffffff; }
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
      for (int k = 0; k < ILP; k++) x[k] = VARIANT == 3 ? F::add(F::sqr(x[k]), y[k]) : F::mul(x[k], y[k]);
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++)
#pragma unroll
      for (int i = 0; i < 12; i++) s ^= x[k].v[i];
    out[tid] = s;
  } else {
    uint32_t x[ILP][NL], y[ILP][NL];
#pragma unroll
    for (int k = 0; k < ILP; k++)
#pragma unroll
      for (int i = 0; i < NL; i++) { x[k][i] = (seed * (i + 1) + tid + k) & MASK; y[k][i] = (seed * 7 + i * tid + k) & MASK; }
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
      for (int k = 0; k < ILP; k++) {
        uint32_t r[NL];
        rr_mul(r, x[k], y[k]);
#pragma unroll
        for (int i = 0; i < NL; i++) x[k][i] = r[i] & MASK;
      }
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++)
#pragma unroll
      for (int i = 0; i < NL; i++) s ^= x[k][i];
    out[tid] = s;
  }
}

// variant 2: the 24-limb (761-bit) multiplication as the kernels call it (out of line, by value)
template <int ILP, bool SQR>
__global__ void __launch_bounds__(128) k_mulbench24(uint32_t iters, uint32_t seed, uint32_t* out) {
  using F = Fp<P_q761>;
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  typename F::T x[ILP], y[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++)
#pragma unroll
    for (int i = 0; i < 24; i++) { x[k].v[i] = seed * (i + 1) + tid + k; y[k].v[i] = seed * 7 + i * tid + k; x[k].v[23] &= 0xffffff; y[k].v[23] &= 0xffffff; }
  for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = SQR ? F::add(F::sqr(x[k]), y[k]) : F::mul(x[k], y[k]);
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++)
#pragma unroll
    for (int i = 0; i < 24; i++) s ^= x[k].v[i];
  out[tid] = s;
}
template <int ILP, bool SQR> void run24(int blocks_per_sm) {
  int dev = 0; cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
  uint32_t blocks = prop.multiProcessorCount * blocks_per_sm, threads = 128, iters = 1000;
  uint32_t* d_out; cudaMalloc(&d_out, (size_t)blocks * threads * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    k_mulbench24<ILP, SQR><<<blocks, threads>>>(iters, 12345 + rep, d_out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_mulbench24<ILP, SQR>);
  double muls = (double)blocks * threads * iters * ILP;
  printf("%-28s ilp=%d blocks/SM=%d regs=%d  %.3f ms  %.2f Gmul/s = %.2f TMAC/s (err=%s)\n", SQR ? "sqr+add 24-limb (q761)" : "cios32 24-limb (q761)", ILP, blocks_per_sm, fa.numRegs, best,
         muls / best / 1e6, muls / best / 1e6 * 1176 / 1e3, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d_out);
}

template <int VARIANT, int ILP> void run(const char* name, int blocks_per_sm) {
  int dev = 0; cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
  uint32_t blocks = prop.multiProcessorCount * blocks_per_sm, threads = 128, iters = 2000;
  uint32_t* d_out; cudaMalloc(&d_out, (size_t)blocks * threads * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    k_mulbench<VARIANT, ILP><<<blocks, threads>>>(iters, 12345 + rep, d_out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_mulbench<VARIANT, ILP>);
  double muls = (double)blocks * threads * iters * ILP;
  printf("%-28s ilp=%d blocks/SM=%d regs=%d  %.3f ms  %.2f Gmul/s  (err=%s)\n", name, ILP, blocks_per_sm, fa.numRegs, best, muls / best / 1e6,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(d_out);
}

int main() {
  // 28-bit limbs of q377 and -p^-1 mod 2^28
  const char* hex = "1ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001";
  unsigned __int128 dummy = 0; (void)dummy;
  // parse hex into 28-bit limbs
  uint32_t limbs[NL] = {0};
  int len = 0; while (hex[len]) len++;
  int bitpos = 0;
  for (int i = len - 1; i >= 0; i--) {
    char ch = hex[i]; uint32_t d = ch <= '9' ? ch - '0' : ch - 'a' + 10;
    for (int b = 0; b < 4; b++) { if ((d >> b) & 1) limbs[(bitpos + b) / W] |= 1u << ((bitpos + b) % W); }
    bitpos += 4;
  }
  uint32_t x = 1; for (int i = 0; i < 6; i++) x *= 2 - limbs[0] * x;
  uint32_t inv = (0u - x) & MASK;
  cudaMemcpyToSymbol(c_p28, limbs, sizeof limbs); cudaMemcpyToSymbol(c_inv28, &inv, 4);
  for (int bps : {2, 4}) { run24<1, false>(bps); run24<1, true>(bps); }
  for (int bps : {2, 4, 8}) {
    run<0, 1>("cios32 even/odd", bps);
    run<0, 2>("cios32 even/odd", bps);
    run<3, 1>("dedicated sqr (+add)", bps);
    run<1, 1>("reduced radix 14x28", bps);
    run<1, 2>("reduced radix 14x28", bps);
  }
  return 0;
}
