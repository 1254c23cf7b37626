// Per-curve kernel instantiations behind a small function table, so that each curve compiles
// in its own translation unit (parallel nvcc) and abi.cu stays curve-agnostic.
#pragma once
#include <type_traits>
#include "host.cuh"
#include "kernels.cuh"

namespace sso {

struct CurveOps {
  // K1-K4 on one vector (mode 0: per-index tau powers, mode 1: one shared scalar)
  int (*batch_exp)(Ctx& c, int si, uint32_t group, const uint8_t* d_in, uint32_t in_compressed, uint64_t n,
                   uint64_t first_index, const uint8_t* tau, const uint8_t* coeff, uint32_t mode, uint8_t* d_out,
                   uint32_t out_compressed, uint32_t check, uint32_t* d_status, char* err, size_t errcap);
  // K3 (+K6) alone
  int (*reencode)(Ctx& c, int si, uint32_t group, const uint8_t* d_in, uint32_t in_compressed, uint64_t n, uint8_t* d_out,
                  uint32_t out_compressed, uint32_t check, uint32_t subgroup, uint32_t* d_aff, uint32_t* d_status, char* err,
                  size_t errcap);
  // phase1_cli::new_challenge: n copies of the group generator, uncompressed or compressed
  int (*fill_generator)(Ctx& c, int si, uint32_t group, uint64_t n, uint8_t* d_out, uint32_t out_compressed, char* err, size_t errcap);
  uint32_t fr_bytes;
};

template <class Fr>
__global__ void k_tau_tables(const uint32_t* tau_canon, const uint32_t* coeff_canon, uint64_t first_index, uint32_t* table) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid < (uint32_t)TAU_TABLE_ELEMS) body_tau_tables<Fr>(tid, tau_canon, coeff_canon, first_index, table);
}

template <class G>
__global__ void __launch_bounds__(128) k_batch_exp(uint32_t n, const uint8_t* in, uint32_t in_compressed, const uint32_t* table,
                                                    uint32_t has_coeff, uint32_t mode, uint32_t check, uint32_t* jac_out,
                                                    uint32_t* status) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  body_batch_exp<G>(tid, n, in, in_compressed, table, has_coeff, mode, check, jac_out, status);
}

template <class G>
__global__ void __launch_bounds__(128) k_normalize_write(uint32_t n, const uint32_t* jac, uint8_t* out, uint32_t out_compressed) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  body_normalize_write<G>(tid, n, jac, out, out_compressed);
}

template <class G>
__global__ void __launch_bounds__(128) k_reencode(uint32_t n, const uint8_t* in, uint32_t in_compressed, uint8_t* out,
                                                   uint32_t out_compressed, uint32_t check, uint32_t subgroup, uint32_t* aff_out,
                                                   uint32_t* status) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  body_reencode<G>(tid, n, in, in_compressed, out, out_compressed, check, subgroup, aff_out, status);
}

// One vector through K1-K4 on stream si.  mode 0: per-index tau powers; mode 1: one shared scalar.
template <class G>
inline int run_batch_exp(Ctx& c, int si, const uint8_t* d_in, uint32_t in_compressed, uint64_t n, uint64_t first_index,
                  const uint8_t* tau, const uint8_t* coeff, uint32_t mode, uint8_t* d_out, uint32_t out_compressed,
                  uint32_t check, uint32_t* d_status, char* err, size_t errcap) {
  using Fr = typename G::Fr;
  using F = typename G::F;
  if (n == 0) return SSO_OK;
  if (n > (1ull << 24)) { set_err(err, errcap, "vector longer than 2^24 elements: split the call"); return SSO_E_ARG; }
  cudaStream_t st = c.s[si];
  uint32_t *d_tau = nullptr, *d_coeff = nullptr, *d_table = nullptr, *d_jac = nullptr;
  int rc;
  if ((rc = upload_scalar(c, tau, Fr::NBYTES, Fr::L, &d_tau, si, err, errcap))) return rc;
  if ((rc = upload_scalar(c, coeff, Fr::NBYTES, Fr::L, &d_coeff, si, err, errcap))) return rc;
  if ((rc = c.alloc((void**)&d_table, (size_t)TAU_TABLE_ELEMS * Fr::L * 4, si))) return rc;
  if ((rc = c.alloc((void**)&d_jac, (size_t)n * 3 * F::WORDS * 4, si))) return rc;
  constexpr bool IS_G1 = G::GROUP == 0;
  c.begin(PK_TAU_TABLES, si, TAU_TABLE_ELEMS);
  k_tau_tables<Fr><<<div_up(TAU_TABLE_ELEMS, 128), 128, 0, st>>>(d_tau, d_coeff, first_index, d_table);
  c.end(si);
  c.begin(IS_G1 ? PK_BATCH_EXP_G1 : PK_BATCH_EXP_G2, si, n);
  k_batch_exp<G><<<div_up(n, 128), 128, 0, st>>>((uint32_t)n, d_in, in_compressed, d_table, coeff != nullptr, mode, check, d_jac, d_status);
  c.end(si);
  c.begin(IS_G1 ? PK_NORMALIZE_G1 : PK_NORMALIZE_G2, si, n);
  k_normalize_write<G><<<div_up(div_up(n, NORM_BATCH), 128), 128, 0, st>>>((uint32_t)n, d_jac, d_out, out_compressed);
  c.end(si);
  CUDA_TRY(cudaGetLastError());
  return SSO_OK;
}


template <class G>
__global__ void k_fill_generator(uint32_t n, uint8_t* out, uint32_t out_compressed) {
  using C = SW<G>;
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n) return;
  typename C::Affine g{G::gen_x(), G::gen_y(), false};
  if (out_compressed) C::write_compressed(out + (size_t)tid * C::SIZE_C, g);
  else C::write_uncompressed(out + (size_t)tid * C::SIZE_U, g);
}

template <class G>
inline int run_fill_generator(Ctx& c, int si, uint64_t n, uint8_t* d_out, uint32_t out_compressed, char* err, size_t errcap) {
  if (n == 0) return SSO_OK;
  if (n > 0xffffffffull) { set_err(err, errcap, "vector too long"); return SSO_E_ARG; }
  c.begin(PK_FILL, si, n);
  k_fill_generator<G><<<div_up(n, 128), 128, 0, c.s[si]>>>((uint32_t)n, d_out, out_compressed);
  c.end(si);
  CUDA_TRY(cudaGetLastError());
  return SSO_OK;
}

template <class G>
inline int run_reencode(Ctx& c, int si, const uint8_t* d_in, uint32_t in_compressed, uint64_t n, uint8_t* d_out,
                        uint32_t out_compressed, uint32_t check, uint32_t subgroup, uint32_t* d_aff, uint32_t* d_status,
                        char* err, size_t errcap) {
  if (n == 0) return SSO_OK;
  if (n > 0xffffffffull) { set_err(err, errcap, "vector too long"); return SSO_E_ARG; }
  constexpr bool IS_G1 = G::GROUP == 0;
  c.begin(IS_G1 ? PK_REENCODE_G1 : PK_REENCODE_G2, si, n);
  k_reencode<G><<<div_up(n, 128), 128, 0, c.s[si]>>>((uint32_t)n, d_in, in_compressed, d_out, out_compressed, check, subgroup, d_aff, d_status);
  c.end(si);
  CUDA_TRY(cudaGetLastError());
  return SSO_OK;
}

template <class G1, class G2> struct CurveImpl {
  static int batch_exp(Ctx& c, int si, uint32_t group, const uint8_t* d_in, uint32_t in_compressed, uint64_t n,
                       uint64_t first_index, const uint8_t* tau, const uint8_t* coeff, uint32_t mode, uint8_t* d_out,
                       uint32_t out_compressed, uint32_t check, uint32_t* d_status, char* err, size_t errcap) {
    if (group == GROUP_G1) return run_batch_exp<G1>(c, si, d_in, in_compressed, n, first_index, tau, coeff, mode, d_out, out_compressed, check, d_status, err, errcap);
    if (group == GROUP_G2) return run_batch_exp<G2>(c, si, d_in, in_compressed, n, first_index, tau, coeff, mode, d_out, out_compressed, check, d_status, err, errcap);
    set_err(err, errcap, "unknown group %u", group);
    return SSO_E_ARG;
  }
  static int reencode(Ctx& c, int si, uint32_t group, const uint8_t* d_in, uint32_t in_compressed, uint64_t n, uint8_t* d_out,
                      uint32_t out_compressed, uint32_t check, uint32_t subgroup, uint32_t* d_aff, uint32_t* d_status, char* err,
                      size_t errcap) {
    if (group == GROUP_G1) return run_reencode<G1>(c, si, d_in, in_compressed, n, d_out, out_compressed, check, subgroup, d_aff, d_status, err, errcap);
    if (group == GROUP_G2) return run_reencode<G2>(c, si, d_in, in_compressed, n, d_out, out_compressed, check, subgroup, d_aff, d_status, err, errcap);
    set_err(err, errcap, "unknown group %u", group);
    return SSO_E_ARG;
  }
  static int fill_generator(Ctx& c, int si, uint32_t group, uint64_t n, uint8_t* d_out, uint32_t out_compressed, char* err, size_t errcap) {
    if (group == GROUP_G1) return run_fill_generator<G1>(c, si, n, d_out, out_compressed, err, errcap);
    if (group == GROUP_G2) return run_fill_generator<G2>(c, si, n, d_out, out_compressed, err, errcap);
    set_err(err, errcap, "unknown group %u", group);
    return SSO_E_ARG;
  }
  static const CurveOps* ops() {
    static const CurveOps o = {&batch_exp, &reencode, &fill_generator, (uint32_t)G1::Fr::NBYTES};
    return &o;
  }
};

// one getter per curve, each defined in its own .cu
const CurveOps* curve_ops_bls12_377();
const CurveOps* curve_ops_bw6_761();
const CurveOps* curve_ops_mnt4_753();
const CurveOps* curve_ops_mnt6_753();

}  // namespace sso
