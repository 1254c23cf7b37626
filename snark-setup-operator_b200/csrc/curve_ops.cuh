// Per-curve kernel instantiations behind a small function table, so that each curve compiles
// in its own translation unit (parallel nvcc) and abi.cu stays curve-agnostic.
#pragma once
#include <type_traits>
#include "host.cuh"
#include "kernels.cuh"

namespace sso {

struct CurveOps {
  // K1: tau-power tables + coefficient slots for one call (shared by every vector of the call)
  int (*tau_tables)(Ctx& c, int si, uint64_t first_index, const uint8_t* tau, const uint8_t* const coeffs[TAU_COEFF_SLOTS],
                    uint32_t** d_table, char* err, size_t errcap);
  // K2-K4 on up to four vectors of one group in one launch pair
  int (*batch_exp)(Ctx& c, int si, uint32_t group, const VecBatch& batch, uint32_t in_compressed, const uint32_t* d_table,
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
__global__ void __launch_bounds__(128) k_batch_exp(const __grid_constant__ VecBatch batch, uint32_t in_compressed,
                                                    const uint32_t* table, uint32_t check, uint32_t* jac_out, uint32_t* status) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  body_batch_exp<G>(tid, batch, in_compressed, table, check, jac_out, status);
}

template <class G>
__global__ void __launch_bounds__(128) k_normalize_write(const __grid_constant__ VecBatch batch, const uint32_t* jac,
                                                          uint32_t out_compressed) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  body_normalize_write<G>(tid, batch, jac, out_compressed);
}

template <class G>
__global__ void __launch_bounds__(128) k_reencode(uint32_t n, const uint8_t* in, uint32_t in_compressed, uint8_t* out,
                                                   uint32_t out_compressed, uint32_t check, uint32_t subgroup, uint32_t* aff_out,
                                                   uint32_t* status) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  body_reencode<G>(tid, n, in, in_compressed, out, out_compressed, check, subgroup, aff_out, status);
}

// tau tables for one call: powers of tau from first_index plus up to three coefficient slots
template <class Fr>
inline int run_tau_tables(Ctx& c, int si, uint64_t first_index, const uint8_t* tau, const uint8_t* const coeffs[TAU_COEFF_SLOTS],
                          uint32_t** d_table, char* err, size_t errcap) {
  c.staging.emplace_back((size_t)(1 + TAU_COEFF_SLOTS) * Fr::L, 0u);
  std::vector<uint32_t>& w = c.staging.back();
  if (tau) memcpy(w.data(), tau, Fr::NBYTES); else w[0] = 1;
  for (int i = 0; i < TAU_COEFF_SLOTS; i++) {
    if (coeffs && coeffs[i]) memcpy(w.data() + (size_t)(1 + i) * Fr::L, coeffs[i], Fr::NBYTES);
    else w[(size_t)(1 + i) * Fr::L] = 1;
  }
  uint32_t* d_sc;
  int rc;
  if ((rc = c.alloc((void**)&d_sc, w.size() * 4, si))) return rc;
  if ((rc = c.alloc((void**)d_table, (size_t)TAU_TABLE_ELEMS * Fr::L * 4, si))) return rc;
  CUDA_TRY(cudaMemcpyAsync(d_sc, w.data(), w.size() * 4, cudaMemcpyHostToDevice, c.s[si]));
  c.begin(PK_TAU_TABLES, si, TAU_TABLE_ELEMS);
  k_tau_tables<Fr><<<div_up(TAU_TABLE_ELEMS, 128), 128, 0, c.s[si]>>>(d_sc, d_sc + Fr::L, first_index, *d_table);
  c.end(si);
  CUDA_TRY(cudaGetLastError());
  return SSO_OK;
}

// K2-K4 for up to four vectors of one group in ONE launch pair (batch_exp + normalize) on stream si
template <class G>
inline int run_batch_exp(Ctx& c, int si, const VecBatch& batch, uint32_t in_compressed, const uint32_t* d_table,
                         uint32_t out_compressed, uint32_t check, uint32_t* d_status, char* err, size_t errcap) {
  using F = typename G::F;
  uint64_t n = batch.total;
  if (n == 0) return SSO_OK;
  for (uint32_t i = 0; i < batch.nseg; i++)
    if (batch.seg[i].n > (1u << 24)) { set_err(err, errcap, "vector longer than 2^24 elements: split the call"); return SSO_E_ARG; }
  cudaStream_t st = c.s[si];
  uint32_t* d_jac = nullptr;
  int rc;
  if ((rc = c.alloc((void**)&d_jac, (size_t)n * 3 * F::WORDS * 4, si))) return rc;
  constexpr bool IS_G1 = G::GROUP == 0;
  c.begin(IS_G1 ? PK_BATCH_EXP_G1 : PK_BATCH_EXP_G2, si, n);
  k_batch_exp<G><<<div_up(n, 128), 128, 0, st>>>(batch, in_compressed, d_table, check, d_jac, d_status);
  c.end(si);
  c.begin(IS_G1 ? PK_NORMALIZE_G1 : PK_NORMALIZE_G2, si, n);
  k_normalize_write<G><<<div_up(div_up(n, NORM_BATCH), 128), 128, 0, st>>>(batch, d_jac, out_compressed);
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
  static int tau_tables(Ctx& c, int si, uint64_t first_index, const uint8_t* tau, const uint8_t* const coeffs[TAU_COEFF_SLOTS],
                        uint32_t** d_table, char* err, size_t errcap) {
    return run_tau_tables<typename G1::Fr>(c, si, first_index, tau, coeffs, d_table, err, errcap);
  }
  static int batch_exp(Ctx& c, int si, uint32_t group, const VecBatch& batch, uint32_t in_compressed, const uint32_t* d_table,
                       uint32_t out_compressed, uint32_t check, uint32_t* d_status, char* err, size_t errcap) {
    if (group == GROUP_G1) return run_batch_exp<G1>(c, si, batch, in_compressed, d_table, out_compressed, check, d_status, err, errcap);
    if (group == GROUP_G2) return run_batch_exp<G2>(c, si, batch, in_compressed, d_table, out_compressed, check, d_status, err, errcap);
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
    static const CurveOps o = {&tau_tables, &batch_exp, &reencode, &fill_generator, (uint32_t)G1::Fr::NBYTES};
    return &o;
  }
};

// one getter per curve, each defined in its own .cu
const CurveOps* curve_ops_bls12_377();
const CurveOps* curve_ops_bw6_761();
const CurveOps* curve_ops_mnt4_753();
const CurveOps* curve_ops_mnt6_753();

}  // namespace sso
