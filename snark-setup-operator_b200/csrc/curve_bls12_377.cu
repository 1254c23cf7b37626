// Kernel instantiations for bls12_377 (G1 and G2).
#include "curve_ops.cuh"
namespace sso {
const CurveOps* curve_ops_bls12_377() { return CurveImpl<Bls12_377_G1, Bls12_377_G2, PAIR_bls12_377>::ops(); }
}  // namespace sso
