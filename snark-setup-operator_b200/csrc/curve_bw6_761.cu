// Kernel instantiations for bw6_761 (G1 and G2).
#include "curve_ops.cuh"
namespace sso {
const CurveOps* curve_ops_bw6_761() { return CurveImpl<Bw6_761_G1, Bw6_761_G2, PAIR_bw6_761>::ops(); }
}  // namespace sso
