// Kernel instantiations for mnt4_753 (G1 and G2).
#include "curve_ops.cuh"
namespace sso {
const CurveOps* curve_ops_mnt4_753() { return CurveImpl<Mnt4_753_G1, Mnt4_753_G2, PAIR_mnt4_753>::ops(); }
}  // namespace sso
