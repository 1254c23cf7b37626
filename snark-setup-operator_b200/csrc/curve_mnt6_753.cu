// Kernel instantiations for mnt6_753 (G1 and G2).
#include "curve_ops.cuh"
namespace sso {
const CurveOps* curve_ops_mnt6_753() { return CurveImpl<Mnt6_753_G1, Mnt6_753_G2, PAIR_mnt6_753>::ops(); }
}  // namespace sso
