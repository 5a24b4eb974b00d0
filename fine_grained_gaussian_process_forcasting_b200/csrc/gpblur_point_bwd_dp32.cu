// Instantiation of the per-point backward for padded input dim DP = 32 (split across files to build in parallel).
#include "gpblur_point_bwd.cuh"

namespace gpblur {
int launch_point_backward_dp32(const PointBwdArgs& a, cudaStream_t st) { return dispatch_bwd_dp<32>(a, st); }
}  // namespace gpblur
