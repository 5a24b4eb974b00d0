// Instantiation of the per-point backward for padded input dim DP = 16 (split across files to build in parallel).
#include "gpblur_point_bwd.cuh"

namespace gpblur {
int launch_point_backward_dp16(const PointBwdArgs& a, cudaStream_t st) { return dispatch_bwd_dp<16>(a, st); }
}  // namespace gpblur
