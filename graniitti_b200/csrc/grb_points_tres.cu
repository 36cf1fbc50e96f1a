// k_points instantiated for one amplitude kind (see grb_kernels_points.cuh)
#include "grb_kernels_points.cuh"

namespace grb {
template void launch_points_amp<AMP_TENSOR_RES>(const PointArgs &, int, unsigned, int, size_t, cudaStream_t);
}  // namespace grb
