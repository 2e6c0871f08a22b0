// smooth.cu -- stand-alone launch of the smoothness backward role (smooth.cuh).  Only the
// deterministic backward uses it: there grad_disp is overwritten first and the view-synthesis
// gradient is added on top in a fixed order.  The default path runs the same role inside the
// fused backward kernel's grid.
#include "smooth.cuh"

namespace ppea {

__global__ void __launch_bounds__(kSmoothThreads) smooth_backward_kernel(const __grid_constant__ VslArgs a) {
  smooth_backward_role<0>(a, blockIdx.x);
}

cudaError_t launch_smooth_backward(const VslArgs& a, cudaStream_t stream) {
  smooth_backward_kernel<<<a.S * a.B * kSmoothChunks, kSmoothThreads, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace ppea
