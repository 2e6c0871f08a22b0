// smooth.cu -- edge-aware disparity smoothness of the fused loss path.
//
// Reference: trainer.py:1147-1149 (mean-normalised disparity) feeding
// get_smooth_loss, layers.py:210-223:
//   nd = disp_s / (mean_hw(disp_s) + 1e-7)
//   smooth = mean(|nd[x]-nd[x+1]| * exp(-mean_c|I[x]-I[x+1]|)) + (same along y)
// All scales of a call are handled by one launch (blockIdx.z = scale,
// blockIdx.y = image, blockIdx.x = chunk of the image).  The per-image mean
// makes the term two-phase: `smooth_disp_sums` runs before the stencil.
//
// Backward uses that the stencil is positively homogeneous of degree 1 in nd:
//   d smooth / d disp_j = ( g_j - S_b / (h*w) ) / (m_b + 1e-7)
// with g = d smooth / d nd (sign stencil) and S_b = sum_i g_i * nd_i, which equals
// image b's own contribution to `smooth` (already reduced by the forward).
#include "vsl_common.cuh"

namespace ppea {

namespace {

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  return t;   // valid in thread 0
}

__device__ __forceinline__ void chunk_range(int n, int chunk, int& lo, int& hi) {
  const int per = (n + kSmoothChunks - 1) / kSmoothChunks;
  lo = chunk * per;
  hi = lo + per < n ? lo + per : n;
}

__device__ __forceinline__ float edge_weight(const float* __restrict__ img, size_t plane, int i, int j) {
  // exp(-mean_c |I[i] - I[j]|)   (layers.py:217-221)
  float g = fabsf(__ldg(img + i) - __ldg(img + j));
  g += fabsf(__ldg(img + plane + i) - __ldg(img + plane + j));
  g += fabsf(__ldg(img + 2 * plane + i) - __ldg(img + 2 * plane + j));
  return __expf(-g * (1.f / 3.f));
}

}  // namespace

__global__ void __launch_bounds__(kSmoothThreads) smooth_disp_sums_kernel(const __grid_constant__ VslArgs a) {
  __shared__ float red[kSmoothThreads / 32];
  const int s = blockIdx.z, b = blockIdx.y, chunk = blockIdx.x;
  const ScaleArgs& sc = a.sc[s];
  const int n = sc.hs * sc.ws;
  int lo, hi;
  chunk_range(n, chunk, lo, hi);
  const float* d = sc.disp + (size_t)b * n;
  float v = 0.f;
  for (int i = lo + threadIdx.x; i < hi; i += kSmoothThreads) v += __ldg(d + i);
  const float t = block_sum(v, red);
  if (threadIdx.x == 0) a.smooth_ws[(((size_t)s * a.B + b) * kSmoothChunks + chunk) * 3 + 0] = t;
}

__device__ __forceinline__ float image_mean(const VslArgs& a, int s, int b, int n) {
  // fixed-order sum of the chunk partials => every block of the image gets the same mean
  const float* w = a.smooth_ws + ((size_t)s * a.B + b) * kSmoothChunks * 3;
  double t = 0;
  for (int c = 0; c < kSmoothChunks; ++c) t += (double)w[c * 3];
  return (float)(t / (double)n);
}

__global__ void __launch_bounds__(kSmoothThreads) smooth_forward_kernel(const __grid_constant__ VslArgs a) {
  __shared__ float red[kSmoothThreads / 32];
  const int s = blockIdx.z, b = blockIdx.y, chunk = blockIdx.x;
  const ScaleArgs& sc = a.sc[s];
  const int h = sc.hs, w = sc.ws, n = h * w;
  const float inv = 1.f / (image_mean(a, s, b, n) + 1e-7f);
  int lo, hi;
  chunk_range(n, chunk, lo, hi);
  const float* d = sc.disp + (size_t)b * n;
  const float* img = sc.color + (size_t)b * 3 * n;
  float sx = 0.f, sy = 0.f;
  for (int i = lo + threadIdx.x; i < hi; i += kSmoothThreads) {
    const int y = i / w, x = i - y * w;
    // mul_rn: every normalised value is rounded once, exactly as the reference's division does; a
    // contracted fma(-d_j, inv, nd_i) would turn exact ties (equal disparities) into +-1 ulp noise
    // whose SIGN the backward would then propagate.
    const float di = mul_rn(__ldg(d + i), inv);
    if (x + 1 < w) sx += fabsf(di - mul_rn(__ldg(d + i + 1), inv)) * edge_weight(img, n, i, i + 1);
    if (y + 1 < h) sy += fabsf(di - mul_rn(__ldg(d + i + w), inv)) * edge_weight(img, n, i, i + w);
  }
  const float tx = block_sum(sx, red);
  const float ty = block_sum(sy, red);
  if (threadIdx.x == 0) {
    float* o = a.smooth_ws + (((size_t)s * a.B + b) * kSmoothChunks + chunk) * 3;
    o[1] = tx;
    o[2] = ty;
  }
}

// Overwrites grad_disp_s with the smoothness gradient (the reprojection /
// consistency gradients are accumulated on top by the view-synthesis backward).
__global__ void __launch_bounds__(kSmoothThreads) smooth_backward_kernel(const __grid_constant__ VslArgs a) {
  const int s = blockIdx.z, b = blockIdx.y, chunk = blockIdx.x;
  const ScaleArgs& sc = a.sc[s];
  const int h = sc.hs, w = sc.ws, n = h * w;
  const float* row = a.sums + (size_t)s * sums_stride(a.B) + PPEA_SUMS_PER_SCALE + 4 * b;
  const float inv = 1.f / (row[0] / (float)n + 1e-7f);
  const float g = scale_grads(a, s).smooth;
  const float gx = g / ((float)a.B * h * (w - 1)), gy = g / ((float)a.B * (h - 1) * w);
  const float S_b = gx * row[1] + gy * row[2];
  const float mean_term = S_b / (float)n;
  int lo, hi;
  chunk_range(n, chunk, lo, hi);
  const float* d = sc.disp + (size_t)b * n;
  const float* img = sc.color + (size_t)b * 3 * n;
  float* out = sc.grad_disp + (size_t)b * n;
  for (int i = lo + threadIdx.x; i < hi; i += kSmoothThreads) {
    const int y = i / w, x = i - y * w;
    const float di = mul_rn(__ldg(d + i), inv);
    float acc = 0.f;
    auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
    if (x + 1 < w) acc += gx * sgn(di - mul_rn(__ldg(d + i + 1), inv)) * edge_weight(img, n, i, i + 1);
    if (x > 0) acc -= gx * sgn(mul_rn(__ldg(d + i - 1), inv) - di) * edge_weight(img, n, i - 1, i);
    if (y + 1 < h) acc += gy * sgn(di - mul_rn(__ldg(d + i + w), inv)) * edge_weight(img, n, i, i + w);
    if (y > 0) acc -= gy * sgn(mul_rn(__ldg(d + i - w), inv) - di) * edge_weight(img, n, i - w, i);
    out[i] = (acc - mean_term) * inv;
  }
}

cudaError_t launch_smooth_disp_sums(const VslArgs& a, cudaStream_t stream) {
  smooth_disp_sums_kernel<<<dim3(kSmoothChunks, a.B, a.S), kSmoothThreads, 0, stream>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_smooth_forward(const VslArgs& a, cudaStream_t stream) {
  smooth_forward_kernel<<<dim3(kSmoothChunks, a.B, a.S), kSmoothThreads, 0, stream>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_smooth_backward(const VslArgs& a, cudaStream_t stream) {
  smooth_backward_kernel<<<dim3(kSmoothChunks, a.B, a.S), kSmoothThreads, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace ppea
