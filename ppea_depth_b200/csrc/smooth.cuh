// smooth.cuh -- edge-aware disparity smoothness of the fused loss path (device code).
//
// Reference: trainer.py:1147-1149 (mean-normalised disparity) feeding get_smooth_loss, layers.py:210-223:
//   nd = disp_s / (mean_hw(disp_s) + 1e-7)
//   smooth = mean(|nd[x]-nd[x+1]| * exp(-mean_c|I[x]-I[x+1]|)) + (same along y)
// The stencil is positively homogeneous of degree 1, so with inv_b = 1/(mean_b + 1e-7) per image
//   smooth = sum_b inv_b * ( X_b / N_x + Y_b / N_y ),   X_b = sum |d[x]-d[x+1]| e_x,  Y_b likewise,
// i.e. ONE pass over disp_s / colour_s collecting (sum d, X_b, Y_b) per image -- no second pass
// after the mean -- and, in the backward,
//   d smooth / d d_j = inv_b * ( g_j - S_b / (h*w) ),  g_j = sum over the 4 edges of +-sign(d_i-d_j) e / N,
//   S_b = inv_b * (X_b / N_x + Y_b / N_y)   (image b's own share of the loss, from the forward).
// Signs are taken on the raw disparities: sign(nd_i - nd_j) == sign(d_i - d_j) exactly, ties included.
//
// These are *roles* run by extra CTAs appended to the grids of the two fused kernels (chunks of one
// image of one scale), so the smoothness term costs no launches of its own.
#pragma once

#include "vsl_common.cuh"

namespace ppea {

__device__ __forceinline__ void smooth_chunk_range(int n, int chunk, int& lo, int& hi) {
  const int per = (n + kSmoothChunks - 1) / kSmoothChunks;
  lo = chunk * per;
  hi = lo + per < n ? lo + per : n;
}

__device__ __forceinline__ float smooth_edge_weight(const float* __restrict__ img, unsigned plane, unsigned i, unsigned j) {
  // exp(-mean_c |I[i] - I[j]|)   (layers.py:217-221)
  float g = fabsf(__ldg(img + i) - __ldg(img + j));
  g += fabsf(__ldg(img + plane + i) - __ldg(img + plane + j));
  g += fabsf(__ldg(img + 2u * plane + i) - __ldg(img + 2u * plane + j));
  return __expf(-g * (1.f / 3.f));
}

// role id -> (scale, image, chunk)
__device__ __forceinline__ void smooth_role_ids(const VslArgs& a, int role, int& s, int& b, int& chunk) {
  chunk = role % kSmoothChunks;
  role /= kSmoothChunks;
  b = role % a.B;
  s = role / a.B;
}

// Forward role: raw sums of one chunk -> smooth_ws[(s, b, chunk)][3] = (sum d, X, Y).  `red` holds 3*nwarps floats.
__device__ __forceinline__ void smooth_forward_role(const VslArgs& a, int role, float* red) {
  int s, b, chunk;
  smooth_role_ids(a, role, s, b, chunk);
  const ScaleArgs& sc = a.sc[s];
  const int h = sc.hs, w = sc.ws, n = h * w;
  int lo, hi;
  smooth_chunk_range(n, chunk, lo, hi);
  const float* d = sc.disp + (size_t)b * n;
  const float* img = sc.color + (size_t)b * 3 * n;
  float* gz = sc.grad_disp ? sc.grad_disp + (size_t)b * n : nullptr;   // optional: pre-zero the backward's accumulator
  float sd = 0.f, sx = 0.f, sy = 0.f;
#pragma unroll 2
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const int y = i / w, x = i - y * w;
    const float di = __ldg(d + i);
    if (gz) gz[i] = 0.f;
    sd += di;
    if (x + 1 < w) sx += fabsf(di - __ldg(d + i + 1)) * smooth_edge_weight(img, n, i, i + 1);
    if (y + 1 < h) sy += fabsf(di - __ldg(d + i + w)) * smooth_edge_weight(img, n, i, i + w);
  }
  sd = warp_sum(sd);
  sx = warp_sum(sx);
  sy = warp_sum(sy);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) {
    red[wid] = sd;
    red[nw + wid] = sx;
    red[2 * nw + wid] = sy;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int k = 0; k < nw; ++k) t += red[threadIdx.x * nw + k];
    a.smooth_ws[(((size_t)s * a.B + b) * kSmoothChunks + chunk) * 3 + threadIdx.x] = t;
  }
}

// Backward role: smoothness gradient of one chunk.  ACCUMULATE: atomically add into a zero-initialised
// (and concurrently accumulated) grad_disp;  otherwise overwrite it.
template <bool ACCUMULATE>
__device__ __forceinline__ void smooth_backward_role(const VslArgs& a, int role) {
  int s, b, chunk;
  smooth_role_ids(a, role, s, b, chunk);
  const ScaleArgs& sc = a.sc[s];
  const int h = sc.hs, w = sc.ws, n = h * w;
  const float* row = a.sums + (size_t)s * sums_stride(a.B) + PPEA_SUMS_PER_SCALE + 4 * b;   // (sum d, X_b, Y_b)
  const float inv = 1.f / (row[0] / (float)n + 1e-7f);
  const float g = scale_grads(a, s).smooth;
  const float gx = g / ((float)a.B * h * (w - 1)), gy = g / ((float)a.B * (h - 1) * w);
  const float mean_term = inv * (gx * row[1] + gy * row[2]) / (float)n;
  int lo, hi;
  smooth_chunk_range(n, chunk, lo, hi);
  const float* d = sc.disp + (size_t)b * n;
  const float* img = sc.color + (size_t)b * 3 * n;
  float* out = sc.grad_disp + (size_t)b * n;
#pragma unroll 2
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const int y = i / w, x = i - y * w;
    const float di = __ldg(d + i);
    float acc = 0.f;
    if (x + 1 < w) acc += gx * sign_of(di - __ldg(d + i + 1)) * smooth_edge_weight(img, n, i, i + 1);
    if (x > 0) acc -= gx * sign_of(__ldg(d + i - 1) - di) * smooth_edge_weight(img, n, i - 1, i);
    if (y + 1 < h) acc += gy * sign_of(di - __ldg(d + i + w)) * smooth_edge_weight(img, n, i, i + w);
    if (y > 0) acc -= gy * sign_of(__ldg(d + i - w) - di) * smooth_edge_weight(img, n, i - w, i);
    const float v = (acc - mean_term) * inv;
    if (ACCUMULATE)
      atomicAdd(out + i, v);
    else
      out[i] = v;
  }
}

}  // namespace ppea
