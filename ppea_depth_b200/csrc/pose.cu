// pose.cu -- pose-network output -> 4x4 camera transform, forward and backward.
//
// Reference: `transformation_from_parameters` (layers.py:26-42) = rot_from_axisangle (:62-100) + get_translation_matrix
// (:45-59) + one (B,4,4) matmul: ~30 tiny launches forward and ~60 in autograd's backward, on the critical path between the
// pose network and the loss (the fused loss hands back dL/dT, this function carries it to the pose head).
// Here: one launch each way, one thread per batch item.  The forward follows the reference op by op in fp32; the backward
// evaluates the same program on dual numbers (value + 6 partials w.r.t. axis-angle and translation) and contracts the
// Jacobian with dL/dT -- no hand-derived Rodrigues adjoint to get wrong.  norm() at the origin has derivative 0, as in
// PyTorch's norm_backward.
#include "vsl_common.cuh"

namespace ppea {

template <int N>
struct Dual {
  float v;
  float d[N];
};
template <int N>
__device__ __forceinline__ Dual<N> dconst(float c) {
  Dual<N> r;
  r.v = c;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = 0.f;
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> dvar(float c, int k) {
  Dual<N> r = dconst<N>(c);
  r.d[k] = 1.f;
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = add_rn(a.v, b.v);
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = sub_rn(a.v, b.v);
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = mul_rn(a.v, b.v);
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = div_rn(a.v, b.v);
  const float ib = 1.f / b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> dneg(const Dual<N>& a) {
  Dual<N> r;
  r.v = -a.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> dsqrt(const Dual<N>& a) {
  Dual<N> r;
  r.v = sqrtf(a.v);
  const float k = r.v > 0.f ? 0.5f / r.v : 0.f;        // norm_backward: zero (sub)gradient at the origin
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = k * a.d[i];
  return r;
}
template <int N>
__device__ __forceinline__ void dsincos(const Dual<N>& a, Dual<N>& s, Dual<N>& c) {
  s.v = sinf(a.v);
  c.v = cosf(a.v);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    s.d[i] = c.v * a.d[i];
    c.d[i] = -s.v * a.d[i];
  }
}

// the reference program on a generic scalar type; M is row-major 4x4
template <int N>
__device__ __forceinline__ void pose_program(const Dual<N> (&aa)[3], const Dual<N> (&tr)[3], bool invert, Dual<N> (&M)[16]) {
  using D = Dual<N>;
  // rot_from_axisangle (layers.py:62-100)
  const D angle = dsqrt(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]);          // torch.norm(vec, 2, 2, True)
  const D den = angle + dconst<N>(1e-7f);
  const D x = aa[0] / den, y = aa[1] / den, z = aa[2] / den;
  D sa, ca;
  dsincos(angle, sa, ca);
  const D C = dconst<N>(1.f) - ca;
  const D xs = x * sa, ys = y * sa, zs = z * sa;
  const D xC = x * C, yC = y * C, zC = z * C;
  const D xyC = x * yC, yzC = y * zC, zxC = z * xC;
  D R[9];
  R[0] = x * xC + ca;
  R[1] = xyC - zs;
  R[2] = zxC + ys;
  R[3] = xyC + zs;
  R[4] = y * yC + ca;
  R[5] = yzC - xs;
  R[6] = zxC - ys;
  R[7] = yzC + xs;
  R[8] = z * zC + ca;
  const D zero = dconst<N>(0.f), one = dconst<N>(1.f);
  if (!invert) {
    // M = T @ R (layers.py:40-41): rotation block of R, translation column t
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) M[i * 4 + j] = R[i * 3 + j];
      M[i * 4 + 3] = tr[i];
    }
  } else {
    // M = R^T @ T(-t) (layers.py:33-39): last column = R^T (-t), accumulated in matmul order
    const D nt[3] = {dneg(tr[0]), dneg(tr[1]), dneg(tr[2])};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) M[i * 4 + j] = R[j * 3 + i];
      M[i * 4 + 3] = (R[0 * 3 + i] * nt[0] + R[1 * 3 + i] * nt[1]) + R[2 * 3 + i] * nt[2];
    }
  }
  M[12] = zero;
  M[13] = zero;
  M[14] = zero;
  M[15] = one;
}

__global__ void pose_to_matrix_forward_kernel(const float* __restrict__ aa, const float* __restrict__ tr, int invert,
                                              float* __restrict__ out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  Dual<1> a[3], t[3], M[16];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    a[i] = dconst<1>(aa[b * 3 + i]);
    t[i] = dconst<1>(tr[b * 3 + i]);
  }
  pose_program<1>(a, t, invert != 0, M);
#pragma unroll
  for (int e = 0; e < 16; ++e) out[b * 16 + e] = M[e].v;
}

__global__ void pose_to_matrix_backward_kernel(const float* __restrict__ aa, const float* __restrict__ tr, int invert,
                                               const float* __restrict__ gT, float* __restrict__ g_aa, float* __restrict__ g_tr,
                                               int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  Dual<6> a[3], t[3], M[16];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    a[i] = dvar<6>(aa[b * 3 + i], i);
    t[i] = dvar<6>(tr[b * 3 + i], 3 + i);
  }
  pose_program<6>(a, t, invert != 0, M);
  float g[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const float w = gT[b * 16 + e];
#pragma unroll
    for (int k = 0; k < 6; ++k) g[k] = fmaf(w, M[e].d[k], g[k]);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    g_aa[b * 3 + k] = g[k];
    g_tr[b * 3 + k] = g[3 + k];
  }
}

// Trainer.compute_matching_mask (trainer.py:859-869): where the cost volume's best depth and the teacher disagree by less
// than a factor of two either way.  matching_depth = 1 / lowest_cost; mask = ((md - mono) / mono < 1) * ((mono - md) / md < 1).
__global__ void __launch_bounds__(256) matching_mask_kernel(const float* __restrict__ mono_depth, const float* __restrict__ lowest_cost,
                                                            uint8_t* __restrict__ mask, size_t n) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float mono = __ldg(mono_depth + i);
  const float md = div_rn(1.f, __ldg(lowest_cost + i));
  const bool a = div_rn(sub_rn(md, mono), mono) < 1.f;
  const bool b = div_rn(sub_rn(mono, md), md) < 1.f;
  mask[i] = (a && b) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Glue between the multi-frame encoder and the loss (remainder), one launch:
//   lowest_cost_up  = F.interpolate(lowest_cost[:, None], [H, W], mode="nearest")[:, 0]            repdepth.py:615-617
//   consistency     = F.interpolate(confidence[:, None], [H, W], mode="nearest")[:, 0]             repdepth.py:618-620
//                     * compute_matching_mask(mono_depth, lowest_cost_up)                           trainer.py:450-451, :859-869
//   per image: min / max of mono_depth over the pixels (DepthBins.update's reductions, trainer.py:54-55)
// The reference spends two interpolate launches, eight elementwise ones and four reductions on it.
// nearest: src = floor(dst * (in / out)) clamped to in - 1  (ATen nearest_neighbor_compute_source_index).
// min / max: positive floats order like their bit patterns, so the block results go through integer atomics
// (order-independent, hence deterministic); minmax_bits must be pre-set to (0x7f800000, 0) per image.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) matching_glue_kernel(const float* __restrict__ lowest_lr, const float* __restrict__ conf_lr,
                                                            const float* __restrict__ mono_depth, float* __restrict__ lowest_up,
                                                            float* __restrict__ consistency, unsigned* __restrict__ minmax_bits, int h,
                                                            int w, int H, int W) {
  __shared__ float s_min[8], s_max[8];
  const int b = blockIdx.y;
  const unsigned n = (unsigned)(H * W);
  const unsigned i = blockIdx.x * 256 + threadIdx.x;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  float vmin = INFINITY, vmax = 0.f;
  if (i < n) {
    const int y = (int)(i / (unsigned)W), x = (int)(i - (unsigned)y * (unsigned)W);
    const int ys = min((int)floorf(mul_rn((float)y, sy)), h - 1), xs = min((int)floorf(mul_rn((float)x, sx)), w - 1);
    const size_t o = (size_t)b * n + i, ol = ((size_t)b * h + ys) * w + xs;
    const float lc = __ldg(lowest_lr + ol);
    const float mono = __ldg(mono_depth + o);
    const float md = div_rn(1.f, lc);
    const bool ok = div_rn(sub_rn(md, mono), mono) < 1.f && div_rn(sub_rn(mono, md), md) < 1.f;
    lowest_up[o] = lc;
    consistency[o] = mul_rn(__ldg(conf_lr + ol), ok ? 1.f : 0.f);
    vmin = vmax = mono;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  }
  if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = vmin, s_max[threadIdx.x >> 5] = vmax;
  __syncthreads();
  if (threadIdx.x == 0 && minmax_bits) {
    for (int k = 1; k < 8; ++k) vmin = fminf(vmin, s_min[k]), vmax = fmaxf(vmax, s_max[k]);
    atomicMin(minmax_bits + 2 * b, __float_as_uint(fmaxf(vmin, 0.f)));
    atomicMax(minmax_bits + 2 * b + 1, __float_as_uint(fmaxf(vmax, 0.f)));
  }
}

// DepthBins.update (trainer.py:52-64) on the device, from the per-image extrema: no tensor leaves the GPU, no host max().
//   min_d = max(opt_min_depth, mean_b(min_b) * 0.9); max_d = mean_b(max_b) * 1.1; state = state * 0.99 + new * 0.01
__global__ void depth_bins_update_kernel(const unsigned* __restrict__ minmax_bits, int batch, float opt_min_depth, float* __restrict__ min_state,
                                         float* __restrict__ max_state) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float smin = 0.f, smax = 0.f;
  for (int b = 0; b < batch; ++b) {        // torch's mean of B values: sum in index order, then one division
    smin = add_rn(smin, __uint_as_float(minmax_bits[2 * b]));
    smax = add_rn(smax, __uint_as_float(minmax_bits[2 * b + 1]));
  }
  const float mn = fmaxf(opt_min_depth, mul_rn(div_rn(smin, (float)batch), 0.9f));
  const float mx = mul_rn(div_rn(smax, (float)batch), 1.1f);
  *max_state = add_rn(mul_rn(*max_state, 0.99f), mul_rn(mx, 0.01f));
  *min_state = add_rn(mul_rn(*min_state, 0.99f), mul_rn(mn, 0.01f));
}

__global__ void minmax_init_kernel(unsigned* __restrict__ minmax_bits, int batch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < batch) minmax_bits[2 * i] = 0x7f800000u, minmax_bits[2 * i + 1] = 0u;
}

// "set missing images to 0 pose" (repdepth.py:502-505): pose[b] *= 0 where the pose features of item b sum to exactly zero.
// The reference asks the host once per batch item (`if feat.sum() == 0`); here the decision stays on the device.
// One CTA per item; fixed-order block reduction.  (An all-zero feature map sums to 0 in any order.)
__global__ void __launch_bounds__(256) zero_missing_poses_kernel(const float* __restrict__ feats, size_t per_item, float* __restrict__ pose,
                                                                 int pose_floats) {
  __shared__ float red[8];
  const int b = blockIdx.x;
  const float* f = feats + (size_t)b * per_item;
  float s = 0.f;
  bool any = false;
  for (size_t i = threadIdx.x; i < per_item; i += 256) {
    const float v = __ldg(f + i);
    s += v;
    any = any || v != 0.f;
  }
  // feat.sum() == 0 <=> (all zero) or an exact cancellation; the latter would need the reference's own summation order to
  // reproduce, and never occurs for features of a real image: decide on "any element non-zero" and on the sum
  s = warp_sum(s);
  const unsigned nz = __ballot_sync(0xffffffffu, any);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = (nz != 0u && s != 0.f) ? 1.f : 0.f;
  __syncthreads();
  bool keep = false;
  for (int k = 0; k < 8; ++k) keep = keep || red[k] != 0.f;
  if (!keep)
    for (int i = threadIdx.x; i < pose_floats; i += 256) pose[(size_t)b * pose_floats + i] *= 0.f;
}

extern "C" int ppea_matching_glue(const float* lowest_cost, const float* confidence, const float* mono_depth, float* lowest_cost_up,
                                  float* consistency_mask, void* minmax_scratch, int batch, int low_h, int low_w, int height, int width,
                                  void* stream) {
  if (!lowest_cost || !confidence || !mono_depth || !lowest_cost_up || !consistency_mask) return PPEA_E_NULL;
  if (batch <= 0 || batch > 65535 || low_h <= 0 || low_w <= 0 || height <= 0 || width <= 0 || (long long)height * width >= (1ll << 31)) return PPEA_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  if (minmax_scratch) minmax_init_kernel<<<ceil_div(batch, 128), 128, 0, st>>>((unsigned*)minmax_scratch, batch);
  const dim3 grid((unsigned)ceil_div(height * width, 256), (unsigned)batch);
  matching_glue_kernel<<<grid, 256, 0, st>>>(lowest_cost, confidence, mono_depth, lowest_cost_up, consistency_mask, (unsigned*)minmax_scratch,
                                             low_h, low_w, height, width);
  return (int)cudaGetLastError();
}

extern "C" int ppea_depth_bins_update(const void* minmax_scratch, int batch, float opt_min_depth, float* min_depth_state, float* max_depth_state,
                                      void* stream) {
  if (!minmax_scratch || !min_depth_state || !max_depth_state) return PPEA_E_NULL;
  if (batch <= 0) return PPEA_E_SHAPE;
  depth_bins_update_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const unsigned*)minmax_scratch, batch, opt_min_depth, min_depth_state, max_depth_state);
  return (int)cudaGetLastError();
}

extern "C" int ppea_zero_missing_poses(const float* pose_feats, size_t floats_per_item, float* pose, int pose_floats, int batch, void* stream) {
  if (!pose_feats || !pose) return PPEA_E_NULL;
  if (batch <= 0 || pose_floats <= 0 || floats_per_item == 0) return PPEA_E_SHAPE;
  zero_missing_poses_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(pose_feats, floats_per_item, pose, pose_floats);
  return (int)cudaGetLastError();
}

extern "C" int ppea_matching_mask(const float* mono_depth, const float* lowest_cost, uint8_t* mask, size_t count, void* stream) {
  if (!mono_depth || !lowest_cost || !mask) return PPEA_E_NULL;
  if (count == 0) return PPEA_OK;
  if (count > ((size_t)1 << 40)) return PPEA_E_SHAPE;
  matching_mask_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mono_depth, lowest_cost, mask, count);
  return (int)cudaGetLastError();
}

extern "C" int ppea_pose_to_matrix_forward(const float* axisangle, const float* translation, int invert, float* T, int batch,
                                           void* stream) {
  if (!axisangle || !translation || !T) return PPEA_E_NULL;
  if (batch <= 0) return PPEA_E_SHAPE;
  pose_to_matrix_forward_kernel<<<ceil_div(batch, 64), 64, 0, (cudaStream_t)stream>>>(axisangle, translation, invert, T, batch);
  return (int)cudaGetLastError();
}

extern "C" int ppea_pose_to_matrix_backward(const float* axisangle, const float* translation, int invert, const float* grad_T,
                                            float* grad_axisangle, float* grad_translation, int batch, void* stream) {
  if (!axisangle || !translation || !grad_T || !grad_axisangle || !grad_translation) return PPEA_E_NULL;
  if (batch <= 0) return PPEA_E_SHAPE;
  pose_to_matrix_backward_kernel<<<ceil_div(batch, 64), 64, 0, (cudaStream_t)stream>>>(axisangle, translation, invert, grad_T,
                                                                                      grad_axisangle, grad_translation, batch);
  return (int)cudaGetLastError();
}

}  // namespace ppea
