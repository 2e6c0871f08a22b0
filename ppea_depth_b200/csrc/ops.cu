// ops.cu -- the piecewise operators behind the reference's nn.Module / function
// API (layers.py SSIM, BackprojectDepth, Project3D, get_smooth_loss; F.grid_sample
// as called at trainer.py:911-914; Trainer.compute_reprojection_loss).  The training
// step never needs them -- the fused kernels in vsl_fwd.cu / vsl_bwd.cu cover the
// whole loss -- they exist for callers that use the modules one by one (the
// cost-volume encoders, debugging, `materialize_warps`).  Element-wise / small-stencil
// kernels reading through L1; same scalar arithmetic (vsl_math.cuh) as the fused path.
#include "vsl_common.cuh"

namespace ppea {
namespace {

constexpr int kT = 256;

struct WinSums {
  float Sx, Sxx, Sxy, Sy, Syy;
};

// reflect-padded 3x3 window sums at (qy, qx) of one plane pair, read from global memory
__device__ __forceinline__ WinSums window_sums(const float* __restrict__ X, const float* __restrict__ Y, int H, int W,
                                               int qy, int qx) {
  WinSums s = {0.f, 0.f, 0.f, 0.f, 0.f};
  const int xs[3] = {reflect_index(qx - 1, W), qx, reflect_index(qx + 1, W)};
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = reflect_index(qy + dy, H);
    const float* xr = X + (size_t)yy * W;
    const float* yr = Y + (size_t)yy * W;
    const float xa = __ldg(xr + xs[0]), xb = __ldg(xr + xs[1]), xc = __ldg(xr + xs[2]);
    const float ya = __ldg(yr + xs[0]), yb = __ldg(yr + xs[1]), yc = __ldg(yr + xs[2]);
    const float hx = xa + xb + xc, hy = ya + yb + yc;
    const float hxx = xa * xa + xb * xb + xc * xc, hyy = ya * ya + yb * yb + yc * yc, hxy = xa * ya + xb * yb + xc * yc;
    s.Sx = dy == -1 ? hx : s.Sx + hx;
    s.Sy = dy == -1 ? hy : s.Sy + hy;
    s.Sxx = dy == -1 ? hxx : s.Sxx + hxx;
    s.Syy = dy == -1 ? hyy : s.Syy + hyy;
    s.Sxy = dy == -1 ? hxy : s.Sxy + hxy;
  }
  return s;
}

// ---------------------------------------------------------------- SSIM (layers.py:243-257)
__global__ void __launch_bounds__(kT) ssim_forward_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                          float* __restrict__ out, int n_planes, int H, int W) {
  const size_t plane = (size_t)H * W;
  const size_t idx = (size_t)blockIdx.x * kT + threadIdx.x;
  if (idx >= plane * n_planes) return;
  const size_t pl = idx / plane;
  const int r = (int)(idx - pl * plane), qy = r / W, qx = r - qy * W;
  const WinSums s = window_sums(x + pl * plane, y + pl * plane, H, W, qy, qx);
  out[idx] = ssim_from_sums<float>(s.Sx, s.Sxx, s.Sxy, ssim_y_stats<float>(s.Sy, s.Syy));
}

// gradient of sum_q g(q) * SSIM(q) wrt x(p) and y(p): gather over the (up to 9) windows containing p,
// with the reflection multiplicity of border windows.  SSIM is symmetric in (x, y), so the y-adjoint
// is the x-adjoint with the roles swapped.
template <bool WANT_Y>
__device__ __forceinline__ void ssim_adjoint_gather(const float* __restrict__ X, const float* __restrict__ Y,
                                                    const float* __restrict__ G, float gscale, int H, int W, int py, int px,
                                                    float& gx_out, float& gy_out) {
  const float xp = __ldg(X + (size_t)py * W + px), yp = __ldg(Y + (size_t)py * W + px);
  float gx = 0.f, gy = 0.f;
  for (int dy = -1; dy <= 1; ++dy) {
    const int qy = py + dy;
    if (qy < 0 || qy >= H) continue;
    const float my = ((qy == 0 && py == 1) || (qy == H - 1 && py == H - 2)) ? 2.f : 1.f;
    for (int dx = -1; dx <= 1; ++dx) {
      const int qx = px + dx;
      if (qx < 0 || qx >= W) continue;
      const float g = __ldg(G + (size_t)qy * W + qx) * gscale;
      if (g == 0.f) continue;
      const float m = my * (((qx == 0 && px == 1) || (qx == W - 1 && px == W - 2)) ? 2.f : 1.f);
      const WinSums s = window_sums(X, Y, H, W, qy, qx);
      const SsimAdjT<float> ax = ssim_adjoint<float>(s.Sx, s.Sxx, s.Sxy, ssim_y_stats<float>(s.Sy, s.Syy), g);
      gx += m * (ax.cA + ax.cB * xp + ax.cC * yp);
      if (WANT_Y) {
        const SsimAdjT<float> ay = ssim_adjoint<float>(s.Sy, s.Syy, s.Sxy, ssim_y_stats<float>(s.Sx, s.Sxx), g);
        gy += m * (ay.cA + ay.cB * yp + ay.cC * xp);
      }
    }
  }
  gx_out = gx;
  gy_out = gy;
}

__global__ void __launch_bounds__(kT) ssim_backward_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                           const float* __restrict__ go, float* __restrict__ gx,
                                                           float* __restrict__ gy, int n_planes, int H, int W) {
  const size_t plane = (size_t)H * W;
  const size_t idx = (size_t)blockIdx.x * kT + threadIdx.x;
  if (idx >= plane * n_planes) return;
  const size_t pl = idx / plane;
  const int r = (int)(idx - pl * plane), py = r / W, px = r - py * W;
  float a, b;
  if (gy) {
    ssim_adjoint_gather<true>(x + pl * plane, y + pl * plane, go + pl * plane, 1.f, H, W, py, px, a, b);
    gy[idx] = b;
  } else {
    ssim_adjoint_gather<false>(x + pl * plane, y + pl * plane, go + pl * plane, 1.f, H, W, py, px, a, b);
  }
  if (gx) gx[idx] = a;
}

// ---------------------------------------------------------------- compute_reprojection_loss (trainer.py:995-1007)
__global__ void __launch_bounds__(kT) reprojection_forward_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                                                  float* __restrict__ out, int B, int H, int W, int no_ssim) {
  const size_t plane = (size_t)H * W;
  const size_t idx = (size_t)blockIdx.x * kT + threadIdx.x;
  if (idx >= plane * B) return;
  const size_t b = idx / plane;
  const int r = (int)(idx - b * plane), qy = r / W, qx = r - qy * W;
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* X = pred + (b * 3 + c) * plane;
    const float* Y = tgt + (b * 3 + c) * plane;
    const float l1 = fabsf(__ldg(Y + r) - __ldg(X + r));
    if (no_ssim) {
      acc += l1 * (1.f / 3.f);
    } else {
      const WinSums s = window_sums(X, Y, H, W, qy, qx);
      acc += PPEA_W_SSIM * ssim_from_sums<float>(s.Sx, s.Sxx, s.Sxy, ssim_y_stats<float>(s.Sy, s.Syy)) + PPEA_W_L1 * l1;
    }
  }
  out[idx] = acc;
}

__global__ void __launch_bounds__(kT) reprojection_backward_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                                                   const float* __restrict__ go, float* __restrict__ gpred,
                                                                   int B, int H, int W, int no_ssim) {
  const size_t plane = (size_t)H * W;
  const size_t idx = (size_t)blockIdx.x * kT + threadIdx.x;
  if (idx >= plane * B * 3) return;
  const size_t pl = idx / plane, b = pl / 3;
  const int r = (int)(idx - pl * plane), py = r / W, px = r - py * W;
  const float* X = pred + pl * plane;
  const float* Y = tgt + pl * plane;
  const float* G = go + b * plane;
  const float d = __ldg(Y + r) - __ldg(X + r);
  const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  float g = -__ldg(G + r) * (no_ssim ? (1.f / 3.f) : PPEA_W_L1) * sgn;
  if (!no_ssim) {
    float a, unused;
    ssim_adjoint_gather<false>(X, Y, G, PPEA_W_SSIM, H, W, py, px, a, unused);
    g += a;
  }
  gpred[idx] = g;
}

// ---------------------------------------------------------------- BackprojectDepth (layers.py:163-168)
__global__ void __launch_bounds__(kT) backproject_forward_kernel(const float* __restrict__ depth, const float* __restrict__ inv_K,
                                                                 float* __restrict__ cam, int B, int H, int W) {
  const size_t plane = (size_t)H * W;
  const size_t idx = (size_t)blockIdx.x * kT + threadIdx.x;
  if (idx >= plane * B) return;
  const size_t b = idx / plane;
  const int r = (int)(idx - b * plane), y = r / W, x = r - y * W;
  float iK[9], ray[3];
#pragma unroll
  for (int e = 0; e < 9; ++e) iK[e] = __ldg(inv_K + b * 16 + (e / 3) * 4 + (e % 3));
  pixel_ray(iK, (float)x, (float)y, ray);
  const float d = depth[idx];
  float* o = cam + b * 4 * plane + r;
  o[0] = mul_rn(d, ray[0]);
  o[plane] = mul_rn(d, ray[1]);
  o[2 * plane] = mul_rn(d, ray[2]);
  o[3 * plane] = 1.f;
}

__global__ void __launch_bounds__(kT) backproject_backward_kernel(const float* __restrict__ gcam, const float* __restrict__ inv_K,
                                                                  float* __restrict__ gdepth, int B, int H, int W) {
  const size_t plane = (size_t)H * W;
  const size_t idx = (size_t)blockIdx.x * kT + threadIdx.x;
  if (idx >= plane * B) return;
  const size_t b = idx / plane;
  const int r = (int)(idx - b * plane), y = r / W, x = r - y * W;
  float iK[9], ray[3];
#pragma unroll
  for (int e = 0; e < 9; ++e) iK[e] = __ldg(inv_K + b * 16 + (e / 3) * 4 + (e % 3));
  pixel_ray(iK, (float)x, (float)y, ray);
  const float* g = gcam + b * 4 * plane + r;
  gdepth[idx] = g[0] * ray[0] + g[plane] * ray[1] + g[2 * plane] * ray[2];
}

// ---------------------------------------------------------------- Project3D (layers.py:184-199)
__device__ __forceinline__ void load_P(const float* __restrict__ K, const float* __restrict__ T, size_t b, float* P) {
  float k[16], t[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    k[e] = __ldg(K + b * 16 + e);
    t[e] = __ldg(T + b * 16 + e);
  }
  compose_P(k, t, P);
}

__global__ void __launch_bounds__(kT) project3d_forward_kernel(const float* __restrict__ pts, const float* __restrict__ K,
                                                               const float* __restrict__ T, float* __restrict__ pix,
                                                               float* __restrict__ zout, int B, int H, int W, float eps) {
  const size_t plane = (size_t)H * W;
  const size_t idx = (size_t)blockIdx.x * kT + threadIdx.x;
  if (idx >= plane * B) return;
  const size_t b = idx / plane;
  const size_t r = idx - b * plane;
  float P[12];
  load_P(K, T, b, P);
  const float* p = pts + b * 4 * plane + r;
  const float X = p[0], Y = p[plane], Z = p[2 * plane], Wh = p[3 * plane];
  float c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float acc = mul_rn(P[i * 4 + 0], X);
    acc = fma_rn(P[i * 4 + 1], Y, acc);
    acc = fma_rn(P[i * 4 + 2], Z, acc);
    acc = fma_rn(P[i * 4 + 3], Wh, acc);
    c[i] = acc;
  }
  const float z = add_rn(c[2], eps);
  const float u = div_rn(c[0], z), v = div_rn(c[1], z);
  pix[idx * 2 + 0] = mul_rn(sub_rn(div_rn(u, (float)(W - 1)), 0.5f), 2.f);
  pix[idx * 2 + 1] = mul_rn(sub_rn(div_rn(v, (float)(H - 1)), 0.5f), 2.f);
  if (zout) zout[idx] = c[2];
}

// grad_points (B,4,HW) and per-block partials of dL/dP (fixed-order reduced by the finish kernel)
__global__ void __launch_bounds__(kT) project3d_backward_kernel(const float* __restrict__ pts, const float* __restrict__ K,
                                                                const float* __restrict__ T, const float* __restrict__ gpix,
                                                                const float* __restrict__ gz, float* __restrict__ gpts,
                                                                float* __restrict__ partials, int B, int H, int W, float eps,
                                                                int blocks_per_image) {
  __shared__ float red[12][kT / 32];
  const size_t plane = (size_t)H * W;
  const int b = blockIdx.x / blocks_per_image, blk = blockIdx.x - b * blocks_per_image;
  const size_t r = (size_t)blk * kT + threadIdx.x;
  float gP[12];
#pragma unroll
  for (int e = 0; e < 12; ++e) gP[e] = 0.f;
  if (r < plane) {
    float P[12];
    load_P(K, T, b, P);
    const float* p = pts + (size_t)b * 4 * plane + r;
    const float v4[4] = {p[0], p[plane], p[2 * plane], p[3 * plane]};
    float c[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float acc = mul_rn(P[i * 4 + 0], v4[0]);
      acc = fma_rn(P[i * 4 + 1], v4[1], acc);
      acc = fma_rn(P[i * 4 + 2], v4[2], acc);
      acc = fma_rn(P[i * 4 + 3], v4[3], acc);
      c[i] = acc;
    }
    const float z = c[2] + eps, inv_z = 1.f / z;
    const size_t idx = (size_t)b * plane + r;
    const float gu = gpix[idx * 2 + 0] * (2.f / (float)(W - 1)), gv = gpix[idx * 2 + 1] * (2.f / (float)(H - 1));
    float gc[3] = {gu * inv_z, gv * inv_z, -(gu * c[0] + gv * c[1]) * inv_z * inv_z};
    if (gz) gc[2] += gz[idx];
    float* go = gpts + (size_t)b * 4 * plane + r;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      go[j * plane] = P[j] * gc[0] + P[4 + j] * gc[1] + P[8 + j] * gc[2];
#pragma unroll
      for (int i = 0; i < 3; ++i) gP[i * 4 + j] = gc[i] * v4[j];
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int e = 0; e < 12; ++e) {
    const float v = warp_sum(gP[e]);
    if (lane == 0) red[e][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    float t = 0.f;
    for (int w = 0; w < kT / 32; ++w) t += red[threadIdx.x][w];
    partials[(size_t)blockIdx.x * 12 + threadIdx.x] = t;
  }
}

// one warp per image: lane l sums the blocks l, l + 32, ... of each of the 12 entries in double, then a fixed xor tree (fixed order of
// addition => bit-reproducible; a single thread per entry walking every block was one dependent DRAM round trip after another)
__global__ void __launch_bounds__(32) project3d_finish_kernel(const float* __restrict__ K, const float* __restrict__ partials,
                                                              float* __restrict__ gT, int blocks_per_image) {
  __shared__ double gP[12];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* p = partials + (size_t)b * blocks_per_image * 12;
#pragma unroll 1
  for (int e = 0; e < 12; ++e) {
    double t = 0;
    for (int i = tid; i < blocks_per_image; i += 32) t += (double)__ldg(p + (size_t)i * 12 + e);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (tid == 0) gP[e] = t;
  }
  __syncwarp();
  if (tid < 16) {
    const int i = tid / 4, j = tid % 4;
    double t = 0;
    for (int r = 0; r < 3; ++r) t += (double)K[b * 16 + r * 4 + i] * gP[r * 4 + j];
    gT[b * 16 + tid] = (float)t;
  }
}

// ---------------------------------------------------------------- grid_sample (bilinear, border, align_corners=True)
struct GridPos {
  float ix, iy, mx, my;
};
__device__ __forceinline__ GridPos grid_unnormalize(float gx, float gy, int W, int H) {
  // GridSampler.h grid_sampler_unnormalize + clip_coordinates(_set_grad)
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  const float fx = mul_rn(add_rn(gx, 1.f), 0.5f * wm1), fy = mul_rn(add_rn(gy, 1.f), 0.5f * hm1);
  GridPos g;
  g.mx = (fx > 0.f && fx < wm1) ? 0.5f * wm1 : 0.f;
  g.my = (fy > 0.f && fy < hm1) ? 0.5f * hm1 : 0.f;
  g.ix = clip_coord(fx, coord_max(W));
  g.iy = clip_coord(fy, coord_max(H));
  return g;
}

__global__ void __launch_bounds__(kT) warp_forward_kernel(const float* __restrict__ src, const float* __restrict__ grid,
                                                          float* __restrict__ out, int B, int C, int H, int W, int oh, int ow) {
  const size_t oplane = (size_t)oh * ow, plane = (size_t)H * W;
  const size_t idx = (size_t)blockIdx.x * kT + threadIdx.x;
  if (idx >= oplane * B) return;
  const size_t b = idx / oplane, r = idx - b * oplane;
  const GridPos g = grid_unnormalize(grid[idx * 2], grid[idx * 2 + 1], W, H);
  const Bilin bl = bilin_setup(g.ix, g.iy, W);
  for (int c = 0; c < C; ++c) {
    const float* S = src + (b * C + c) * plane;
    out[(b * C + c) * oplane + r] = bilin_value(bl, __ldg(S + bl.o00), __ldg(S + bl.o00 + 1), __ldg(S + bl.o00 + W), __ldg(S + bl.o00 + W + 1));
  }
}

__global__ void __launch_bounds__(kT) warp_backward_kernel(const float* __restrict__ src, const float* __restrict__ grid,
                                                           const float* __restrict__ go, float* __restrict__ ggrid, int B, int C,
                                                           int H, int W, int oh, int ow) {
  const size_t oplane = (size_t)oh * ow, plane = (size_t)H * W;
  const size_t idx = (size_t)blockIdx.x * kT + threadIdx.x;
  if (idx >= oplane * B) return;
  const size_t b = idx / oplane, r = idx - b * oplane;
  const GridPos g = grid_unnormalize(grid[idx * 2], grid[idx * 2 + 1], W, H);
  const Bilin bl = bilin_setup(g.ix, g.iy, W);
  float gx = 0.f, gy = 0.f;
  for (int c = 0; c < C; ++c) {
    const float* S = src + (b * C + c) * plane;
    const float nw = __ldg(S + bl.o00), ne = __ldg(S + bl.o00 + 1), sw = __ldg(S + bl.o00 + W), se = __ldg(S + bl.o00 + W + 1);
    const float gv = go[(b * C + c) * oplane + r];
    gx += gv * bilin_ddx(bl, nw, ne, sw, se);
    gy += gv * bilin_ddy(bl, nw, ne, sw, se);
  }
  ggrid[idx * 2] = gx * g.mx;
  ggrid[idx * 2 + 1] = gy * g.my;
}

// ---------------------------------------------------------------- get_smooth_loss (layers.py:210-223)
__device__ __forceinline__ float edge_w(const float* __restrict__ img, size_t plane, size_t i, size_t j) {
  float g = fabsf(__ldg(img + i) - __ldg(img + j));
  g += fabsf(__ldg(img + plane + i) - __ldg(img + plane + j));
  g += fabsf(__ldg(img + 2 * plane + i) - __ldg(img + 2 * plane + j));
  return __expf(-g * (1.f / 3.f));
}

__global__ void __launch_bounds__(kT) smooth_op_forward_kernel(const float* __restrict__ disp, const float* __restrict__ img,
                                                               float* __restrict__ partials, int B, int H, int W) {
  __shared__ float red[2][kT / 32];
  const size_t plane = (size_t)H * W, total = plane * B;
  float sx = 0.f, sy = 0.f;
  for (size_t idx = (size_t)blockIdx.x * kT + threadIdx.x; idx < total; idx += (size_t)gridDim.x * kT) {
    const size_t b = idx / plane, r = idx - b * plane;
    const int y = (int)(r / W), x = (int)(r - (size_t)y * W);
    const float* d = disp + b * plane;
    const float* im = img + b * 3 * plane;
    if (x + 1 < W) sx += fabsf(d[r] - d[r + 1]) * edge_w(im, plane, r, r + 1);
    if (y + 1 < H) sy += fabsf(d[r] - d[r + W]) * edge_w(im, plane, r, r + W);
  }
  sx = warp_sum(sx);
  sy = warp_sum(sy);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    red[0][wid] = sx;
    red[1][wid] = sy;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float t = 0.f;
    for (int w = 0; w < kT / 32; ++w) t += red[threadIdx.x][w];
    partials[blockIdx.x * 2 + threadIdx.x] = t;
  }
}

__global__ void __launch_bounds__(32) smooth_op_finish_kernel(const float* __restrict__ partials, int nblk, float* out, int B,
                                                              int H, int W) {
  // one warp: lane-strided sums in double + a fixed xor tree (fixed order of addition)
  double sx = 0, sy = 0;
  for (int i = threadIdx.x; i < nblk; i += 32) {
    sx += (double)__ldg(partials + i * 2);
    sy += (double)__ldg(partials + i * 2 + 1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
  }
  if (threadIdx.x == 0) out[0] = (float)(sx / ((double)B * H * (W - 1))) + (float)(sy / ((double)B * (H - 1) * W));
}

__global__ void __launch_bounds__(kT) smooth_op_backward_kernel(const float* __restrict__ disp, const float* __restrict__ img,
                                                                const float* __restrict__ gs, float* __restrict__ gd, int B, int H,
                                                                int W) {
  const size_t plane = (size_t)H * W;
  const size_t idx = (size_t)blockIdx.x * kT + threadIdx.x;
  if (idx >= plane * B) return;
  const size_t b = idx / plane, r = idx - b * plane;
  const int y = (int)(r / W), x = (int)(r - (size_t)y * W);
  const float* d = disp + b * plane;
  const float* im = img + b * 3 * plane;
  const float g = gs[0];
  const float gx = g / ((float)B * H * (W - 1)), gy = g / ((float)B * (H - 1) * W);
  auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
  const float di = d[r];
  float acc = 0.f;
  if (x + 1 < W) acc += gx * sgn(di - d[r + 1]) * edge_w(im, plane, r, r + 1);
  if (x > 0) acc -= gx * sgn(d[r - 1] - di) * edge_w(im, plane, r - 1, r);
  if (y + 1 < H) acc += gy * sgn(di - d[r + W]) * edge_w(im, plane, r, r + W);
  if (y > 0) acc -= gy * sgn(d[r - W] - di) * edge_w(im, plane, r - W, r);
  gd[idx] = acc;
}

inline int blocks_for(size_t n) { return (int)((n + kT - 1) / kT); }
constexpr int kSmoothOpBlocks = 296;   // 2 x 148 SMs

inline int shape_ok(int a, int b, int c) {
  if (a <= 0 || b <= 0 || c <= 0) return 0;
  return (size_t)a * b * c < ((size_t)1 << 31);
}

}  // namespace
}  // namespace ppea

using namespace ppea;

#define PPEA_RET_LAST()                      \
  do {                                       \
    cudaError_t e__ = cudaGetLastError();    \
    return e__ == cudaSuccess ? PPEA_OK : (int)e__; \
  } while (0)

// ---------------------------------------------------------------------------
// uint8 -> float32 image expansion (ToTensor on the device): 16 pixels per thread, 128-bit accesses.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float u8_unit(unsigned k) { return __fdiv_rn((float)k, 255.f); }

__global__ void __launch_bounds__(256) images_u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t count,
                                                               int vec_ok) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t base = t * 16;
  if (base >= count) return;
  if (vec_ok && base + 16 <= count) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + base));
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
    float4* o = reinterpret_cast<float4*>(dst + base);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      o[i] = make_float4(u8_unit(w[i] & 0xffu), u8_unit((w[i] >> 8) & 0xffu), u8_unit((w[i] >> 16) & 0xffu), u8_unit(w[i] >> 24));
  } else {
    for (size_t i = base; i < count && i < base + 16; ++i) dst[i] = u8_unit(src[i]);
  }
}

extern "C" {

int ppea_ssim_forward(const float* x, const float* y, float* out, int n_planes, int height, int width, void* stream) {
  if (!x || !y || !out) return PPEA_E_NULL;
  if (!shape_ok(n_planes, height, width) || height < 2 || width < 2) return PPEA_E_SHAPE;
  ssim_forward_kernel<<<blocks_for((size_t)n_planes * height * width), kT, 0, (cudaStream_t)stream>>>(x, y, out, n_planes, height, width);
  PPEA_RET_LAST();
}

int ppea_ssim_backward(const float* x, const float* y, const float* grad_out, float* grad_x, float* grad_y, int n_planes,
                       int height, int width, void* stream) {
  if (!x || !y || !grad_out || (!grad_x && !grad_y)) return PPEA_E_NULL;
  if (!shape_ok(n_planes, height, width) || height < 2 || width < 2) return PPEA_E_SHAPE;
  ssim_backward_kernel<<<blocks_for((size_t)n_planes * height * width), kT, 0, (cudaStream_t)stream>>>(x, y, grad_out, grad_x, grad_y,
                                                                                                      n_planes, height, width);
  PPEA_RET_LAST();
}

int ppea_reprojection_forward(const float* pred, const float* target, float* out, int batch, int height, int width, int no_ssim,
                              void* stream) {
  if (!pred || !target || !out) return PPEA_E_NULL;
  if (!shape_ok(batch * 3, height, width) || height < 2 || width < 2) return PPEA_E_SHAPE;
  reprojection_forward_kernel<<<blocks_for((size_t)batch * height * width), kT, 0, (cudaStream_t)stream>>>(pred, target, out, batch,
                                                                                                         height, width, no_ssim);
  PPEA_RET_LAST();
}

int ppea_reprojection_backward(const float* pred, const float* target, const float* grad_out, float* grad_pred, int batch,
                               int height, int width, int no_ssim, void* stream) {
  if (!pred || !target || !grad_out || !grad_pred) return PPEA_E_NULL;
  if (!shape_ok(batch * 3, height, width) || height < 2 || width < 2) return PPEA_E_SHAPE;
  reprojection_backward_kernel<<<blocks_for((size_t)batch * 3 * height * width), kT, 0, (cudaStream_t)stream>>>(
      pred, target, grad_out, grad_pred, batch, height, width, no_ssim);
  PPEA_RET_LAST();
}

int ppea_backproject_forward(const float* depth, const float* inv_K, float* cam_points, int batch, int height, int width,
                             void* stream) {
  if (!depth || !inv_K || !cam_points) return PPEA_E_NULL;
  if (!shape_ok(batch * 4, height, width)) return PPEA_E_SHAPE;
  backproject_forward_kernel<<<blocks_for((size_t)batch * height * width), kT, 0, (cudaStream_t)stream>>>(depth, inv_K, cam_points,
                                                                                                        batch, height, width);
  PPEA_RET_LAST();
}

int ppea_backproject_backward(const float* grad_cam, const float* inv_K, float* grad_depth, int batch, int height, int width,
                              void* stream) {
  if (!grad_cam || !inv_K || !grad_depth) return PPEA_E_NULL;
  if (!shape_ok(batch * 4, height, width)) return PPEA_E_SHAPE;
  backproject_backward_kernel<<<blocks_for((size_t)batch * height * width), kT, 0, (cudaStream_t)stream>>>(grad_cam, inv_K,
                                                                                                         grad_depth, batch, height, width);
  PPEA_RET_LAST();
}

int ppea_project3d_forward(const float* points, const float* K, const float* T, float* pix, float* z_or_null, int batch,
                           int height, int width, float eps, void* stream) {
  if (!points || !K || !T || !pix) return PPEA_E_NULL;
  if (!shape_ok(batch * 4, height, width) || height < 2 || width < 2) return PPEA_E_SHAPE;
  project3d_forward_kernel<<<blocks_for((size_t)batch * height * width), kT, 0, (cudaStream_t)stream>>>(points, K, T, pix, z_or_null,
                                                                                                      batch, height, width, eps);
  PPEA_RET_LAST();
}

size_t ppea_project3d_partials_bytes(int batch, int height, int width) {
  if (!shape_ok(batch, height, width)) return 0;
  return (size_t)batch * blocks_for((size_t)height * width) * 12 * sizeof(float);
}

int ppea_project3d_backward(const float* points, const float* K, const float* T, const float* grad_pix,
                            const float* grad_z_or_null, float* grad_points, float* grad_T, void* partials, int batch,
                            int height, int width, float eps, void* stream) {
  if (!points || !K || !T || !grad_pix || !grad_points || !grad_T || !partials) return PPEA_E_NULL;
  if (!shape_ok(batch * 4, height, width) || height < 2 || width < 2) return PPEA_E_SHAPE;
  const int bpi = blocks_for((size_t)height * width);
  project3d_backward_kernel<<<batch * bpi, kT, 0, (cudaStream_t)stream>>>(points, K, T, grad_pix, grad_z_or_null, grad_points,
                                                                        (float*)partials, batch, height, width, eps, bpi);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  project3d_finish_kernel<<<batch, 32, 0, (cudaStream_t)stream>>>(K, (const float*)partials, grad_T, bpi);
  PPEA_RET_LAST();
}

int ppea_warp_forward(const float* src, const float* grid, float* out, int batch, int channels, int height, int width,
                      int out_h, int out_w, void* stream) {
  if (!src || !grid || !out) return PPEA_E_NULL;
  if (!shape_ok(batch * channels, height, width) || !shape_ok(batch * channels, out_h, out_w)) return PPEA_E_SHAPE;
  warp_forward_kernel<<<blocks_for((size_t)batch * out_h * out_w), kT, 0, (cudaStream_t)stream>>>(src, grid, out, batch, channels,
                                                                                                height, width, out_h, out_w);
  PPEA_RET_LAST();
}

int ppea_warp_backward(const float* src, const float* grid, const float* grad_out, float* grad_grid, int batch, int channels,
                       int height, int width, int out_h, int out_w, void* stream) {
  if (!src || !grid || !grad_out || !grad_grid) return PPEA_E_NULL;
  if (!shape_ok(batch * channels, height, width) || !shape_ok(batch * channels, out_h, out_w)) return PPEA_E_SHAPE;
  warp_backward_kernel<<<blocks_for((size_t)batch * out_h * out_w), kT, 0, (cudaStream_t)stream>>>(
      src, grid, grad_out, grad_grid, batch, channels, height, width, out_h, out_w);
  PPEA_RET_LAST();
}

size_t ppea_smooth_workspace_bytes(int batch, int height, int width) {
  if (!shape_ok(batch, height, width)) return 0;
  return (size_t)kSmoothOpBlocks * 2 * sizeof(float);
}

int ppea_smooth_forward(const float* disp, const float* img, float* out_scalar, void* workspace, int batch, int height,
                        int width, void* stream) {
  if (!disp || !img || !out_scalar || !workspace) return PPEA_E_NULL;
  if (!shape_ok(batch * 3, height, width) || height < 2 || width < 2) return PPEA_E_SHAPE;
  smooth_op_forward_kernel<<<kSmoothOpBlocks, kT, 0, (cudaStream_t)stream>>>(disp, img, (float*)workspace, batch, height, width);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  smooth_op_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const float*)workspace, kSmoothOpBlocks, out_scalar, batch, height, width);
  PPEA_RET_LAST();
}

int ppea_smooth_backward(const float* disp, const float* img, const float* grad_scalar, float* grad_disp, int batch,
                         int height, int width, void* stream) {
  if (!disp || !img || !grad_scalar || !grad_disp) return PPEA_E_NULL;
  if (!shape_ok(batch * 3, height, width) || height < 2 || width < 2) return PPEA_E_SHAPE;
  smooth_op_backward_kernel<<<blocks_for((size_t)batch * height * width), kT, 0, (cudaStream_t)stream>>>(disp, img, grad_scalar,
                                                                                                       grad_disp, batch, height, width);
  PPEA_RET_LAST();
}

int ppea_images_u8_to_f32(const uint8_t* src, float* dst, size_t count, void* stream) {
  if (!src || !dst) return PPEA_E_NULL;
  if (count == 0) return PPEA_OK;
  if (count > ((size_t)1 << 40)) return PPEA_E_SHAPE;
  if (reinterpret_cast<uintptr_t>(dst) & 3) return PPEA_E_ALIGN;
  const int vec_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) ? 1 : 0;
  const size_t threads = (count + 15) / 16;
  images_u8_to_f32_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, count, vec_ok);
  PPEA_RET_LAST();
}

}  // extern "C"
