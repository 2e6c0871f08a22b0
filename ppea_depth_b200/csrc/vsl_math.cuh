// vsl_math.cuh -- per-pixel arithmetic of the view-synthesis loss path.
//
// Scalar building blocks shared by every kernel in this directory.  They are
// written as host/device functions so that tests/emul/vsl_emul.cpp can run
// the very same arithmetic pixel-by-pixel on the CPU and pin it against the
// oracle where no GPU exists (test tooling only -- the product never runs it).
//
// Reference arithmetic being restated (file:line under /root/reference/ppeadepth):
//   bilinear upsample   trainer.py:886-887 -> ATen UpSample.h area_pixel_compute_source_index /
//                       compute_source_index_and_lambda (align_corners=False)
//   disp_to_depth       layers.py:14-23
//   BackprojectDepth    layers.py:163-168
//   Project3D           layers.py:184-199
//   grid_sample         trainer.py:911-914 -> ATen GridSampler.h (bilinear, border, align_corners=True)
//   SSIM                layers.py:243-257
//   reprojection loss   trainer.py:995-1007
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PPEA_HD __host__ __device__ __forceinline__
#else
#define PPEA_HD inline
#endif

namespace ppea {

// ---- rounding-controlled primitives: the projection chain follows the
// reference's op-by-op fp32 rounding (each torch op rounds once), so these
// must not be contracted into FMAs by the compiler.
PPEA_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;
  return r;
#endif
}
PPEA_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}
PPEA_HD float sub_rn(float a, float b) { return add_rn(a, -b); }
PPEA_HD float div_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b;
  return r;
#endif
}
PPEA_HD float fma_rn(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}
PPEA_HD float fast_div(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdividef(a, b);
#else
  return a / b;
#endif
}
PPEA_HD float clamp01(float v) {
#if defined(__CUDA_ARCH__)
  return __saturatef(v);
#else
  return v < 0.f ? 0.f : (v > 1.f ? 1.f : v);   // NaN -> NaN on host, 0 on device (never hit: d > 0)
#endif
}

PPEA_HD int reflect_index(int i, int n) {
  // nn.ReflectionPad2d(1) index map (layers.py:238): -1 -> 1, n -> n-2; further out is clamped
  // (only reached by cells of partial tiles whose results are discarded).
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  if (i < 0) i = 0;
  if (i >= n) i = n - 1;
  return i;
}

// ---------------------------------------------------------------- upsample
struct UpCoef {
  int i0, i1;
  float l0, l1;
};

PPEA_HD float up_scale(int in_size, int out_size) { return (float)in_size / (float)out_size; }

PPEA_HD UpCoef up_coef(int dst, int in_size, float scale) {
  // ATen area_pixel_compute_source_index(scale, dst, align_corners=false, cubic=false)
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  UpCoef c;
  c.i0 = (int)src;
  c.i1 = c.i0 + ((c.i0 < in_size - 1) ? 1 : 0);
  c.l1 = src - (float)c.i0;
  c.l0 = 1.f - c.l1;
  return c;
}

PPEA_HD float up_sample(const float* __restrict__ d, int w_s, const UpCoef& cy, const UpCoef& cx) {
  const float* r0 = d + (size_t)cy.i0 * w_s;
  const float* r1 = d + (size_t)cy.i1 * w_s;
  return cy.l0 * (cx.l0 * r0[cx.i0] + cx.l1 * r0[cx.i1]) + cy.l1 * (cx.l0 * r1[cx.i0] + cx.l1 * r1[cx.i1]);
}

// ---------------------------------------------------------------- depth
PPEA_HD float depth_from_disp(float disp, float lo, float range) {
  // layers.py:21-22: scaled = lo + range*disp (two torch ops => two roundings); depth = 1/scaled
  return div_rn(1.f, add_rn(lo, mul_rn(range, disp)));
}
// d depth / d disp = -range * depth^2
PPEA_HD float ddepth_ddisp(float depth, float range) { return -range * depth * depth; }

// ---------------------------------------------------------------- geometry
// P = (K @ T)[:3, :]  (layers.py:185), row-major 3x4.
PPEA_HD void compose_P(const float* __restrict__ K, const float* __restrict__ T, float* __restrict__ P) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) {
      // ATen's small-matrix bmm path (4*4*4 < 400 elements) accumulates separately
      // rounded products in k order; the large (., HW) products go through sgemm,
      // whose FMA chain pixel_ray / project_point follow.
      float acc = mul_rn(K[i * 4 + 0], T[0 * 4 + j]);
      acc = add_rn(acc, mul_rn(K[i * 4 + 1], T[1 * 4 + j]));
      acc = add_rn(acc, mul_rn(K[i * 4 + 2], T[2 * 4 + j]));
      acc = add_rn(acc, mul_rn(K[i * 4 + 3], T[3 * 4 + j]));
      P[i * 4 + j] = acc;
    }
}

// ray = inv_K[:3,:3] @ (x, y, 1)   (layers.py:164); iK is the 3x3 block, row-major.
PPEA_HD void pixel_ray(const float* __restrict__ iK, float x, float y, float* __restrict__ ray) {
  for (int j = 0; j < 3; ++j) {
    float acc = mul_rn(iK[j * 3 + 0], x);
    acc = fma_rn(iK[j * 3 + 1], y, acc);
    acc = fma_rn(iK[j * 3 + 2], 1.f, acc);
    ray[j] = acc;
  }
}

struct Proj {
  float ix, iy;   // source-image sampling position after un-normalise + border clip
  float gx, gy;   // normalised grid coordinates (the ("sample", f, s) tensor of the reference)
  float z;        // c2 + eps
  float u, v;     // pixel coordinates c0/z, c1/z
  float mx, my;   // d(ix)/d(u), d(iy)/d(v): 1 inside the image, 0 where grid_sample clips (GridSampler.h clip_coordinates_set_grad)
};

// Project3D (layers.py:185-194) followed by grid_sample's un-normalise + clip
// (GridSampler.h grid_sampler_unnormalize / clip_coordinates, align_corners=True, border).
PPEA_HD Proj project_point(const float* __restrict__ P, const float* __restrict__ cam, float eps, float wm1, float hm1) {
  float c[3];
  for (int i = 0; i < 3; ++i) {
    float acc = mul_rn(P[i * 4 + 0], cam[0]);
    acc = fma_rn(P[i * 4 + 1], cam[1], acc);
    acc = fma_rn(P[i * 4 + 2], cam[2], acc);
    acc = fma_rn(P[i * 4 + 3], 1.f, acc);
    c[i] = acc;
  }
  Proj o;
  o.z = add_rn(c[2], eps);
  o.u = div_rn(c[0], o.z);
  o.v = div_rn(c[1], o.z);
  o.gx = mul_rn(sub_rn(div_rn(o.u, wm1), 0.5f), 2.f);
  o.gy = mul_rn(sub_rn(div_rn(o.v, hm1), 0.5f), 2.f);
  float fx = mul_rn(add_rn(o.gx, 1.f), 0.5f * wm1);
  float fy = mul_rn(add_rn(o.gy, 1.f), 0.5f * hm1);
  o.mx = (fx > 0.f && fx < wm1) ? 1.f : 0.f;
  o.my = (fy > 0.f && fy < hm1) ? 1.f : 0.f;
  o.ix = fminf(wm1, fmaxf(fx, 0.f));
  o.iy = fminf(hm1, fmaxf(fy, 0.f));
  return o;
}

// ---------------------------------------------------------------- bilinear gather
struct Bilin {
  int o00, o01, o10, o11;   // element offsets of the nw, ne, sw, se corners inside one channel plane
  float wnw, wne, wsw, wse; // corner weights
  float tx, ty;             // fractional position
};

PPEA_HD Bilin bilin_setup(float ix, float iy, int W, int H) {
  float x0f = floorf(ix), y0f = floorf(iy);
  Bilin b;
  b.tx = ix - x0f;
  b.ty = iy - y0f;
  float ex = 1.f - b.tx, ey = 1.f - b.ty;
  b.wnw = ey * ex;
  b.wne = ey * b.tx;
  b.wsw = b.ty * ex;
  b.wse = b.ty * b.tx;
  int x0 = (int)x0f, y0 = (int)y0f;
  // ix <= W-1 after the clip, so the +1 corners leave the image only when the
  // weight on them is exactly 0; clamping keeps the loads in bounds.
  int x1 = x0 + 1 < W ? x0 + 1 : W - 1;
  int y1 = y0 + 1 < H ? y0 + 1 : H - 1;
  b.o00 = y0 * W + x0;
  b.o01 = y0 * W + x1;
  b.o10 = y1 * W + x0;
  b.o11 = y1 * W + x1;
  return b;
}

PPEA_HD float bilin_value(const Bilin& b, float nw, float ne, float sw, float se) {
  // ATen's CPU grid_sampler accumulates the four corners as an FMA chain (checked bitwise)
  return fma_rn(se, b.wse, fma_rn(sw, b.wsw, fma_rn(ne, b.wne, mul_rn(nw, b.wnw))));
}
// d value / d ix and d value / d iy   (GridSampler grid_sampler_2d_backward)
PPEA_HD float bilin_ddx(const Bilin& b, float nw, float ne, float sw, float se) {
  return (ne - nw) * (1.f - b.ty) + (se - sw) * b.ty;
}
PPEA_HD float bilin_ddy(const Bilin& b, float nw, float ne, float sw, float se) {
  return (sw - nw) * (1.f - b.tx) + (se - ne) * b.tx;
}

// ---------------------------------------------------------------- SSIM
// layers.py:243-257 with every mean written as (window sum)/9 and the common
// factor 81^2 cancelled between numerator and denominator:
//   n/d = (2 Sx Sy + 81 C1)(18 Sxy - 2 Sx Sy + 81 C2) / ((Sx^2 + Sy^2 + 81 C1)(9 Sxx - Sx^2 + 9 Syy - Sy^2 + 81 C2))
// -- the same real number as the reference's expression, without divisions by 9.
#define PPEA_SSIM_K1 (81.0f * 1e-4f)
#define PPEA_SSIM_K2 (81.0f * 9e-4f)
#define PPEA_W_SSIM (0.85f / 3.0f)
#define PPEA_W_L1 (0.15f / 3.0f)

struct SsimY {  // per-window statistics of the target image (shared by every source)
  float s;      // Sy
  float d1;     // Sy^2 + 81 C1
  float v;      // 9 Syy - Sy^2   (81 x variance)
};

// The cancelling differences (9 Sxx - Sx^2, 18 Sxy - 2 Sx Sy) are formed FIRST, with a
// single rounding each (FMA), and the constants added afterwards: adding 81*C2 to a
// ~40-magnitude operand before the subtraction rounds it onto that operand's ulp grid
// and biases the whole loss by ~6e-6 relative (measured; see tests/test_emul.py).
PPEA_HD SsimY ssim_y_stats(float Sy, float Syy) {
  SsimY y;
  y.s = Sy;
  y.d1 = fma_rn(Sy, Sy, PPEA_SSIM_K1);
  y.v = fma_rn(9.f, Syy, -mul_rn(Sy, Sy));
  return y;
}

struct SsimTerms {
  float n1, n2, d1, d2;
};

PPEA_HD SsimTerms ssim_terms(float Sx, float Sxx, float Sxy, const SsimY& y) {
  SsimTerms t;
  const float a2 = 2.f * mul_rn(Sx, y.s);
  const float sx2 = mul_rn(Sx, Sx);
  t.n1 = add_rn(a2, PPEA_SSIM_K1);
  t.n2 = add_rn(fma_rn(18.f, Sxy, -a2), PPEA_SSIM_K2);
  t.d1 = add_rn(sx2, y.d1);
  t.d2 = add_rn(add_rn(fma_rn(9.f, Sxx, -sx2), y.v), PPEA_SSIM_K2);
  return t;
}

// SSIM dissimilarity clamp((1 - n/d)/2, 0, 1) from the 3x3 window sums
PPEA_HD float ssim_from_sums(float Sx, float Sxx, float Sxy, const SsimY& y) {
  const SsimTerms t = ssim_terms(Sx, Sxx, Sxy, y);
  const float R = fast_div(t.n1 * t.n2, t.d1 * t.d2);
  return clamp01(fma_rn(-0.5f, R, 0.5f));
}

// Adjoint of ssim_from_sums wrt the x-dependent window sums, scaled by `g`
// (upstream weight of this window's SSIM value):
//   d(g*S)/dx(p) = cA + cB*x(p) + cC*y(p)   for every tap p of the window
// (dSx/dx(p) = 1, dSxx/dx(p) = 2 x(p), dSxy/dx(p) = y(p)).
struct SsimAdj {
  float cA, cB, cC;
};

PPEA_HD SsimAdj ssim_adjoint(float Sx, float Sxx, float Sxy, const SsimY& y, float g) {
  const SsimTerms t = ssim_terms(Sx, Sxx, Sxy, y);
  const float inv_d = 1.f / (t.d1 * t.d2);
  const float R = t.n1 * t.n2 * inv_d;
  const float v = fma_rn(-0.5f, R, 0.5f);
  // torch.clamp backward passes the gradient where min <= v <= max (inclusive)
  const float k = (v >= 0.f && v <= 1.f) ? -0.5f * g * inv_d : 0.f;
  SsimAdj o;
  o.cA = k * (2.f * y.s * (t.n2 - t.n1) - 2.f * R * Sx * (t.d2 - t.d1));
  o.cB = k * (-18.f * R * t.d1);
  o.cC = k * (18.f * t.n1);
  return o;
}

// ---------------------------------------------------------------- selection (trainer.py:1076-1091)
struct Select {
  float r;      // reprojection loss after min over sources and the selec_reproj overrides
  int src;      // 0 / 1: source whose loss is propagated; 2: none
};

PPEA_HD Select select_source(float L0, float L1, float csum0, float csum1, bool selec_reproj) {
  Select s;
  // torch.min(dim=1) keeps the first index on ties
  if (L1 < L0) { s.r = L1; s.src = 1; } else { s.r = L0; s.src = 0; }
  if (selec_reproj) {
    bool dark0 = csum0 < 0.1f, dark1 = csum1 < 0.1f;
    if (dark0) { s.r = L1; s.src = 1; }
    if (dark1) { s.r = L0; s.src = 0; }
    if (dark0 && dark1) { s.r = 0.f; s.src = 2; }
  }
  return s;
}

}  // namespace ppea
