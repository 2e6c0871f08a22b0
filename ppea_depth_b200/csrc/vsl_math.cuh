// vsl_math.cuh -- per-pixel arithmetic of the view-synthesis loss path.
//
// Scalar / 2-wide building blocks shared by every kernel in this directory.  They are
// host/device functions so that tests/emul/vsl_emul.cpp can run the very same arithmetic
// pixel-by-pixel on the CPU and pin it against the oracle where no GPU exists (test tooling
// only -- the product never runs it).
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
//
// Numerical policy.  depth (an OUTPUT tensor of the path) follows the reference's op-by-op fp32
// rounding.  The sampling position is computed as  u = c0/(c2+eps)  directly: the reference's
// normalise (/(W-1), -0.5, *2) followed by grid_sample's un-normalise is the identity in real
// arithmetic and only adds ~2e-5 px of fp32 noise of its own, so no evaluation order can
// reproduce it bit for bit anyway; the contract (BASELINE.json) is per-pixel maps within fp32
// noise, selection bit-exact outside that margin, loss 1e-5, gradients 1e-4 -- tests measure it.
// The composition  (K T)[:3,:3] K^-1  is formed once per image in double.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PPEA_HD __host__ __device__ __forceinline__
#else
#define PPEA_HD inline
#endif

namespace ppea {

// ---- rounding-controlled scalar primitives (never contracted by the compiler)
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
PPEA_HD float rcp_fast(float a) {      // one MUFU.RCP (<= 1 ulp) on the device
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
#else
  return 1.f / a;
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
PPEA_HD float sign_of(float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }

// floor of 0 <= v < 2^22 without the conversion (XU) pipe: adding 2^23 with round-down leaves
// floor(v) in the low mantissa bits.
struct FloorIF {
  int i;
  float f;
};
PPEA_HD FloorIF floor_if(float v) {
  FloorIF o;
#if defined(__CUDA_ARCH__)
  const float t = __fadd_rd(v, 8388608.f);
  o.i = __float_as_int(t) & 0x7fffff;
  o.f = t - 8388608.f;
#else
  o.f = floorf(v);
  o.i = (int)o.f;
#endif
  return o;
}
PPEA_HD float int_to_float(int v) {   // 0 <= v < 2^23, exact, no I2F
#if defined(__CUDA_ARCH__)
  return __int_as_float(0x4b000000 | v) - 8388608.f;
#else
  return (float)v;
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

// ---------------------------------------------------------------- 2-wide fp32 (lanes = the two source frames)
// Blackwell issues packed fp32 pairs (FFMA2 / FADD2 / FMUL2, PTX *.f32x2) as ONE instruction: the
// photometric arithmetic of both sources advances in lockstep at half the issue cost.
struct __attribute__((aligned(8))) f2 {
  float x, y;
};
PPEA_HD f2 mk2(float a, float b) {
  f2 r;
  r.x = a;
  r.y = b;
  return r;
}
PPEA_HD f2 dup2(float a) { return mk2(a, a); }

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ unsigned long long f2_pack(f2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ f2 f2_unpack(unsigned long long r) {
  f2 a;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
  return a;
}
#endif

PPEA_HD f2 vfma(f2 a, f2 b, f2 c) {
#if defined(__CUDA_ARCH__)
  unsigned long long r;
  asm("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
  return f2_unpack(r);
#else
  return mk2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
PPEA_HD f2 vadd(f2 a, f2 b) {
#if defined(__CUDA_ARCH__)
  unsigned long long r;
  asm("add.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(r);
#else
  return mk2(add_rn(a.x, b.x), add_rn(a.y, b.y));
#endif
}
PPEA_HD f2 vmul(f2 a, f2 b) {
#if defined(__CUDA_ARCH__)
  unsigned long long r;
  asm("mul.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(r);
#else
  return mk2(mul_rn(a.x, b.x), mul_rn(a.y, b.y));
#endif
}
PPEA_HD f2 vneg(f2 a) { return mk2(-a.x, -a.y); }
PPEA_HD f2 vsub(f2 a, f2 b) { return vadd(a, vneg(b)); }
// scalar overloads so the same templates serve the CPU emulator / one-source callers
PPEA_HD float vfma(float a, float b, float c) { return fma_rn(a, b, c); }
PPEA_HD float vadd(float a, float b) { return add_rn(a, b); }
PPEA_HD float vmul(float a, float b) { return mul_rn(a, b); }
PPEA_HD float vsub(float a, float b) { return add_rn(a, -b); }
PPEA_HD float vneg(float a) { return -a; }
PPEA_HD void splat(float v, float& o) { o = v; }
PPEA_HD void splat(float v, f2& o) { o = dup2(v); }
template <class T>
PPEA_HD T vconst(float v) {
  T o;
  splat(v, o);
  return o;
}
PPEA_HD float vrcp(float a) { return rcp_fast(a); }
PPEA_HD f2 vrcp(f2 a) { return mk2(rcp_fast(a.x), rcp_fast(a.y)); }
PPEA_HD float vsat_half_minus_half(float R) { return clamp01(fma_rn(-0.5f, R, 0.5f)); }   // clamp((1-R)/2, 0, 1)
PPEA_HD f2 vsat_half_minus_half(f2 R) { return mk2(vsat_half_minus_half(R.x), vsat_half_minus_half(R.y)); }

// ---------------------------------------------------------------- upsample
struct UpCoef {
  int i0, i1;
  float l0, l1;
};

PPEA_HD float up_scale(int in_size, int out_size) { return (float)in_size / (float)out_size; }

PPEA_HD UpCoef up_coef(int dst, int in_size, float scale) {
  // ATen area_pixel_compute_source_index(scale, dst, align_corners=false, cubic=false)
  float src = scale * (int_to_float(dst) + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  const FloorIF fl = floor_if(src);
  UpCoef c;
  c.i0 = fl.i;
  c.i1 = c.i0 + ((c.i0 < in_size - 1) ? 1 : 0);
  c.l1 = src - fl.f;
  c.l0 = 1.f - c.l1;
  return c;
}

PPEA_HD float up_sample(const float* __restrict__ d, int w_s, const UpCoef& cy, const UpCoef& cx) {
  // 32-bit element offsets from one base pointer: one IMAD.WIDE.U32 per address on the device
  const unsigned r0 = (unsigned)cy.i0 * (unsigned)w_s, r1 = (unsigned)cy.i1 * (unsigned)w_s;
  // explicit roundings (mul, fma per lerp) so that every kernel that inlines this gets the same bits
  const float top = fma_rn(cx.l1, d[r0 + (unsigned)cx.i1], mul_rn(cx.l0, d[r0 + (unsigned)cx.i0]));
  const float bot = fma_rn(cx.l1, d[r1 + (unsigned)cx.i1], mul_rn(cx.l0, d[r1 + (unsigned)cx.i0]));
  return fma_rn(cy.l1, bot, mul_rn(cy.l0, top));
}

// ---------------------------------------------------------------- depth
PPEA_HD float depth_from_disp(float disp, float lo, float range) {
  // layers.py:21-22: scaled = lo + range*disp (two torch ops => two roundings); depth = 1/scaled
  return div_rn(1.f, add_rn(lo, mul_rn(range, disp)));
}
// The kernels' version: MUFU.RCP + one Newton step (<= 1 ulp from the correctly rounded quotient)
// instead of the ~11-instruction IEEE division sequence.
PPEA_HD float depth_from_disp_fast(float disp, float lo, float range) {
#if defined(__CUDA_ARCH__)
  const float x = add_rn(lo, mul_rn(range, disp));
  const float r = rcp_fast(x);
  return fma_rn(r, fma_rn(-x, r, 1.f), r);
#else
  return depth_from_disp(disp, lo, range);
#endif
}
// d depth / d disp = -range * depth^2
PPEA_HD float ddepth_ddisp(float depth, float range) { return -range * depth * depth; }

// ---------------------------------------------------------------- geometry
// For one image and one source:  c = (K T)[:3,:] (depth * K^-1 (x,y,1), 1)  (layers.py:164-166, 185-187)
//                                  = depth * (M (x,y,1)) + t,   M = (K T)[:3,:3] K^-1[:3,:3],  t = (K T)[:3,3].
struct Geom {
  float M[9];   // row-major 3x3
  float t[3];
};

// single entry e (0..11: M[0..8], t[0..2]) -- the kernels compute one entry per thread
PPEA_HD float geom_entry(const float* __restrict__ K, const float* __restrict__ T, const float* __restrict__ inv_K, int e) {
  const int i = e < 9 ? e / 3 : e - 9;
  double P[4];
  for (int j = 0; j < 4; ++j) {
    double acc = 0;
    for (int k = 0; k < 4; ++k) acc += (double)K[i * 4 + k] * (double)T[k * 4 + j];
    P[j] = acc;
  }
  if (e >= 9) return (float)P[3];
  const int j = e % 3;
  double acc = 0;
  for (int k = 0; k < 3; ++k) acc += P[k] * (double)inv_K[k * 4 + j];
  return (float)acc;
}
PPEA_HD void compose_geom(const float* __restrict__ K, const float* __restrict__ T, const float* __restrict__ inv_K,
                          Geom& g) {
  for (int e = 0; e < 9; ++e) g.M[e] = geom_entry(K, T, inv_K, e);
  for (int e = 0; e < 3; ++e) g.t[e] = geom_entry(K, T, inv_K, 9 + e);
}

// Sampling position of one pixel in one source.  A = M (x,y,1) (the caller forms it; the row part
// is hoisted out of inner loops), c = depth*A + t.
template <class T>
struct ProjT {
  T ix, iy;     // sampling position after the border clip (GridSampler.h clip_coordinates)
  T u, v;       // un-clipped pixel coordinates c0/z, c1/z
  T rz;         // 1/(c2 + eps)
};

PPEA_HD float clip_coord(float v, float vmax) { return fminf(vmax, fmaxf(v, 0.f)); }   // NaN -> 0 like ATen CUDA

PPEA_HD ProjT<float> project_fast(float depth, float A0, float A1, float A2, float t0, float t1, float t2, float eps,
                                  float wmax, float hmax) {
  ProjT<float> o;
  const float c0 = fma_rn(depth, A0, t0), c1 = fma_rn(depth, A1, t1), c2 = fma_rn(depth, A2, t2);
  o.rz = rcp_fast(add_rn(c2, eps));
  o.u = mul_rn(c0, o.rz);
  o.v = mul_rn(c1, o.rz);
  o.ix = clip_coord(o.u, wmax);
  o.iy = clip_coord(o.v, hmax);
  return o;
}
PPEA_HD ProjT<f2> project_fast(float depth, f2 A0, f2 A1, f2 A2, f2 t0, f2 t1, f2 t2, float eps, float wmax, float hmax) {
  ProjT<f2> o;
  const f2 d = dup2(depth);
  const f2 c0 = vfma(d, A0, t0), c1 = vfma(d, A1, t1), c2 = vfma(d, A2, t2);
  o.rz = mk2(rcp_fast(add_rn(c2.x, eps)), rcp_fast(add_rn(c2.y, eps)));
  o.u = vmul(c0, o.rz);
  o.v = vmul(c1, o.rz);
  o.ix = mk2(clip_coord(o.u.x, wmax), clip_coord(o.u.y, wmax));
  o.iy = mk2(clip_coord(o.v.x, hmax), clip_coord(o.v.y, hmax));
  return o;
}
// largest sampling coordinate: just below size-1, so that floor(ix)+1 is always a valid index
// (the weight of that neighbour is then ~1 instead of exactly 1 at ix == size-1: a 1-ulp change)
PPEA_HD float coord_max(int size) {
  const float m = (float)(size - 1);
  return m * (1.f - 5.9604645e-8f);
}
// d(ix)/d(u): 1 inside the image, 0 where grid_sample clips (GridSampler.h clip_coordinates_set_grad)
PPEA_HD float clip_mask(float u, float size_m1) { return (u > 0.f && u < size_m1) ? 1.f : 0.f; }

// ---------------------------------------------------------------- bilinear gather
struct Bilin {
  int o00;                  // element offset of the north-west corner inside one channel plane (ne = +1, sw = +W, se = +W+1)
  float wnw, wne, wsw, wse; // corner weights
  float tx, ty;             // fractional position
};

PPEA_HD Bilin bilin_setup(float ix, float iy, int W) {
  const FloorIF fx = floor_if(ix), fy = floor_if(iy);
  Bilin b;
  b.tx = ix - fx.f;
  b.ty = iy - fy.f;
  const float ex = 1.f - b.tx, ey = 1.f - b.ty;
  b.wnw = ey * ex;
  b.wne = ey * b.tx;
  b.wsw = b.ty * ex;
  b.wse = b.ty * b.tx;
  b.o00 = fy.i * W + fx.i;
  return b;
}

PPEA_HD float bilin_value(const Bilin& b, float nw, float ne, float sw, float se) {
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
// layers.py:243-257 with every mean written as (window sum)/9 and the common factor 81^2
// cancelled between numerator and denominator:
//   n/d = (2 Sx Sy + 81 C1)(18 Sxy - 2 Sx Sy + 81 C2) / ((Sx^2 + Sy^2 + 81 C1)(9 Sxx - Sx^2 + 9 Syy - Sy^2 + 81 C2))
// -- the same real number as the reference's expression, without divisions by 9.
#define PPEA_SSIM_K1 (81.0f * 1e-4f)
#define PPEA_SSIM_K2 (81.0f * 9e-4f)
#define PPEA_W_SSIM (0.85f / 3.0f)
#define PPEA_W_L1 (0.15f / 3.0f)

// horizontal 3-tap sums of one window row (T = float: one source, T = f2: both sources)
template <class T>
PPEA_HD void row_sums_x(T x0, T x1, T x2, T y0, T y1, T y2, T& hx, T& hxx, T& hxy) {
  hx = vadd(vadd(x0, x1), x2);
  hxx = vfma(x2, x2, vfma(x1, x1, vmul(x0, x0)));
  hxy = vfma(x2, y2, vfma(x1, y1, vmul(x0, y0)));
}
template <class T>
PPEA_HD void row_sums_y(T y0, T y1, T y2, T& hy, T& hyy) {
  hy = vadd(vadd(y0, y1), y2);
  hyy = vfma(y2, y2, vfma(y1, y1, vmul(y0, y0)));
}
template <class T>
PPEA_HD T sum3(T a, T b, T c) {
  return vadd(vadd(a, b), c);
}

template <class T>
struct SsimYT {   // per-window statistics of the target image (shared by every source)
  T s;            // Sy
  T d1;           // Sy^2 + 81 C1
  T v;            // 9 Syy - Sy^2   (81 x variance)
};

// The cancelling differences (9 Sxx - Sx^2, 18 Sxy - 2 Sx Sy) are formed FIRST, with a single
// rounding each (FMA), and the constants added afterwards: adding 81*C2 to a ~40-magnitude
// operand before the subtraction rounds it onto that operand's ulp grid and biases the whole
// loss by ~6e-6 relative (measured; see tests/test_emul.py).
template <class T>
PPEA_HD SsimYT<T> ssim_y_stats(T Sy, T Syy) {
  SsimYT<T> y;
  y.s = Sy;
  y.d1 = vfma(Sy, Sy, vconst<T>(PPEA_SSIM_K1));
  y.v = vfma(vconst<T>(9.f), Syy, vneg(vmul(Sy, Sy)));
  return y;
}

template <class T>
struct SsimTermsT {
  T n1, n2, d1, d2;
};

template <class T>
PPEA_HD SsimTermsT<T> ssim_terms(T Sx, T Sxx, T Sxy, const SsimYT<T>& y) {
  SsimTermsT<T> t;
  const T a2 = vmul(vconst<T>(2.f), vmul(Sx, y.s));
  const T sx2 = vmul(Sx, Sx);
  t.n1 = vadd(a2, vconst<T>(PPEA_SSIM_K1));
  t.n2 = vadd(vfma(vconst<T>(18.f), Sxy, vneg(a2)), vconst<T>(PPEA_SSIM_K2));
  t.d1 = vadd(sx2, y.d1);
  t.d2 = vadd(vadd(vfma(vconst<T>(9.f), Sxx, vneg(sx2)), y.v), vconst<T>(PPEA_SSIM_K2));
  return t;
}

// SSIM dissimilarity clamp((1 - n/d)/2, 0, 1) from the 3x3 window sums
template <class T>
PPEA_HD T ssim_from_sums(T Sx, T Sxx, T Sxy, const SsimYT<T>& y) {
  const SsimTermsT<T> t = ssim_terms(Sx, Sxx, Sxy, y);
  const T R = vmul(vmul(t.n1, t.n2), vrcp(vmul(t.d1, t.d2)));
  return vsat_half_minus_half(R);
}

// Adjoint of ssim_from_sums wrt the x-dependent window sums, scaled by `g` (upstream weight of
// this window's SSIM value):   d(g*S)/dx(p) = cA + cB*x(p) + cC*y(p)   for every tap p of the window
// (dSx/dx(p) = 1, dSxx/dx(p) = 2 x(p), dSxy/dx(p) = y(p)).
template <class T>
struct SsimAdjT {
  T cA, cB, cC;
};

PPEA_HD float clamp_pass(float v, float k) { return (v >= 0.f && v <= 1.f) ? k : 0.f; }   // torch.clamp backward (inclusive)
PPEA_HD f2 clamp_pass(f2 v, f2 k) { return mk2(clamp_pass(v.x, k.x), clamp_pass(v.y, k.y)); }

template <class T>
PPEA_HD SsimAdjT<T> ssim_adjoint(T Sx, T Sxx, T Sxy, const SsimYT<T>& y, T g) {
  const SsimTermsT<T> t = ssim_terms(Sx, Sxx, Sxy, y);
  const T inv_d = vrcp(vmul(t.d1, t.d2));
  const T R = vmul(vmul(t.n1, t.n2), inv_d);
  const T v = vfma(vconst<T>(-0.5f), R, vconst<T>(0.5f));
  const T k = clamp_pass(v, vmul(vmul(vconst<T>(-0.5f), g), inv_d));
  SsimAdjT<T> o;
  // cA = k * (2 Sy (n2 - n1) - 2 R Sx (d2 - d1))
  const T p = vmul(vmul(vconst<T>(2.f), y.s), vsub(t.n2, t.n1));
  const T q = vmul(vmul(vmul(vconst<T>(2.f), R), Sx), vsub(t.d2, t.d1));
  o.cA = vmul(k, vsub(p, q));
  o.cB = vmul(k, vmul(vmul(vconst<T>(-18.f), R), t.d1));
  o.cC = vmul(k, vmul(vconst<T>(18.f), t.n1));
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


// ---------------------------------------------------------------- reference-order geometry (piecewise operators)
// BackprojectDepth / Project3D as stand-alone modules hand their results (camera points, the normalised
// sampling grid) to the caller, so they follow the reference's op-by-op fp32 rounding.
// P = (K @ T)[:3, :]  (layers.py:185), row-major 3x4.
PPEA_HD void compose_P(const float* __restrict__ K, const float* __restrict__ T, float* __restrict__ P) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) {
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

}  // namespace ppea
