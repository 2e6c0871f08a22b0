// pyramid.cu -- input format on the device: the LANCZOS image pyramid of the dataset and the packed
// RGBx frame format.
//
// Reference: MonoDataset.preprocess (datasets/mono_dataset.py:96-112) resizes every colour frame on the CPU workers with
// `transforms.Resize((h // 2^i, w // 2^i), interpolation=Image.LANCZOS)` (:79-85), scale i from scale i - 1, then ToTensor.
// PIL's resampler (libImaging/Resample.c) is an exact integer algorithm for 8-bit images: per output coordinate a window
// [xmin, xmin + xmax) of taps, double-precision Lanczos-3 weights normalised to sum 1, converted to 22-bit fixed point
// (round half away from zero), accumulated in int32 from 2^21, shifted down and clipped to 0..255 -- horizontal pass first
// (into an 8-bit intermediate), then vertical.  Here the tap tables are built on the host by the same statements
// (ppea_lanczos_table: same libm sin, so the same integers) and the two passes run as two launches on uint8 planes:
// bit-identical to PIL, so the frames the loss reads are the reference's own, without the CPU resize and with a quarter of the
// host->device bytes.
//
// ppea_pack_rgbx_u8: planar (N,3,H,W) or interleaved (N,H,W,3) uint8 frames -> one 32-bit word r | g << 8 | b << 16 per
// pixel, the format the streaming loss kernel gathers from (vsl_stream.cu; its preparation launch builds the same words
// from float frames).
#include <cmath>

#include "vsl_common.cuh"

namespace ppea {

constexpr int kPrecisionBits = 32 - 8 - 2;      // Resample.c PRECISION_BITS

static double sinc_filter(double x) {
  if (x == 0.0) return 1.0;
  x = x * M_PI;
  return sin(x) / x;
}
static double lanczos_filter(double x) {      // Resample.c lanczos_filter, support 3
  if (-3.0 <= x && x < 3.0) return sinc_filter(x) * sinc_filter(x / 3);
  return 0.0;
}

__device__ __forceinline__ uint8_t clip8(int v) {      // Resample.c clip8: (v >> PRECISION_BITS) clamped to 0..255
  v >>= kPrecisionBits;
  return (uint8_t)min(255, max(0, v));
}

// one thread per output value; rows = n_planes * height
__global__ void resize_h_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t rows, int in_w, int out_w,
                                const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * out_w) return;
  const int xx = (int)(idx % out_w);
  const size_t row = idx / out_w;
  const int xmin = __ldg(bounds + 2 * xx), xmax = __ldg(bounds + 2 * xx + 1);
  const int* k = kk + (size_t)xx * ksize;
  const uint8_t* s = src + row * in_w + xmin;
  int ss = 1 << (kPrecisionBits - 1);
  for (int x = 0; x < xmax; ++x) ss += (int)__ldg(s + x) * __ldg(k + x);
  dst[idx] = clip8(ss);
}

__global__ void resize_v_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t n_planes, int in_h, int out_h, int w,
                                const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_planes * out_h * w) return;
  const int x = (int)(idx % w);
  const size_t r = idx / w;
  const int yy = (int)(r % out_h);
  const size_t plane = r / out_h;
  const int ymin = __ldg(bounds + 2 * yy), ymax = __ldg(bounds + 2 * yy + 1);
  const int* k = kk + (size_t)yy * ksize;
  const uint8_t* s = src + (plane * in_h + ymin) * w + x;
  int ss = 1 << (kPrecisionBits - 1);
  for (int y = 0; y < ymax; ++y) ss += (int)__ldg(s + (size_t)y * w) * __ldg(k + y);
  dst[idx] = clip8(ss);
}

__global__ void pack_rgbx_kernel(const uint8_t* __restrict__ src, uint32_t* __restrict__ dst, size_t n_images, size_t plane, int interleaved) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_images * plane) return;
  unsigned r, g, b;
  if (interleaved) {
    const uint8_t* p = src + idx * 3;
    r = __ldg(p), g = __ldg(p + 1), b = __ldg(p + 2);
  } else {
    const size_t n = idx / plane, o = idx % plane;
    const uint8_t* p = src + n * 3 * plane + o;
    r = __ldg(p), g = __ldg(p + plane), b = __ldg(p + 2 * plane);
  }
  dst[idx] = r | (g << 8) | (b << 16);
}

}  // namespace ppea

using namespace ppea;

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for the whole-image box (in0 = 0, in1 = in_size), Lanczos-3.
extern "C" int ppea_lanczos_ksize(int in_size, int out_size) {
  if (in_size <= 0 || out_size <= 0) return PPEA_E_SHAPE;
  double filterscale = (double)in_size / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 3.0 * filterscale;
  return (int)ceil(support) * 2 + 1;
}

extern "C" int ppea_lanczos_table(int in_size, int out_size, int* bounds, int* coeffs) {
  if (!bounds || !coeffs) return PPEA_E_NULL;
  const int ksize = ppea_lanczos_ksize(in_size, out_size);
  if (ksize < 0) return ksize;
  const double scale = (double)in_size / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 3.0 * filterscale;
  const double ss = 1.0 / filterscale;
  double* k = new double[ksize];
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    int x;
    for (x = 0; x < xmax; x++) {
      const double w = lanczos_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (x = 0; x < xmax; x++)
      if (ww != 0.0) k[x] /= ww;
    for (; x < ksize; x++) k[x] = 0;
    bounds[xx * 2 + 0] = xmin;
    bounds[xx * 2 + 1] = xmax;
    int* kk = coeffs + (size_t)xx * ksize;
    for (x = 0; x < ksize; x++) {      // normalize_coeffs_8bpc
      if (k[x] < 0)
        kk[x] = (int)(-0.5 + k[x] * (1 << kPrecisionBits));
      else
        kk[x] = (int)(0.5 + k[x] * (1 << kPrecisionBits));
    }
  }
  delete[] k;
  return ksize;
}

// src (n_planes, in_h, in_w) uint8 -> dst (n_planes, out_h, out_w) uint8; tmp holds n_planes * in_h * out_w bytes (the
// horizontally resized intermediate; unused when only one direction changes).  The tables are DEVICE copies of what
// ppea_lanczos_table returned for (in_w, out_w) and (in_h, out_h); a direction whose size does not change is skipped, as PIL does.
extern "C" int ppea_resize_lanczos_u8(const uint8_t* src, uint8_t* dst, uint8_t* tmp, size_t n_planes, int in_h, int in_w, int out_h,
                                      int out_w, const int* bounds_x, const int* coeffs_x, int ksize_x, const int* bounds_y,
                                      const int* coeffs_y, int ksize_y, void* stream) {
  if (!src || !dst) return PPEA_E_NULL;
  if (n_planes == 0 || in_h <= 0 || in_w <= 0 || out_h <= 0 || out_w <= 0) return PPEA_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const bool need_h = out_w != in_w, need_v = out_h != in_h;
  if ((need_h && (!bounds_x || !coeffs_x || ksize_x <= 0)) || (need_v && (!bounds_y || !coeffs_y || ksize_y <= 0))) return PPEA_E_NULL;
  if (need_h && need_v && !tmp) return PPEA_E_NULL;
  if (!need_h && !need_v) return (int)cudaMemcpyAsync(dst, src, n_planes * in_h * in_w, cudaMemcpyDeviceToDevice, st);
  const uint8_t* cur = src;
  if (need_h) {
    uint8_t* out = need_v ? tmp : dst;
    const size_t n = n_planes * in_h * out_w;
    resize_h_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cur, out, n_planes * in_h, in_w, out_w, bounds_x, coeffs_x, ksize_x);
    cur = out;
  }
  if (need_v) {
    const size_t n = n_planes * out_h * out_w;
    resize_v_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cur, dst, n_planes, in_h, out_h, out_w, bounds_y, coeffs_y, ksize_y);
  }
  return (int)cudaGetLastError();
}

extern "C" int ppea_pack_rgbx_u8(const uint8_t* src, uint32_t* dst, size_t n_images, int height, int width, int interleaved, void* stream) {
  if (!src || !dst) return PPEA_E_NULL;
  if (n_images == 0 || height <= 0 || width <= 0) return PPEA_E_SHAPE;
  const size_t n = n_images * height * width;
  pack_rgbx_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, n_images, (size_t)height * width, interleaved);
  return (int)cudaGetLastError();
}
