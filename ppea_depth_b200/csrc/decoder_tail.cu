// decoder_tail.cu -- the disparity head of DepthDecoderV2.
//
// Reference: `self.outputs[("disp", 0)] = self.sigmoid(self.disp_convs[0](x))` (networks/depth_decoder_v2.py:123-129, :239)
// with Conv3x3 = nn.ReflectionPad2d(1) + nn.Conv2d(C, 1, 3) (layers.py:119-135), followed on the loss side by disp_to_depth
// (layers.py:14-23, trainer.py:888).  A convolution with ONE output channel is not a contraction worth a tensor core: it
// reads C values per output and is bound by the read of x.  In the reference it is a pad launch (a full padded copy of x),
// a cuDNN launch, a sigmoid launch and -- in autograd -- their three backwards; here
//
//   disp_head_forward_kernel     x (B,C,H,W), w (C,3,3), bias -> disp (B,1,H,W) [+ depth]            reads x once
//   disp_head_dx_kernel          g = grad_disp * disp * (1 - disp)  -> grad_x (B,C,H,W)              never reads x
//   disp_head_dw_kernel          x, g -> per-task partial sums of grad_w (C,3,3) and grad_bias       reads x once
//   disp_head_dw_finish_kernel   fixed-order sum of the partials (bit-reproducible)
//
// All three main kernels are warp-streaming column strips (lane == column, lanes 0 / 31 are the halo columns of the 3x3
// window, horizontal neighbours by warp shuffle), like the loss kernels.  Reflection padding never exists in memory: loads
// go through reflect_index, and its adjoint is folded into the gradient taps (a pixel next to the border collects the taps
// that the padding mirrored onto it).
#include "vsl_common.cuh"

namespace ppea {

constexpr int kHeadStripW = 30;      // output columns per warp
constexpr int kHeadRows = 8;         // output rows per warp task (forward, dx)
constexpr int kHeadThreads = 128;
constexpr int kHeadMaxC = 256;       // weights live in shared memory: C * 12 floats
#ifndef PPEA_DW_GROUP
#define PPEA_DW_GROUP 4
#endif
#ifndef PPEA_DW_BLOCK
#define PPEA_DW_BLOCK 8
#endif
constexpr int kDwGroup = PPEA_DW_GROUP;   // channels per warp task of the weight-gradient kernel
constexpr int kDwRows = 64;          // rows per warp task there

__device__ __forceinline__ float sigmoid_ref(float v) { return div_rn(1.f, add_rn(1.f, expf(-v))); }

struct HeadTask {
  int b, strip, seg, lane, gx, px;
  bool own;
};
__device__ __forceinline__ bool head_task(int strips, int segs, int B, int W, HeadTask& t) {
  int id = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (id >= B * strips * segs) return false;
  t.lane = threadIdx.x & 31;
  t.strip = id % strips;
  id /= strips;
  t.seg = id % segs;
  t.b = id / segs;
  t.gx = t.strip * kHeadStripW - 1 + t.lane;
  t.px = reflect_index(t.gx, W);
  t.own = t.lane >= 1 && t.lane <= 30 && t.gx < W;
  return true;
}

// weights to shared memory as [c][12] (three rows of (w0, w1, w2, pad)): a channel costs three 128-bit broadcast loads
__device__ __forceinline__ void stage_weights(const float* __restrict__ w, int C, float4* s_w) {
  for (int i = threadIdx.x; i < C * 3; i += blockDim.x) {
    const float* r = w + (size_t)i * 3;
    s_w[i] = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), 0.f);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kHeadThreads) disp_head_forward_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                         const float* __restrict__ bias, float* __restrict__ disp,
                                                                         float* __restrict__ depth, int B, int C, int H, int W, int strips,
                                                                         int segs, float disp_lo, float disp_range) {
  extern __shared__ float4 s_w[];
  stage_weights(w, C, s_w);
  HeadTask t;
  if (!head_task(strips, segs, B, W, t)) return;
  const int r0 = t.seg * kHeadRows;
  const unsigned plane = (unsigned)(H * W);
  const float* xb = x + (size_t)t.b * C * plane;
  // row offsets of the kHeadRows + 2 input rows (reflected: nn.ReflectionPad2d(1))
  unsigned off[kHeadRows + 2];
#pragma unroll
  for (int j = 0; j < kHeadRows + 2; ++j) off[j] = (unsigned)reflect_index(r0 - 1 + j, H) * (unsigned)W + (unsigned)t.px;
  float acc[kHeadRows];
  const float bv = __ldg(bias);
#pragma unroll
  for (int i = 0; i < kHeadRows; ++i) acc[i] = bv;
  // the kernel is a stream of x through 9 FMAs per value: what bounds it is the number of loads in flight, so the rows of
  // channel c + 1 are requested before channel c is accumulated (two register sets), and tasks are short (many warps)
  float v[kHeadRows + 2], vn[kHeadRows + 2];
#pragma unroll
  for (int j = 0; j < kHeadRows + 2; ++j) vn[j] = __ldg(xb + off[j]);
#pragma unroll 1
  for (int c = 0; c < C; ++c) {
#pragma unroll
    for (int j = 0; j < kHeadRows + 2; ++j) v[j] = vn[j];
    if (c + 1 < C) {
      const float* xc = xb + (size_t)(c + 1) * plane;
#pragma unroll
      for (int j = 0; j < kHeadRows + 2; ++j) vn[j] = __ldg(xc + off[j]);
    }
    const float4 w0 = s_w[c * 3], w1 = s_w[c * 3 + 1], w2 = s_w[c * 3 + 2];
#pragma unroll
    for (int j = 0; j < kHeadRows + 2; ++j) {
      const float l = __shfl_up_sync(0xffffffffu, v[j], 1), r = __shfl_down_sync(0xffffffffu, v[j], 1);
      // input row j is window row ky of output row j - ky
      if (j < kHeadRows) acc[j] = fmaf(w0.z, r, fmaf(w0.y, v[j], fmaf(w0.x, l, acc[j])));
      if (j >= 1 && j <= kHeadRows) acc[j - 1] = fmaf(w1.z, r, fmaf(w1.y, v[j], fmaf(w1.x, l, acc[j - 1])));
      if (j >= 2) acc[j - 2] = fmaf(w2.z, r, fmaf(w2.y, v[j], fmaf(w2.x, l, acc[j - 2])));
    }
  }
  if (!t.own) return;
#pragma unroll
  for (int i = 0; i < kHeadRows; ++i) {
    const int y = r0 + i;
    if (y < H) {
      const size_t o = (size_t)t.b * plane + (unsigned)y * (unsigned)W + (unsigned)t.gx;
      const float s = sigmoid_ref(acc[i]);
      disp[o] = s;
      if (depth) depth[o] = depth_from_disp(s, disp_lo, disp_range);      // layers.py:19-22
    }
  }
}

// g = grad_disp * s (1 - s) of one row at the three window columns, with the adjoint of the column reflection folded in:
// tap kx of the window of q reads column q.x + kx - 1, so pixel p.x collects kx = 0 from its RIGHT neighbour, kx = 2 from its
// left one; column 1 also collects what the padding mirrored from column 0 (kx = 0 of q.x = 0), column W - 2 from column W - 1.
struct GRow {
  float k0, k1, k2;      // multiplies w[.][0], w[.][1], w[.][2]
};
__device__ __forceinline__ GRow grad_row(const float* __restrict__ gd, const float* __restrict__ sd, int y, int H, int W, const HeadTask& t,
                                         unsigned img) {
  float g = 0.f;
  if (y >= 0 && y < H && t.gx >= 0 && t.gx < W) {
    const size_t o = (size_t)img + (unsigned)y * (unsigned)W + (unsigned)t.gx;
    const float s = __ldg(sd + o);
    g = __ldg(gd + o) * s * (1.f - s);
  }
  const float gl = __shfl_up_sync(0xffffffffu, g, 1), gr = __shfl_down_sync(0xffffffffu, g, 1);      // (zero outside the image)
  GRow r;
  r.k1 = g;
  r.k0 = gr + (t.gx == 1 ? gl : 0.f);
  r.k2 = gl + (t.gx == W - 2 ? gr : 0.f);
  return r;
}

__global__ void __launch_bounds__(kHeadThreads) disp_head_dx_kernel(const float* __restrict__ w, const float* __restrict__ disp,
                                                                    const float* __restrict__ grad_disp, float* __restrict__ grad_x, int B,
                                                                    int C, int H, int W, int strips, int segs) {
  extern __shared__ float4 s_w[];
  stage_weights(w, C, s_w);
  HeadTask t;
  if (!head_task(strips, segs, B, W, t)) return;
  const int r0 = t.seg * kHeadRows;
  const unsigned plane = (unsigned)(H * W);
  const unsigned img = (unsigned)t.b * plane;
  // rows r0 - 1 .. r0 + kHeadRows of g (zero outside the image); window row ky of q reads row q.y + ky - 1, so output row y collects
  // ky = 0 from row y + 1, ky = 2 from row y - 1, and rows 1 / H - 2 also what the padding mirrored from rows 0 / H - 1
  GRow G[kHeadRows + 2];
#pragma unroll
  for (int j = 0; j < kHeadRows + 2; ++j) G[j] = grad_row(grad_disp, disp, r0 - 1 + j, H, W, t, img);
  GRow top = {0.f, 0.f, 0.f}, bot = {0.f, 0.f, 0.f};      // mirrored contributions (row 0 -> row 1 under ky = 0, row H-1 -> row H-2 under ky = 2)
  const bool has_top = r0 <= 1 && 1 < r0 + kHeadRows, has_bot = r0 <= H - 2 && H - 2 < r0 + kHeadRows;
  if (has_top) top = grad_row(grad_disp, disp, 0, H, W, t, img);
  if (has_bot) bot = grad_row(grad_disp, disp, H - 1, H, W, t, img);
  float* gxb = grad_x + (size_t)t.b * C * plane;
#pragma unroll 1
  for (int c = 0; c < C; ++c) {
    const float4 w0 = s_w[c * 3], w1 = s_w[c * 3 + 1], w2 = s_w[c * 3 + 2];
    const float e_top = fmaf(w0.z, top.k2, fmaf(w0.y, top.k1, w0.x * top.k0));
    const float e_bot = fmaf(w2.z, bot.k2, fmaf(w2.y, bot.k1, w2.x * bot.k0));
    float* gc = gxb + (size_t)c * plane;
#pragma unroll
    for (int i = 0; i < kHeadRows; ++i) {
      const int y = r0 + i;
      float a = fmaf(w0.z, G[i + 2].k2, fmaf(w0.y, G[i + 2].k1, w0.x * G[i + 2].k0));
      a = fmaf(w1.z, G[i + 1].k2, fmaf(w1.y, G[i + 1].k1, fmaf(w1.x, G[i + 1].k0, a)));
      a = fmaf(w2.z, G[i].k2, fmaf(w2.y, G[i].k1, fmaf(w2.x, G[i].k0, a)));
      if (y == 1) a += e_top;
      if (y == H - 2) a += e_bot;
      if (t.own && y < H) gc[(unsigned)y * (unsigned)W + (unsigned)t.gx] = a;
    }
  }
}

// grad_w[c][ky][kx] = sum_q g(q) xpad[c](q + k - 1) = sum_p x[c](p) * tap[ky][kx](p) with the same folded taps as above.
// A warp walks kDwRows rows of one strip for kDwGroup channels at once: the taps of a row are formed once and used by the
// four channels; 36 + 1 per-lane partial sums, reduced over the warp once per task.
#ifndef PPEA_DW_CTAS
#define PPEA_DW_CTAS 4
#endif
__global__ void __launch_bounds__(kHeadThreads, PPEA_DW_CTAS) disp_head_dw_kernel(const float* __restrict__ x, const float* __restrict__ disp,
                                                                    const float* __restrict__ grad_disp, float* __restrict__ partials, int B,
                                                                    int C, int H, int W, int strips, int segs, int groups) {
  int id = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (id >= B * strips * segs * groups) return;
  const int task = id;
  const int grp = id % groups;
  id /= groups;
  HeadTask t;
  t.lane = threadIdx.x & 31;
  t.strip = id % strips;
  id /= strips;
  t.seg = id % segs;
  t.b = id / segs;
  t.gx = t.strip * kHeadStripW - 1 + t.lane;
  t.px = reflect_index(t.gx, W);
  t.own = t.lane >= 1 && t.lane <= 30 && t.gx < W;
  const int r0 = t.seg * kDwRows, r1 = min(r0 + kDwRows, H);
  const unsigned plane = (unsigned)(H * W);
  const unsigned img = (unsigned)t.b * plane;
  const int c0 = grp * kDwGroup;
  const float* xb = x + ((size_t)t.b * C + c0) * plane;
  float acc[kDwGroup][9];
#pragma unroll
  for (int k = 0; k < kDwGroup; ++k)
#pragma unroll
    for (int e = 0; e < 9; ++e) acc[k][e] = 0.f;
  float accb = 0.f;
  const GRow zero = {0.f, 0.f, 0.f};
  GRow up = grad_row(grad_disp, disp, r0 - 1, H, W, t, img), cur = grad_row(grad_disp, disp, r0, H, W, t, img);
  const GRow top = grad_row(grad_disp, disp, 0, H, W, t, img), bot = grad_row(grad_disp, disp, H - 1, H, W, t, img);
  const bool in_col = t.gx >= 0 && t.gx < W;
  // (bases the optimiser cannot see through: every address formed from them is one IMAD.WIDE.U32 of a 32-bit offset)
  auto opaque = [](const float* p) {
    unsigned long long v = reinterpret_cast<unsigned long long>(p);
    asm volatile("" : "+l"(v));
    return reinterpret_cast<const float*>(v);
  };
  const float* gd_c = opaque(grad_disp + (size_t)img + (unsigned)(in_col ? t.gx : 0));
  const float* sd_c = opaque(disp + (size_t)img + (unsigned)(in_col ? t.gx : 0));
  const float* x_c = opaque(xb + (unsigned)(t.own ? t.gx : 0));
  unsigned koff[kDwGroup];              // plane offsets of the group's channels
#pragma unroll
  for (int k = 0; k < kDwGroup; ++k) koff[k] = (unsigned)min(k, C - c0 - 1) * plane;
  constexpr int R = PPEA_DW_BLOCK;      // rows per block: R x (kDwGroup + 2) loads in flight per lane (the kernel is bound by loads in flight)
#pragma unroll 1
  for (int yb = r0; yb < r1; yb += R) {
    float xv[R][kDwGroup], gr[R], sr[R];
    // Loads of the block, with no per-load tests: a row past the segment is clamped to its last row (the block loop below
    // skips it), a channel past C to the last channel of the group (its sums land in partial slots nobody reads), offsets are
    // 32-bit (C * H * W < 2^31 is checked by the entry point).  The first version spent ~9 integer instructions per load on
    // row / channel predicates and 64-bit addresses: 42 % of the loop's SASS.
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int yc = min(yb + i, r1 - 1);
      const unsigned orow = (unsigned)yc * (unsigned)W;
#pragma unroll
      for (int k = 0; k < kDwGroup; ++k) xv[i][k] = t.own ? __ldg(x_c + (koff[k] + orow)) : 0.f;
      const bool dn_ok = yc + 1 < H && in_col;               // row y + 1 of g (zero outside the image)
      gr[i] = dn_ok ? __ldg(gd_c + (orow + (unsigned)W)) : 0.f;
      sr[i] = dn_ok ? __ldg(sd_c + (orow + (unsigned)W)) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int y = yb + i;
      if (y < r1) {      // (warp-uniform)
        const float g = gr[i] * sr[i] * (1.f - sr[i]);
        const float gl = __shfl_up_sync(0xffffffffu, g, 1), grt = __shfl_down_sync(0xffffffffu, g, 1);
        GRow dn;
        dn.k1 = g;
        dn.k0 = grt + (t.gx == 1 ? gl : 0.f);
        dn.k2 = gl + (t.gx == W - 2 ? grt : 0.f);
        // taps of pixel (y, gx): ky = 0 <- row y + 1 (+ row 0 at y == 1), ky = 1 <- row y, ky = 2 <- row y - 1 (+ row H - 1 at y == H - 2)
        const GRow e0 = (y == 1) ? top : zero, e2 = (y == H - 2) ? bot : zero;
        const float t0 = dn.k0 + e0.k0, t1 = dn.k1 + e0.k1, t2 = dn.k2 + e0.k2;
        const float t6 = up.k0 + e2.k0, t7 = up.k1 + e2.k1, t8 = up.k2 + e2.k2;
#pragma unroll
        for (int k = 0; k < kDwGroup; ++k) {
          acc[k][0] = fmaf(xv[i][k], t0, acc[k][0]);
          acc[k][1] = fmaf(xv[i][k], t1, acc[k][1]);
          acc[k][2] = fmaf(xv[i][k], t2, acc[k][2]);
          acc[k][3] = fmaf(xv[i][k], cur.k0, acc[k][3]);
          acc[k][4] = fmaf(xv[i][k], cur.k1, acc[k][4]);
          acc[k][5] = fmaf(xv[i][k], cur.k2, acc[k][5]);
          acc[k][6] = fmaf(xv[i][k], t6, acc[k][6]);
          acc[k][7] = fmaf(xv[i][k], t7, acc[k][7]);
          acc[k][8] = fmaf(xv[i][k], t8, acc[k][8]);
        }
        if (t.own) accb += cur.k1;
        up = cur;
        cur = dn;
      }
    }
  }
  float* out = partials + (size_t)task * (kDwGroup * 9 + 1);
#pragma unroll
  for (int k = 0; k < kDwGroup; ++k)
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      const float s = warp_sum(acc[k][e]);
      if (t.lane == 0) out[k * 9 + e] = s;
    }
  accb = warp_sum(accb);
  if (t.lane == 0) out[kDwGroup * 9] = accb;
}

// one WARP per weight (and one for the bias): lane l sums the tasks l, l + 32, ... in double, then a fixed xor tree -- the order of
// addition is fixed, so the result is bit-reproducible.  (The first version, one THREAD per weight walking all ~800 task partials,
// took 67 us for 289 threads: one dependent DRAM round trip after another, ncu r2v: long scoreboard 21.8 per issue.)
__global__ void __launch_bounds__(128) disp_head_dw_finish_kernel(const float* __restrict__ partials, float* __restrict__ grad_w,
                                                                   float* __restrict__ grad_b, int C, int n_spatial, int groups) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i > C * 9) return;          // (whole warp)
  const bool bias = i == C * 9;
  const int c = bias ? 0 : i / 9, e = i % 9;
  const int grp = c / kDwGroup, k4 = c % kDwGroup;
  const size_t stride = (size_t)groups * (kDwGroup * 9 + 1);
  const float* p = partials + (size_t)grp * (kDwGroup * 9 + 1) + (bias ? kDwGroup * 9 : k4 * 9 + e);   // bias: the partials of channel group 0
  double s = 0.0;
#pragma unroll 4
  for (int k = lane; k < n_spatial; k += 32) s += (double)__ldg(p + (size_t)k * stride);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    if (bias) {
      if (grad_b) *grad_b = (float)s;
    } else {
      grad_w[i] = (float)s;
    }
  }
}

static bool head_shape_ok(int B, int C, int H, int W) {
  return B > 0 && C > 0 && C <= kHeadMaxC && H >= 2 && W >= 2 && (long long)B * C * H * W < (1ll << 40) && (long long)C * H * W < (1ll << 31) &&
         (long long)B * H * W < (1ll << 31);
}

}  // namespace ppea

using namespace ppea;

extern "C" int ppea_disp_head_forward(const float* x, const float* weight, const float* bias, float* disp, float* depth_or_null, int batch,
                                      int channels, int height, int width, float min_depth, float max_depth, void* stream) {
  if (!x || !weight || !bias || !disp) return PPEA_E_NULL;
  if (!head_shape_ok(batch, channels, height, width)) return PPEA_E_SHAPE;
  if (depth_or_null && !(min_depth > 0.f && max_depth > min_depth)) return PPEA_E_SHAPE;
  const int strips = ceil_div(width, kHeadStripW), segs = ceil_div(height, kHeadRows);
  const int ctas = ceil_div(batch * strips * segs, kHeadThreads / 32);
  const float lo = depth_or_null ? 1.f / max_depth : 0.f, range = depth_or_null ? 1.f / min_depth - 1.f / max_depth : 0.f;      // layers.py:19-21
  disp_head_forward_kernel<<<ctas, kHeadThreads, (size_t)channels * 3 * sizeof(float4), (cudaStream_t)stream>>>(
      x, weight, bias, disp, depth_or_null, batch, channels, height, width, strips, segs, lo, range);
  return (int)cudaGetLastError();
}

extern "C" size_t ppea_disp_head_workspace_bytes(int batch, int channels, int height, int width) {
  if (!head_shape_ok(batch, channels, height, width)) return 0;
  const size_t n_spatial = (size_t)batch * ceil_div(width, kHeadStripW) * ceil_div(height, kDwRows);
  return n_spatial * ceil_div(channels, kDwGroup) * (kDwGroup * 9 + 1) * sizeof(float);
}

extern "C" int ppea_disp_head_backward(const float* x, const float* weight, const float* disp, const float* grad_disp, float* grad_x_or_null,
                                       float* grad_weight_or_null, float* grad_bias_or_null, void* workspace, int batch, int channels,
                                       int height, int width, void* stream) {
  if (!weight || !disp || !grad_disp) return PPEA_E_NULL;
  if (!head_shape_ok(batch, channels, height, width)) return PPEA_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const int strips = ceil_div(width, kHeadStripW);
  if (grad_x_or_null) {
    const int segs = ceil_div(height, kHeadRows);
    const int ctas = ceil_div(batch * strips * segs, kHeadThreads / 32);
    disp_head_dx_kernel<<<ctas, kHeadThreads, (size_t)channels * 3 * sizeof(float4), st>>>(weight, disp, grad_disp, grad_x_or_null, batch,
                                                                                          channels, height, width, strips, segs);
  }
  if (grad_weight_or_null || grad_bias_or_null) {
    if (!x || !workspace) return PPEA_E_NULL;
    if (!grad_weight_or_null) return PPEA_E_NULL;
    const int segs = ceil_div(height, kDwRows), groups = ceil_div(channels, kDwGroup);
    const int n_spatial = batch * strips * segs;
    const int ctas = ceil_div(n_spatial * groups, kHeadThreads / 32);
    disp_head_dw_kernel<<<ctas, kHeadThreads, 0, st>>>(x, disp, grad_disp, (float*)workspace, batch, channels, height, width, strips, segs, groups);
    disp_head_dw_finish_kernel<<<ceil_div(channels * 9 + 1, 4), 128, 0, st>>>((const float*)workspace, grad_weight_or_null, grad_bias_or_null,
                                                                              channels, n_spatial, groups);
  }
  return (int)cudaGetLastError();
}
