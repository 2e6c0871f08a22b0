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

// A role CTA owns one chunk of one image of one scale: a strip of blockDim.x columns x a band of rows.
// Threads walk DOWN their column: the row just loaded serves as "down neighbour" of the previous row and as
// "current" of the next, and the right neighbour comes by warp shuffle, so a pixel costs one 4-word load
// (disp + 3 colour channels) instead of 12-29.
struct SmoothChunk {
  int x, y_lo, y_hi;      // column of this thread (may be >= w), row band
  bool active;            // chunk id maps to work
};
__device__ __forceinline__ SmoothChunk smooth_chunk(int w, int h, int chunk, int bd, int tid) {
  const int strips = (w + bd - 1) / bd;                    // <= kSmoothChunks is checked by the API (w <= 32 * 128)
  const int bands = kSmoothChunks / strips > 0 ? kSmoothChunks / strips : 1;
  const int rows = (h + bands - 1) / bands;
  SmoothChunk c;
  const int strip = chunk % strips, band = chunk / strips;
  c.active = band < bands && band * rows < h;
  c.x = strip * bd + tid;
  c.y_lo = band * rows;
  c.y_hi = c.y_lo + rows < h ? c.y_lo + rows : h;
  return c;
}
__device__ __forceinline__ SmoothChunk smooth_chunk(int w, int h, int chunk) { return smooth_chunk(w, h, chunk, blockDim.x, threadIdx.x); }

struct SmoothPx {
  float d, i0, i1, i2;
};
__device__ __forceinline__ SmoothPx smooth_load(const float* __restrict__ d, const float* __restrict__ img, unsigned plane,
                                                unsigned o, bool ok) {
  SmoothPx p = {0.f, 0.f, 0.f, 0.f};
  if (ok) {
    p.d = __ldg(d + o);
    p.i0 = __ldg(img + o);
    p.i1 = __ldg(img + plane + o);
    p.i2 = __ldg(img + 2u * plane + o);
  }
  return p;
}
__device__ __forceinline__ SmoothPx smooth_shfl_down(const SmoothPx& p) {
  SmoothPx q;
  q.d = __shfl_down_sync(0xffffffffu, p.d, 1);
  q.i0 = __shfl_down_sync(0xffffffffu, p.i0, 1);
  q.i1 = __shfl_down_sync(0xffffffffu, p.i1, 1);
  q.i2 = __shfl_down_sync(0xffffffffu, p.i2, 1);
  return q;
}
__device__ __forceinline__ float smooth_weight(const SmoothPx& a, const SmoothPx& b) {
  // exp(-mean_c |I_a - I_b|)   (layers.py:217-221)
  return __expf(-(fabsf(a.i0 - b.i0) + fabsf(a.i1 - b.i1) + fabsf(a.i2 - b.i2)) * (1.f / 3.f));
}

// Forward role: raw sums of one chunk -> smooth_ws[(s, b, chunk)][3] = (sum d, X, Y).  `red` holds 3*nwarps floats.
__device__ __forceinline__ void smooth_forward_role(const VslArgs& a, int role, float* red) {
  int s, b, chunk;
  smooth_role_ids(a, role, s, b, chunk);
  const ScaleArgs& sc = a.sc[s];
  const int h = sc.hs, w = sc.ws;
  const unsigned n = (unsigned)(h * w);
  const float* d = sc.disp + (size_t)b * n;
  const float* img = sc.color + (size_t)b * 3 * n;
  float* gz = sc.grad_disp ? sc.grad_disp + (size_t)b * n : nullptr;   // optional: pre-zero the backward's accumulator
  const SmoothChunk ck = smooth_chunk(w, h, chunk);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float sd = 0.f, sx = 0.f, sy = 0.f;
  if (ck.active) {
    const bool xin = ck.x < w, has_r = ck.x + 1 < w;
    SmoothPx cur = smooth_load(d, img, n, (unsigned)(ck.y_lo * w + ck.x), xin);
    for (int y = ck.y_lo; y < ck.y_hi; ++y) {
      const unsigned o = (unsigned)(y * w + ck.x);
      const bool has_d = y + 1 < h;
      const SmoothPx nxt = smooth_load(d, img, n, o + (unsigned)w, xin && has_d);
      SmoothPx rgt = smooth_shfl_down(cur);
      if (lane == 31) rgt = smooth_load(d, img, n, o + 1u, has_r);     // the neighbour lives in the next warp / strip
      if (xin) {
        sd += cur.d;
        if (gz) gz[o] = 0.f;
        if (has_r) sx += fabsf(cur.d - rgt.d) * smooth_weight(cur, rgt);
        if (has_d) sy += fabsf(cur.d - nxt.d) * smooth_weight(cur, nxt);
      }
      cur = nxt;
    }
  }
  sd = warp_sum(sd);
  sx = warp_sum(sx);
  sy = warp_sum(sy);
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

// Fused-step role (vsl_fused.cu): the forward sums of one chunk AND the un-normalised stencil field
//   st(x,y) = ( R(x,y) - R(x-1,y) ) / N_x + ( D(x,y) - D(x,y-1) ) / N_y,
//   R = sign(d - d_right) * e_right,  D = sign(d - d_down) * e_down,
// from the same column walk (the edge weights are the forward's own), so that the backward proper is
// elementwise:  d smooth / d d_j = inv_b * ( st_j - inv_b * (X_b / N_x + Y_b / N_y) / (h*w) )  (times upstream).
// The column walk of one (virtual) thread `tid` of a `bd`-wide chunk CTA; adds its share to (sd, sx, sy).
// A warp covers 32 adjacent columns of which the inner 30 are its own: both horizontal neighbours of an owned column
// live in the warp and come by shuffle, so a row costs each lane ONE 4-word load (disp + 3 colour channels), issued a row
// ahead -- no lane waits for a neighbour it had to load itself.  A chunk is a strip of 30 * (bd / 32) columns x a band of rows.
__device__ __forceinline__ void smooth_fused_walk(const VslArgs& a, int s, int b, int chunk, int bd, int tid, float& sd, float& sx, float& sy) {
  const ScaleArgs& sc = a.sc[s];
  const int h = sc.hs, w = sc.ws;
  const unsigned n = (unsigned)(h * w);
  const float* d = sc.disp + (size_t)b * n;
  const float* img = sc.color + (size_t)b * 3 * n;
  float* st = sc.grad_st + (size_t)b * n;
  const int lane = tid & 31, cols = 30 * (bd >> 5);
  const int strips = (w + cols - 1) / cols;                 // <= kSmoothChunks is checked by the API (w <= 32 * 120)
  const int bands = kSmoothChunks / strips > 0 ? kSmoothChunks / strips : 1;
  const int rows = (h + bands - 1) / bands;
  const int strip = chunk % strips, band = chunk / strips;
  if (!(band < bands && band * rows < h)) return;
  const int x = strip * cols + (tid >> 5) * 30 + lane - 1;
  const int y_lo = band * rows, y_hi = y_lo + rows < h ? y_lo + rows : h;
  const float cx = w > 1 ? 1.f / ((float)a.B * h * (w - 1)) : 0.f, cy = h > 1 ? 1.f / ((float)a.B * (h - 1) * w) : 0.f;
  const bool xin = x >= 0 && x < w, has_r = xin && x + 1 < w;
  const bool own = xin && lane >= 1 && lane <= 30;
  const unsigned ux = (unsigned)(xin ? x : 0);
  // Three row registers in rotation (the loop is unrolled by three so that the roles are names, not moves): a row is
  // requested two iterations before it is used, and nothing touches its registers in between -- a register move of a
  // load result that has not arrived yet would stall for the whole memory latency.
  SmoothPx r0 = smooth_load(d, img, n, (unsigned)(y_lo * w) + ux, xin);
  SmoothPx r1 = smooth_load(d, img, n, (unsigned)((y_lo + 1) * w) + ux, xin && y_lo + 1 < h);
  SmoothPx r2;
  float d_up = 0.f;                      // D(x, y-1)
  if (y_lo > 0 && xin) {
    const SmoothPx up = smooth_load(d, img, n, (unsigned)((y_lo - 1) * w) + ux, true);
    d_up = sign_of(up.d - r0.d) * smooth_weight(up, r0);
  }
  auto row = [&](int y, const SmoothPx& cur, const SmoothPx& nxt, SmoothPx& nn) {
    const unsigned o = (unsigned)(y * w) + ux;
    const bool has_d = y + 1 < h;
    nn = smooth_load(d, img, n, o + 2u * (unsigned)w, xin && y + 2 < h && y + 2 <= y_hi);
    const SmoothPx rgt = smooth_shfl_down(cur);
    const float dr = cur.d - rgt.d, dd = cur.d - nxt.d;
    const float er = has_r ? smooth_weight(cur, rgt) : 0.f;
    const float ed = (xin && has_d) ? smooth_weight(cur, nxt) : 0.f;
    const float r_here = sign_of(dr) * er, d_here = sign_of(dd) * ed;
    const float r_left = __shfl_up_sync(0xffffffffu, r_here, 1);      // (0 at the image border: lane 0 is outside, er = 0)
    if (own) {
      sd += cur.d;
      sx += fabsf(dr) * er;
      sy += fabsf(dd) * ed;
      st[o] = (r_here - r_left) * cx + (d_here - d_up) * cy;
    }
    d_up = d_here;
  };
#pragma unroll 1
  for (int y = y_lo; y < y_hi; y += 3) {
    row(y, r0, r1, r2);
    if (y + 1 < y_hi) row(y + 1, r1, r2, r0);
    if (y + 2 < y_hi) row(y + 2, r2, r0, r1);
  }
}

__device__ __forceinline__ void smooth_fused_role(const VslArgs& a, int role, float* red) {
  int s, b, chunk;
  smooth_role_ids(a, role, s, b, chunk);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float sd = 0.f, sx = 0.f, sy = 0.f;
  smooth_fused_walk(a, s, b, chunk, blockDim.x, threadIdx.x, sd, sx, sy);
  sd = warp_sum(sd);
  sx = warp_sum(sx);
  sy = warp_sum(sy);
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

// The same role run by ONE warp (a 32-thread CTA of the streaming kernel): the warp plays the four warps of a
// kSmoothThreads-wide chunk CTA one after the other; same chunk geometry, same per-warp sums, same order of addition.
__device__ __forceinline__ void smooth_fused_role_warp(const VslArgs& a, int role) {
  int s, b, chunk;
  smooth_role_ids(a, role, s, b, chunk);
  const int lane = threadIdx.x & 31;
  float tot[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
  for (int sub = 0; sub < kSmoothThreads / 32; ++sub) {
    float sd = 0.f, sx = 0.f, sy = 0.f;
    smooth_fused_walk(a, s, b, chunk, kSmoothThreads, sub * 32 + lane, sd, sx, sy);
    tot[0] += warp_sum(sd);
    tot[1] += warp_sum(sx);
    tot[2] += warp_sum(sy);
  }
  if (lane < 3) a.smooth_ws[(((size_t)s * a.B + b) * kSmoothChunks + chunk) * 3 + lane] = lane == 0 ? tot[0] : (lane == 1 ? tot[1] : tot[2]);
}

// Backward role: smoothness gradient of one chunk.  MODE 0: overwrite grad_disp;  1: atomically add into a
// zero-initialised (and concurrently accumulated) grad_disp;  2 (fused step, vsl_fused.cu): grad_disp =
// smoothness gradient + w_s * grad_raw with w_s = upstream(reproj_s) / (sum(mask_s) + 1e-7).
//   grad(x,y) = inv * ( R(x,y) - R(x-1,y) + D(x,y) - D(x,y-1) - mean_term ),
//   R = gx * sign(d - d_right) * e_right,  D = gy * sign(d - d_down) * e_down.
template <int MODE>
__device__ __forceinline__ void smooth_backward_role(const VslArgs& a, int role) {
  int s, b, chunk;
  smooth_role_ids(a, role, s, b, chunk);
  const ScaleArgs& sc = a.sc[s];
  const int h = sc.hs, w = sc.ws;
  const unsigned n = (unsigned)(h * w);
  const SmoothChunk ck = smooth_chunk(w, h, chunk);
  if (!ck.active) return;
  const float* row = a.sums + (size_t)s * sums_stride(a.B) + PPEA_SUMS_PER_SCALE + 4 * b;   // (sum d, X_b, Y_b)
  const float inv = 1.f / (row[0] / (float)n + 1e-7f);
  const float g = scale_grads(a, s).smooth;
  const float gx = g / ((float)a.B * h * (w - 1)), gy = g / ((float)a.B * (h - 1) * w);
  const float mean_term = inv * (gx * row[1] + gy * row[2]) / (float)n;
  const float* d = sc.disp + (size_t)b * n;
  const float* img = sc.color + (size_t)b * 3 * n;
  float* out = sc.grad_disp + (size_t)b * n;
  const float* raw = (MODE == 2) ? sc.grad_raw + (size_t)b * n : nullptr;
  const float w_raw = (MODE == 2) ? scale_grads(a, s).reproj / (a.sums[(size_t)s * sums_stride(a.B) + 1] + 1e-7f) : 0.f;
  const int lane = threadIdx.x & 31;
  const bool xin = ck.x < w, has_r = ck.x + 1 < w, has_l = xin && ck.x > 0;
  SmoothPx cur = smooth_load(d, img, n, (unsigned)(ck.y_lo * w + ck.x), xin);
  float d_up = 0.f;                      // D(x, y-1)
  if (ck.y_lo > 0 && xin) {
    const SmoothPx up = smooth_load(d, img, n, (unsigned)((ck.y_lo - 1) * w + ck.x), true);
    d_up = gy * sign_of(up.d - cur.d) * smooth_weight(up, cur);
  }
  for (int y = ck.y_lo; y < ck.y_hi; ++y) {
    const unsigned o = (unsigned)(y * w + ck.x);
    const bool has_d = y + 1 < h;
    const SmoothPx nxt = smooth_load(d, img, n, o + (unsigned)w, xin && has_d);
    SmoothPx rgt = smooth_shfl_down(cur);
    if (lane == 31) rgt = smooth_load(d, img, n, o + 1u, has_r);
    const float r_here = (xin && has_r) ? gx * sign_of(cur.d - rgt.d) * smooth_weight(cur, rgt) : 0.f;
    float r_left = __shfl_up_sync(0xffffffffu, r_here, 1);
    if (lane == 0) {
      r_left = 0.f;
      if (has_l) {
        const SmoothPx lft = smooth_load(d, img, n, o - 1u, true);
        r_left = gx * sign_of(lft.d - cur.d) * smooth_weight(lft, cur);
      }
    }
    const float d_here = (xin && has_d) ? gy * sign_of(cur.d - nxt.d) * smooth_weight(cur, nxt) : 0.f;
    if (xin) {
      const float v = ((r_here - r_left) + (d_here - d_up) - mean_term) * inv;
      if (MODE == 1)
        atomicAdd(out + o, v);
      else if (MODE == 2)
        out[o] = fmaf(w_raw, __ldg(raw + o), v);
      else
        out[o] = v;
    }
    d_up = d_here;
    cur = nxt;
  }
}

}  // namespace ppea
