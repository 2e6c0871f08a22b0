// matching.cu -- plane-sweep cost volume of the multi-frame encoder.
//
// Reference: `match_features`, networks/replk_matching_adapter.py:261-340 (the same code sits in
// replk_matching.py:127-206 and resnet_encoder.py:164-246).  Per batch item the reference repeats the lookup
// features num_depth_bins times, back-projects / projects a (D,h,w) grid through BackprojectDepth + Project3D,
// calls F.grid_sample(padding_mode="zeros", align_corners=True), takes the channel-mean L1 difference to the
// current features, masks it at the borders, averages over the lookup frames and fills the depth bins that never
// landed inside the image with the per-pixel maximum -- ~15 launches and (D,C,h,w) temporaries per item, in a
// Python loop over the batch.
//
// Here: one launch for the volume (+ one small fix-up launch).  A thread owns one pixel of one batch item and a chunk of
// kMatchChunk depth bins (grid.z; measured best with one group of four bins per CTA: 17 280 CTAs at the KITTI shape, the
// 128 pixels of a CTA sample one contiguous footprint per bin) and walks them kMatchBins at a time: the
// projection of the hypotheses is set up once (corner offsets + bilinear weights in registers), then the channel loop
// reads the current feature once per channel and the four corners of every hypothesis (lanes are adjacent pixels, so
// every load of the warp is a contiguous run of one channel plane).  "Set missing to max" needs the per-pixel maximum
// over ALL bins: a second, bandwidth-bound kernel sweeps the finished volume (missing <=> cost == 0).  Nothing but the
// cost volume and its mask is written; there is no workspace.
//
// Arithmetic follows the reference op by op where a rounding can flip a decision (the border masks compare
// x_vals / y_vals with 2 and size-2): P = K @ T, ray = inv_K[:3,:3] @ (x,y,1), X = depth * ray (layers.py:164-167),
// cam = P @ (X,1), pix = cam.xy / (cam.z + eps), normalised to [-1,1] (layers.py:187-198), x_vals = (g/2 + .5)(w-1)
// (:302-304), sampling position ((g+1)/2)(w-1) (ATen grid_sampler_unnormalize, align_corners=True).
#include "vsl_common.cuh"

namespace ppea {

#ifndef PPEA_MATCH_THREADS
#define PPEA_MATCH_THREADS 128
#endif
constexpr int kMatchThreads = PPEA_MATCH_THREADS;
#ifndef PPEA_MATCH_BINS
#define PPEA_MATCH_BINS 4
#endif
#ifndef PPEA_MATCH_CHUNK
#define PPEA_MATCH_CHUNK 8
#endif
constexpr int kMatchBins = PPEA_MATCH_BINS;        // depth hypotheses in flight per thread
constexpr int kMatchChunk = PPEA_MATCH_CHUNK;     // depth bins per CTA (grid.z = ceil(D / kMatchChunk))
constexpr int kMatchMaxBins = 4096;
#ifndef PPEA_MATCH_TILE_W
#define PPEA_MATCH_TILE_W 32
#endif
// A CTA covers a kMatchTileW x (kMatchThreads / kMatchTileW) pixel tile (lane == column inside a row of the tile): the footprint
// of a hypothesis in the lookup features is then ~(TW+1) x (TH+1) cells instead of the 2 x 129 of a 128-pixel row segment, so
// fewer sectors come from L2 per CTA.
#if PPEA_MATCH_TILE_W == 0
constexpr int kMatchTileW = 0, kMatchTileH = 0;
inline int match_tiles(int h, int w) { return ceil_div(h * w, kMatchThreads); }
#else
constexpr int kMatchTileW = PPEA_MATCH_TILE_W, kMatchTileH = kMatchThreads / kMatchTileW;
inline int match_tiles(int h, int w) { return ceil_div(w, kMatchTileW) * ceil_div(h, kMatchTileH); }
#endif

struct MatchArgs {
  const float* cur;     // (B,C,h,w)
  const float* look;    // (B,F,C,h,w)
  const float* poses;   // (B,F,4,4)
  const float* K;       // (B,4,4) of the matching scale
  const float* invK;    // (B,4,4)
  const float* bins;    // (D)
  float* cost;          // (B,D,h,w)
  float* missing;       // (B,D,h,w)
  int B, F, C, h, w, D;
  int set_missing_to_max;
  float eps;
};

struct MatchTap {
  int off;                     // element offset of the north-west corner inside a channel plane
  float wnw, wne, wsw, wse;    // bilinear weights with the zero padding folded in (0 for corners outside the image)
  float edge;                  // border mask of this hypothesis (:302-312)
};

__device__ __forceinline__ MatchTap match_setup(const float* __restrict__ P, const float* __restrict__ ray, float depth, int h, int w,
                                                float eps, float cur_mask) {
  // BackprojectDepth: X = depth * ray; Project3D: cam = P @ (X, 1)
  const float X = mul_rn(depth, ray[0]), Y = mul_rn(depth, ray[1]), Z = mul_rn(depth, ray[2]);
  float c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float acc = mul_rn(P[i * 4 + 0], X);
    acc = fma_rn(P[i * 4 + 1], Y, acc);
    acc = fma_rn(P[i * 4 + 2], Z, acc);
    acc = fma_rn(P[i * 4 + 3], 1.f, acc);
    c[i] = acc;
  }
  const float z = add_rn(c[2], eps);
  const float gx = mul_rn(sub_rn(div_rn(div_rn(c[0], z), (float)(w - 1)), 0.5f), 2.f);
  const float gy = mul_rn(sub_rn(div_rn(div_rn(c[1], z), (float)(h - 1)), 0.5f), 2.f);
  // border mask on the pixel coordinates the reference derives from the grid
  const float xv = mul_rn(add_rn(div_rn(gx, 2.f), 0.5f), (float)(w - 1));
  const float yv = mul_rn(add_rn(div_rn(gy, 2.f), 0.5f), (float)(h - 1));
  MatchTap t;
  t.edge = (xv >= 2.f && xv <= (float)(w - 2) && yv >= 2.f && yv <= (float)(h - 2)) ? cur_mask : 0.f;
  // grid_sample, zeros padding, align_corners=True
  const float ix = mul_rn(div_rn(add_rn(gx, 1.f), 2.f), (float)(w - 1));
  const float iy = mul_rn(div_rn(add_rn(gy, 1.f), 2.f), (float)(h - 1));
  // (NaN / far-out coordinates: every corner falls outside, all weights 0)
  const float fx = floorf(ix), fy = floorf(iy);
  const bool finite = (ix > -2.f && ix < (float)(w + 1) && iy > -2.f && iy < (float)(h + 1));
  const int x0 = finite ? (int)fx : -2, y0 = finite ? (int)fy : -2;
  const float tx = ix - fx, ty = iy - fy;
  const bool xin0 = x0 >= 0 && x0 < w, xin1 = x0 + 1 >= 0 && x0 + 1 < w;
  const bool yin0 = y0 >= 0 && y0 < h, yin1 = y0 + 1 >= 0 && y0 + 1 < h;
  t.wnw = (xin0 && yin0) ? (1.f - tx) * (1.f - ty) : 0.f;
  t.wne = (xin1 && yin0) ? tx * (1.f - ty) : 0.f;
  t.wsw = (xin0 && yin1) ? (1.f - tx) * ty : 0.f;
  t.wse = (xin1 && yin1) ? tx * ty : 0.f;
  // clamp the base so that the four addresses o, o+1, o+w, o+w+1 stay inside the plane (their weights are 0 where clamped)
  const int xc = min(max(x0, 0), w - 2), yc = min(max(y0, 0), h - 2);
  t.off = yc * w + xc;
  if (xc != x0) {          // the in-image corner moved: re-route its weight
    // x0 == -1: only the east corners are inside -> they sit at column 0 == the clamped west column
    // x0 == w-1: only the west corners are inside -> they sit at column w-1 == the clamped east column
    if (x0 == -1) { t.wnw = t.wne; t.wsw = t.wse; t.wne = 0.f; t.wse = 0.f; }
    else if (x0 == w - 1) { t.wne = t.wnw; t.wse = t.wsw; t.wnw = 0.f; t.wsw = 0.f; }
  }
  if (yc != y0) {
    if (y0 == -1) { t.wnw = t.wsw; t.wne = t.wse; t.wsw = 0.f; t.wse = 0.f; }
    else if (y0 == h - 1) { t.wsw = t.wnw; t.wse = t.wne; t.wnw = 0.f; t.wne = 0.f; }
  }
  return t;
}

__global__ void __launch_bounds__(kMatchThreads) match_features_kernel(const MatchArgs a) {
  __shared__ float sP[12];         // (K @ T)[:3,:] of the current lookup frame
  __shared__ float siK[9];
  __shared__ int s_skip;
  const int b = blockIdx.y;
  const int h = a.h, w = a.w, D = a.D, C = a.C;
  const unsigned plane = (unsigned)(h * w);
#if PPEA_MATCH_TILE_W == 0      // (tuning reference: 128 consecutive pixels of the flattened image)
  const unsigned lin = blockIdx.x * kMatchThreads + threadIdx.x;
  const int tx0 = (int)(lin % (unsigned)w), ty0 = (int)(lin / (unsigned)w);
#else
  const int tiles_x = (w + kMatchTileW - 1) / kMatchTileW;
  const int tx0 = (blockIdx.x % tiles_x) * kMatchTileW + (threadIdx.x % kMatchTileW), ty0 = (blockIdx.x / tiles_x) * kMatchTileH + (threadIdx.x / kMatchTileW);
#endif
  const bool live = tx0 < w && ty0 < h;
  const int y = live ? ty0 : 0, x = live ? tx0 : 0;
  const unsigned pix = (unsigned)y * (unsigned)w + (unsigned)x;
  if (threadIdx.x < 9) siK[threadIdx.x] = a.invK[b * 16 + (threadIdx.x / 3) * 4 + threadIdx.x % 3];
  __syncthreads();
  float ray[3];
  pixel_ray(siK, (float)x, (float)y, ray);
  const float cur_mask = (y >= 2 && y < h - 2 && x >= 2 && x < w - 2) ? 1.f : 0.f;      // current_mask[:, 2:-2, 2:-2] = 1 (:310-312)
  const float* cur_b = a.cur + (size_t)b * C * plane + pix;
  float* cost_b = a.cost + (size_t)b * D * plane + pix;
  float* miss_b = a.missing + (size_t)b * D * plane + pix;
  const float fC = (float)C;
  const int d_lo = blockIdx.z * kMatchChunk, d_hi = min(D, d_lo + kMatchChunk);

  for (int d0 = d_lo; d0 < d_hi; d0 += kMatchBins) {
    float cost[kMatchBins], cnt[kMatchBins];
#pragma unroll
    for (int j = 0; j < kMatchBins; ++j) cost[j] = cnt[j] = 0.f;
    for (int f = 0; f < a.F; ++f) {
      __syncthreads();
      if (threadIdx.x == 0) {
        const float* T = a.poses + ((size_t)b * a.F + f) * 16;
        float s = 0.f;
        for (int e = 0; e < 16; ++e) s += T[e];
        s_skip = (s == 0.f) ? 1 : 0;                      // "missing lookup frame" (:289-291)
        compose_P(a.K + b * 16, T, sP);
      }
      __syncthreads();
      if (s_skip) continue;
      MatchTap tap[kMatchBins];
      float acc[kMatchBins];
#pragma unroll
      for (int j = 0; j < kMatchBins; ++j) {
        const int d = min(d0 + j, D - 1);
        tap[j] = match_setup(sP, ray, __ldg(a.bins + d), h, w, a.eps, cur_mask);
        acc[j] = 0.f;
      }
      if (live) {
        const float* look_f = a.look + ((size_t)b * a.F + f) * C * plane;
#pragma unroll 2
        for (int c = 0; c < C; ++c) {
          const float cv = __ldg(cur_b + (size_t)c * plane);
          const float* lp = look_f + (size_t)c * plane;
#pragma unroll
          for (int j = 0; j < kMatchBins; ++j) {
            const float* q = lp + tap[j].off;
            const float v = fmaf(tap[j].wse, __ldg(q + w + 1), fmaf(tap[j].wsw, __ldg(q + w), fmaf(tap[j].wne, __ldg(q + 1), tap[j].wnw * __ldg(q))));
            acc[j] += fabsf(v - cv);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kMatchBins; ++j) {
        const float diff = mul_rn(div_rn(acc[j], fC), tap[j].edge);        // .mean(1) * edge_mask (:314-315)
        cost[j] += diff;
        cnt[j] += diff > 0.f ? 1.f : 0.f;
      }
    }
    if (live) {
#pragma unroll
      for (int j = 0; j < kMatchBins; ++j) {
        const int d = d0 + j;
        if (d < d_hi) {
          const float v = cost[j] / (cnt[j] + 1e-7f);             // average over lookup images (:321)
          cost_b[(size_t)d * plane] = v;
          miss_b[(size_t)d * plane] = (v == 0.f) ? 1.f : 0.f;     // missing_val_mask (:324)
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Channel-quad variant (the path ppea_match_features_ws takes when C % 4 == 0).  The planar kernel issues one 32-bit
// load per channel and corner (21 instructions per hypothesis and channel, a warp's 33-float corner row straddles two
// 128-byte lines).  Here both feature tensors are first re-laid (match_repack_kernel, one pass over 2 x 23.6 MB at the
// KITTI shape, inside the timed call) as (N, C/4, h, w, 4): the four channels of a cell are one aligned 16-byte word,
// adjacent pixels are adjacent words, so a corner of a hypothesis is ONE 128-bit load for four channels and a warp's
// corner row is 33 words = 528 contiguous bytes.  Same arithmetic in the same channel order as the planar kernel: the
// two produce bit-identical volumes (tests/test_matching.py).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) match_repack_kernel(const float* __restrict__ cur, const float* __restrict__ look, float4* __restrict__ out,
                                                           unsigned plane, size_t n_cur, size_t total) {
  // one launch for both tensors: output word i = (n * C/4 + group) * plane + pixel, the current features first
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const bool second = i >= n_cur;
  const size_t k = second ? i - n_cur : i;
  const size_t g = k / plane;
  const unsigned pix = (unsigned)(k - g * plane);
  const float* p = (second ? look : cur) + g * 4 * (size_t)plane + pix;
  out[i] = make_float4(__ldg(p), __ldg(p + plane), __ldg(p + 2 * (size_t)plane), __ldg(p + 3 * (size_t)plane));
}

#ifndef PPEA_MATCHQ_BINS
#define PPEA_MATCHQ_BINS 4
#endif
#ifndef PPEA_MATCHQ_CTAS
#define PPEA_MATCHQ_CTAS 4
#endif
#ifndef PPEA_MATCHQ_CHUNK
#define PPEA_MATCHQ_CHUNK 4
#endif
constexpr int kMatchQBins = PPEA_MATCHQ_BINS;      // depth hypotheses in flight per thread (16 x 128-bit loads at four)
constexpr int kMatchQChunk = PPEA_MATCHQ_CHUNK;    // depth bins per CTA (measured: 4 -> 0.484 ms, 8 -> 0.513, 16 -> 0.530 at the KITTI shape)

__global__ void __launch_bounds__(kMatchThreads, PPEA_MATCHQ_CTAS) match_features_quad_kernel(const MatchArgs a, const float4* __restrict__ cur_q,
                                                                                              const float4* __restrict__ look_q) {
  __shared__ float sP[12];         // (K @ T)[:3,:] of the current lookup frame
  __shared__ float siK[9];
  __shared__ int s_skip;
  const int b = blockIdx.y;
  const int h = a.h, w = a.w, D = a.D, C4 = a.C >> 2;
  const unsigned plane = (unsigned)(h * w);
  const int tiles_x = (w + 31) / 32;
  const int tx0 = (blockIdx.x % tiles_x) * 32 + (threadIdx.x & 31), ty0 = (blockIdx.x / tiles_x) * (kMatchThreads / 32) + (threadIdx.x >> 5);
  const bool live = tx0 < w && ty0 < h;
  const int y = live ? ty0 : 0, x = live ? tx0 : 0;
  const unsigned pix = (unsigned)y * (unsigned)w + (unsigned)x;
  if (threadIdx.x < 9) siK[threadIdx.x] = a.invK[b * 16 + (threadIdx.x / 3) * 4 + threadIdx.x % 3];
  __syncthreads();
  float ray[3];
  pixel_ray(siK, (float)x, (float)y, ray);
  const float cur_mask = (y >= 2 && y < h - 2 && x >= 2 && x < w - 2) ? 1.f : 0.f;      // current_mask[:, 2:-2, 2:-2] = 1 (:310-312)
  const float4* cur_p = cur_q + (size_t)b * C4 * plane + pix;
  float* cost_b = a.cost + (size_t)b * D * plane + pix;
  float* miss_b = a.missing + (size_t)b * D * plane + pix;
  const float fC = (float)a.C;
  const int d_lo = blockIdx.z * kMatchQChunk, d_hi = min(D, d_lo + kMatchQChunk);

  for (int d0 = d_lo; d0 < d_hi; d0 += kMatchQBins) {
    float cost[kMatchQBins], cnt[kMatchQBins];
#pragma unroll
    for (int j = 0; j < kMatchQBins; ++j) cost[j] = cnt[j] = 0.f;
    for (int f = 0; f < a.F; ++f) {
      __syncthreads();
      if (threadIdx.x == 0) {
        const float* T = a.poses + ((size_t)b * a.F + f) * 16;
        float s = 0.f;
        for (int e = 0; e < 16; ++e) s += T[e];
        s_skip = (s == 0.f) ? 1 : 0;                      // "missing lookup frame" (:289-291)
        compose_P(a.K + b * 16, T, sP);
      }
      __syncthreads();
      if (s_skip) continue;
      MatchTap tap[kMatchQBins];
      float acc[kMatchQBins];
#pragma unroll
      for (int j = 0; j < kMatchQBins; ++j) {
        const int d = min(d0 + j, D - 1);
        tap[j] = match_setup(sP, ray, __ldg(a.bins + d), h, w, a.eps, cur_mask);
        acc[j] = 0.f;
      }
      if (live) {
        const float4* look_f = look_q + ((size_t)b * a.F + f) * C4 * plane;
#pragma unroll 1
        for (int g = 0; g < C4; ++g) {
          const float4 cv = __ldg(cur_p + (size_t)g * plane);
          const float4* lp = look_f + (size_t)g * plane;
          float4 nw[kMatchQBins], ne[kMatchQBins], sw[kMatchQBins], se[kMatchQBins];
#pragma unroll
          for (int j = 0; j < kMatchQBins; ++j) {          // all loads of the group first: 4 x kMatchQBins requests in flight
            const float4* q = lp + tap[j].off;
            nw[j] = __ldg(q), ne[j] = __ldg(q + 1), sw[j] = __ldg(q + w), se[j] = __ldg(q + w + 1);
          }
#pragma unroll
          for (int j = 0; j < kMatchQBins; ++j) {
            const MatchTap& t = tap[j];
            const float v0 = fmaf(t.wse, se[j].x, fmaf(t.wsw, sw[j].x, fmaf(t.wne, ne[j].x, t.wnw * nw[j].x)));
            acc[j] += fabsf(v0 - cv.x);
            const float v1 = fmaf(t.wse, se[j].y, fmaf(t.wsw, sw[j].y, fmaf(t.wne, ne[j].y, t.wnw * nw[j].y)));
            acc[j] += fabsf(v1 - cv.y);
            const float v2 = fmaf(t.wse, se[j].z, fmaf(t.wsw, sw[j].z, fmaf(t.wne, ne[j].z, t.wnw * nw[j].z)));
            acc[j] += fabsf(v2 - cv.z);
            const float v3 = fmaf(t.wse, se[j].w, fmaf(t.wsw, sw[j].w, fmaf(t.wne, ne[j].w, t.wnw * nw[j].w)));
            acc[j] += fabsf(v3 - cv.w);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kMatchQBins; ++j) {
        const float diff = mul_rn(div_rn(acc[j], fC), tap[j].edge);        // .mean(1) * edge_mask (:314-315)
        cost[j] += diff;
        cnt[j] += diff > 0.f ? 1.f : 0.f;
      }
    }
    if (live) {
#pragma unroll
      for (int j = 0; j < kMatchQBins; ++j) {
        const int d = d0 + j;
        if (d < d_hi) {
          const float v = cost[j] / (cnt[j] + 1e-7f);             // average over lookup images (:321)
          cost_b[(size_t)d * plane] = v;
          miss_b[(size_t)d * plane] = (v == 0.f) ? 1.f : 0.f;     // missing_val_mask (:324)
        }
      }
    }
  }
}

// (A row-pair variant -- a thread owning two vertically adjacent pixels and re-using the upper pixel's south corners as the
// lower pixel's north corners where the footprints abut -- was measured bit-identical and slower, 0.505 ms against 0.482 ms:
// profiles/README.md; it lives in the history, commit "match_features: row-pair channel-quad variant".)

// cost = cost * (1 - missing) + max_d(cost) * missing   (:325-328): one thread per pixel, two sweeps over its bins
__global__ void __launch_bounds__(256) match_fill_missing_kernel(float* __restrict__ cost, int D, unsigned plane) {
  const unsigned pix = blockIdx.x * 256 + threadIdx.x;
  if (pix >= plane) return;
  float* c = cost + (size_t)blockIdx.y * D * plane + pix;
  float vmax = -INFINITY;
  bool any = false;
#pragma unroll 8
  for (int d = 0; d < D; ++d) {
    const float v = c[(size_t)d * plane];
    vmax = fmaxf(vmax, v);
    any = any || v == 0.f;
  }
  if (!any || vmax == 0.f) return;
  for (int d = 0; d < D; ++d)
    if (c[(size_t)d * plane] == 0.f) c[(size_t)d * plane] = vmax;
}

// The same fix-up with the bins of a pixel split over the four warps of a CTA (lane == pixel, warp == quarter of the bins, held in
// registers): four times the loads in flight, one read of the volume instead of two (the one-thread-per-pixel sweep above is
// latency-bound: 26 us for 35 MB, ncu r2k).  max is exact, so the split does not change a bit.  Used when D <= 4 * kFillMaxPer.
constexpr int kFillParts = 4, kFillMaxPer = 32;
__global__ void __launch_bounds__(32 * kFillParts) match_fill_missing_split_kernel(float* __restrict__ cost, int D, unsigned plane, int per) {
  __shared__ float s_max[kFillParts][32];
  __shared__ int s_any[kFillParts][32];
  const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
  const unsigned pix = blockIdx.x * 32 + lane;
  const bool live = pix < plane;
  float* c = cost + (size_t)blockIdx.y * D * plane + (live ? pix : 0u);
  const int d0 = part * per;
  float v[kFillMaxPer];
  float vmax = -INFINITY;
  bool any = false;
#pragma unroll
  for (int k = 0; k < kFillMaxPer; ++k) {
    const bool ok = live && k < per && d0 + k < D;
    v[k] = ok ? c[(size_t)(d0 + k) * plane] : -INFINITY;
  }
#pragma unroll
  for (int k = 0; k < kFillMaxPer; ++k) {
    vmax = fmaxf(vmax, v[k]);
    any = any || v[k] == 0.f;
  }
  s_max[part][lane] = vmax;
  s_any[part][lane] = any ? 1 : 0;
  __syncthreads();
  float m = s_max[0][lane];
  int a = s_any[0][lane];
#pragma unroll
  for (int q = 1; q < kFillParts; ++q) m = fmaxf(m, s_max[q][lane]), a |= s_any[q][lane];
  if (!any || !a || m == 0.f) return;          // (nothing of this thread's quarter to fill)
#pragma unroll
  for (int k = 0; k < kFillMaxPer; ++k)
    if (v[k] == 0.f) c[(size_t)(d0 + k) * plane] = m;
}

static cudaError_t launch_match_fill_missing(float* cost, int batch, int D, int height, int width, cudaStream_t stream) {
  const unsigned plane = (unsigned)(height * width);
  if (D <= kFillParts * kFillMaxPer) {
    const dim3 g((unsigned)ceil_div(height * width, 32), (unsigned)batch);
    match_fill_missing_split_kernel<<<g, 32 * kFillParts, 0, stream>>>(cost, D, plane, ceil_div(D, kFillParts));
  } else {
    const dim3 g((unsigned)ceil_div(height * width, 256), (unsigned)batch);
    match_fill_missing_kernel<<<g, 256, 0, stream>>>(cost, D, plane);
  }
  return cudaGetLastError();
}

// Tail of the encoder's matching block (replk_matching_adapter.py:380-387, :439-453), one sweep per pixel over its bins:
//   confidence = [ #(cost * (1 - missing) > 0) == threshold ]                       compute_confidence_mask
//   (mins, argmin) = min_d viz,  viz = cost with exact zeros replaced by 100          (first minimum, like torch.min)
//   cost *= confidence                                                                (optional, in place)
__global__ void __launch_bounds__(256) match_tail_kernel(float* __restrict__ cost, const float* __restrict__ missing,
                                                         float* __restrict__ confidence, float* __restrict__ mins,
                                                         long long* __restrict__ argmin, int D, unsigned plane, int threshold,
                                                         int mask_volume) {
  const unsigned pix = blockIdx.x * 256 + threadIdx.x;
  if (pix >= plane) return;
  const size_t base = (size_t)blockIdx.y * D * plane + pix;
  float* c = cost + base;
  const float* m = missing ? missing + base : nullptr;
  int count = 0, best = 0;
  float vmin = INFINITY;
#pragma unroll 4
  for (int d = 0; d < D; ++d) {
    const float v = c[(size_t)d * plane];
    const float keep = m ? mul_rn(v, sub_rn(1.f, __ldg(m + (size_t)d * plane))) : v;
    count += keep > 0.f ? 1 : 0;
    const float viz = (v == 0.f) ? 100.f : v;
    if (viz < vmin) {
      vmin = viz;
      best = d;
    }
  }
  const float conf = (count == threshold) ? 1.f : 0.f;
  const size_t o = (size_t)blockIdx.y * plane + pix;
  if (confidence) confidence[o] = conf;
  if (mins) mins[o] = vmin;
  if (argmin) argmin[o] = best;
  if (mask_volume && conf == 0.f)
    for (int d = 0; d < D; ++d) c[(size_t)d * plane] = mul_rn(c[(size_t)d * plane], 0.f);
}

// The same tail with the bins of a pixel split over the four warps of a CTA (like match_fill_missing_split_kernel): every warp
// holds its quarter of the bins in registers, the partial (count, min, first argmin) are combined in ascending bin order -- a
// later quarter only wins with a strictly smaller value, which is torch.min's first-minimum rule -- and the masking writes come
// from the registers instead of a second read.
__global__ void __launch_bounds__(32 * kFillParts) match_tail_split_kernel(float* __restrict__ cost, const float* __restrict__ missing,
                                                                           float* __restrict__ confidence, float* __restrict__ mins,
                                                                           long long* __restrict__ argmin, int D, unsigned plane, int threshold,
                                                                           int mask_volume, int per) {
  __shared__ float s_min[kFillParts][32];
  __shared__ int s_cnt[kFillParts][32], s_best[kFillParts][32];
  const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
  const unsigned pix = blockIdx.x * 32 + lane;
  const bool live = pix < plane;
  const size_t base = (size_t)blockIdx.y * D * plane + (live ? pix : 0u);
  float* c = cost + base;
  const float* m = missing ? missing + base : nullptr;
  const int d0 = part * per;
  float v[kFillMaxPer], keep[kFillMaxPer];
#pragma unroll
  for (int k = 0; k < kFillMaxPer; ++k) {
    const bool ok = live && k < per && d0 + k < D;
    v[k] = ok ? c[(size_t)(d0 + k) * plane] : INFINITY;
    keep[k] = (ok && m) ? __ldg(m + (size_t)(d0 + k) * plane) : 0.f;
  }
  int count = 0, best = 0;
  float vmin = INFINITY;
#pragma unroll
  for (int k = 0; k < kFillMaxPer; ++k) {
    const bool ok = live && k < per && d0 + k < D;
    const float kv = m ? mul_rn(v[k], sub_rn(1.f, keep[k])) : v[k];
    count += (ok && kv > 0.f) ? 1 : 0;
    const float viz = (v[k] == 0.f) ? 100.f : v[k];
    if (ok && viz < vmin) {
      vmin = viz;
      best = d0 + k;
    }
  }
  s_min[part][lane] = vmin;
  s_cnt[part][lane] = count;
  s_best[part][lane] = best;
  __syncthreads();
  float gmin = s_min[0][lane];
  int gcnt = s_cnt[0][lane], gbest = s_best[0][lane];
#pragma unroll
  for (int q = 1; q < kFillParts; ++q) {
    gcnt += s_cnt[q][lane];
    if (s_min[q][lane] < gmin) gmin = s_min[q][lane], gbest = s_best[q][lane];
  }
  if (!live) return;
  const float conf = (gcnt == threshold) ? 1.f : 0.f;
  if (part == 0) {
    const size_t o = (size_t)blockIdx.y * plane + pix;
    if (confidence) confidence[o] = conf;
    if (mins) mins[o] = gmin;
    if (argmin) argmin[o] = gbest;
  }
  if (mask_volume && conf == 0.f) {
#pragma unroll
    for (int k = 0; k < kFillMaxPer; ++k)
      if (k < per && d0 + k < D) c[(size_t)(d0 + k) * plane] = mul_rn(v[k], 0.f);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// match_features_dyn (replk_matching_adapter.py:163-258), the variant the encoder takes when it is handed a teacher depth
// (:400, :439-442): the same plane sweep, plus (a) an occlusion map of the lookup image projected into every layer of the
// volume -- where it exceeds pool_th, and the item is not augmented, the warped features become 1 (set_1) or the 3-D max
// over the (2 pool_r + 1)^3 neighbourhood of the un-occluded warped features (pool, F.max_pool3d :206, out-of-volume
// neighbours ignored like its -inf padding) -- and (b) the lookup frames combined by minimum instead of the average
// (cv_min :238-246).  The reference materialises the (D,C,h,w) warped volume to pool it; here an occluded cell re-derives
// the samples of its neighbours (their own projections, their own occlusion decision) on the fly -- occluded cells are few.
// A thread owns one pixel and walks the depth bins of its CTA's chunk one at a time.
// ---------------------------------------------------------------------------------------------------------------
struct MatchDynArgs {
  MatchArgs m;
  const float* occ;       // (N,h,w) 0/1 occlusion map of the lookup images at the matching resolution, indexed by batch item (:166, :198)
  const float* aug;       // (B) augmentation mask (item skipped when != 0, :196)
  int cv_min, set_1, pool, pool_r;
  float pool_th;
};

__device__ __forceinline__ float match_sample(const float* __restrict__ plane_ptr, const MatchTap& t, int w) {
  const float* q = plane_ptr + t.off;
  return fmaf(t.wse, __ldg(q + w + 1), fmaf(t.wsw, __ldg(q + w), fmaf(t.wne, __ldg(q + 1), t.wnw * __ldg(q))));
}

constexpr int kMatchDynMaxR = 2;

__global__ void __launch_bounds__(kMatchThreads) match_features_dyn_kernel(const MatchDynArgs da) {
  const MatchArgs& a = da.m;
  __shared__ float sP[12];
  __shared__ float siK[9];
  __shared__ int s_skip;
  const int b = blockIdx.y;
  const int h = a.h, w = a.w, D = a.D, C = a.C;
  const unsigned plane = (unsigned)(h * w);
#if PPEA_MATCH_TILE_W == 0      // (tuning reference: 128 consecutive pixels of the flattened image)
  const unsigned lin = blockIdx.x * kMatchThreads + threadIdx.x;
  const int tx0 = (int)(lin % (unsigned)w), ty0 = (int)(lin / (unsigned)w);
#else
  const int tiles_x = (w + kMatchTileW - 1) / kMatchTileW;
  const int tx0 = (blockIdx.x % tiles_x) * kMatchTileW + (threadIdx.x % kMatchTileW), ty0 = (blockIdx.x / tiles_x) * kMatchTileH + (threadIdx.x / kMatchTileW);
#endif
  const bool live = tx0 < w && ty0 < h;
  const int y = live ? ty0 : 0, x = live ? tx0 : 0;
  const unsigned pix = (unsigned)y * (unsigned)w + (unsigned)x;
  if (threadIdx.x < 9) siK[threadIdx.x] = a.invK[b * 16 + (threadIdx.x / 3) * 4 + threadIdx.x % 3];
  __syncthreads();
  float ray[3];
  pixel_ray(siK, (float)x, (float)y, ray);
  const float cur_mask = (y >= 2 && y < h - 2 && x >= 2 && x < w - 2) ? 1.f : 0.f;
  const float* cur_b = a.cur + (size_t)b * C * plane + pix;
  float* cost_b = a.cost + (size_t)b * D * plane + pix;
  float* miss_b = a.missing + (size_t)b * D * plane + pix;
  const float fC = (float)C;
  const bool occl = (da.set_1 || da.pool) && (__ldg(da.aug + b) == 0.f);
  const float* occ_b = da.occ + (size_t)b * plane;
  const int r = da.pool_r;
  const int d_lo = blockIdx.z * kMatchChunk, d_hi = min(D, d_lo + kMatchChunk);

  for (int d = d_lo; d < d_hi; ++d) {
    float cost = da.cv_min ? 1.f : 0.f, cnt = 0.f;
    for (int f = 0; f < a.F; ++f) {
      __syncthreads();
      if (threadIdx.x == 0) {
        const float* T = a.poses + ((size_t)b * a.F + f) * 16;
        float s = 0.f;
        for (int e = 0; e < 16; ++e) s += T[e];
        s_skip = (s == 0.f) ? 1 : 0;
        compose_P(a.K + b * 16, T, sP);
      }
      __syncthreads();
      if (s_skip) continue;
      const MatchTap tap = match_setup(sP, ray, __ldg(a.bins + d), h, w, a.eps, cur_mask);
      float acc = 0.f;
      if (live) {
        const float* look_f = a.look + ((size_t)b * a.F + f) * C * plane;
        const bool masked = occl && match_sample(occ_b, tap, w) > da.pool_th;
        if (!masked) {
          for (int c = 0; c < C; ++c) acc += fabsf(match_sample(look_f + (size_t)c * plane, tap, w) - __ldg(cur_b + (size_t)c * plane));
        } else if (da.set_1) {
          for (int c = 0; c < C; ++c) acc += fabsf(1.f - __ldg(cur_b + (size_t)c * plane));
        } else {
          // pooled: the un-occluded neighbours' own samples (the occluded ones, this cell included, count as 0)
          MatchTap nb[(2 * kMatchDynMaxR + 1) * (2 * kMatchDynMaxR + 1) * (2 * kMatchDynMaxR + 1)];
          int n_nb = 0;
          bool any_zero = false;         // some in-volume neighbour is occluded (contributes the value 0), e.g. this cell
          for (int dd = -r; dd <= r; ++dd)
            for (int dy = -r; dy <= r; ++dy)
              for (int dx = -r; dx <= r; ++dx) {
                const int d2 = d + dd, y2 = y + dy, x2 = x + dx;
                if (d2 < 0 || d2 >= D || y2 < 0 || y2 >= h || x2 < 0 || x2 >= w) continue;
                float ray2[3];
                pixel_ray(siK, (float)x2, (float)y2, ray2);
                const MatchTap t2 = match_setup(sP, ray2, __ldg(a.bins + d2), h, w, a.eps, 1.f);
                if (match_sample(occ_b, t2, w) > da.pool_th)
                  any_zero = true;
                else
                  nb[n_nb++] = t2;
              }
          for (int c = 0; c < C; ++c) {
            const float* lp = look_f + (size_t)c * plane;
            float vmax = any_zero ? 0.f : -INFINITY;
            for (int k = 0; k < n_nb; ++k) vmax = fmaxf(vmax, match_sample(lp, nb[k], w));
            acc += fabsf(vmax - __ldg(cur_b + (size_t)c * plane));
          }
        }
      }
      float diff = mul_rn(div_rn(acc, fC), tap.edge);
      if (da.cv_min) {
        if (diff == 0.f) diff = 1.f;                       // :238-240
        cost = fminf(diff, cost);
      } else {
        cost += diff;
        cnt += diff > 0.f ? 1.f : 0.f;
      }
    }
    if (live) {
      const float v = da.cv_min ? (cost == 1.f ? 0.f : cost) : cost / (cnt + 1e-7f);     // :245-248
      cost_b[(size_t)d * plane] = v;
      miss_b[(size_t)d * plane] = (v == 0.f) ? 1.f : 0.f;
    }
  }
}

extern "C" int ppea_match_features_dyn(const float* current_feats, const float* lookup_feats, const float* relative_poses, const float* K,
                                       const float* inv_K, const float* depth_bins, const float* occlusion, const float* aug_mask,
                                       float* cost_volume, float* missing_mask, int batch, int num_lookup, int channels, int height,
                                       int width, int num_bins, int set_missing_to_max, int cv_min, int set_1, int pool, int pool_radius,
                                       float pool_threshold, float eps, void* stream) {
  if (!current_feats || !lookup_feats || !relative_poses || !K || !inv_K || !depth_bins || !cost_volume || !missing_mask) return PPEA_E_NULL;
  if ((set_1 || pool) && (!occlusion || !aug_mask)) return PPEA_E_FLAGS;
  if (batch <= 0 || batch > 65535 || num_lookup < 0 || channels <= 0 || height < 2 || width < 2 || num_bins <= 0 || num_bins > kMatchMaxBins ||
      ceil_div(num_bins, kMatchChunk) > 65535 || (long long)height * width >= (1ll << 30) || pool_radius < 0 || pool_radius > kMatchDynMaxR)
    return PPEA_E_SHAPE;
  MatchDynArgs da;
  MatchArgs& a = da.m;
  a.cur = current_feats;
  a.look = lookup_feats;
  a.poses = relative_poses;
  a.K = K;
  a.invK = inv_K;
  a.bins = depth_bins;
  a.cost = cost_volume;
  a.missing = missing_mask;
  a.B = batch;
  a.F = num_lookup;
  a.C = channels;
  a.h = height;
  a.w = width;
  a.D = num_bins;
  a.set_missing_to_max = set_missing_to_max;
  a.eps = eps;
  da.occ = occlusion;
  da.aug = aug_mask;
  da.cv_min = cv_min;
  da.set_1 = set_1;
  da.pool = set_1 ? 0 : pool;          // (`if set_1 ... elif pool`, :202-205)
  da.pool_r = pool_radius;
  da.pool_th = pool_threshold;
  const dim3 grid((unsigned)match_tiles(height, width), (unsigned)batch, (unsigned)ceil_div(num_bins, kMatchChunk));
  match_features_dyn_kernel<<<grid, kMatchThreads, 0, (cudaStream_t)stream>>>(da);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  if (set_missing_to_max) e = launch_match_fill_missing(cost_volume, batch, num_bins, height, width, (cudaStream_t)stream);
  return (int)e;
}

extern "C" int ppea_match_tail(float* cost_volume, const float* missing_mask_or_null, float* confidence_or_null, float* mins_or_null,
                               long long* argmin_or_null, int batch, int num_bins, int height, int width, int threshold,
                               int mask_volume, void* stream) {
  if (!cost_volume) return PPEA_E_NULL;
  if (batch <= 0 || batch > 65535 || num_bins <= 0 || height <= 0 || width <= 0 || (long long)height * width >= (1ll << 30)) return PPEA_E_SHAPE;
  if (num_bins <= kFillParts * kFillMaxPer) {
    const dim3 grid((unsigned)ceil_div(height * width, 32), (unsigned)batch);
    match_tail_split_kernel<<<grid, 32 * kFillParts, 0, (cudaStream_t)stream>>>(cost_volume, missing_mask_or_null, confidence_or_null, mins_or_null,
                                                                                argmin_or_null, num_bins, (unsigned)(height * width), threshold,
                                                                                mask_volume, ceil_div(num_bins, kFillParts));
  } else {
    const dim3 grid((unsigned)ceil_div(height * width, 256), (unsigned)batch);
    match_tail_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(cost_volume, missing_mask_or_null, confidence_or_null, mins_or_null, argmin_or_null,
                                                           num_bins, (unsigned)(height * width), threshold, mask_volume);
  }
  return (int)cudaGetLastError();
}

// bytes of the channel-quad copies of current_feats and lookup_feats (0: this shape takes the planar kernel)
extern "C" size_t ppea_match_workspace_bytes(int batch, int num_lookup, int channels, int height, int width) {
  if (batch <= 0 || num_lookup <= 0 || channels <= 0 || (channels & 3) || height < 2 || width < 2) return 0;
  return ((size_t)batch * (1 + (size_t)num_lookup) * channels * height * width) * sizeof(float);
}

extern "C" int ppea_match_features_ws(const float* current_feats, const float* lookup_feats, const float* relative_poses, const float* K,
                                      const float* inv_K, const float* depth_bins, float* cost_volume, float* missing_mask, int batch,
                                      int num_lookup, int channels, int height, int width, int num_bins, int set_missing_to_max, float eps,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  if (!current_feats || !lookup_feats || !relative_poses || !K || !inv_K || !depth_bins || !cost_volume || !missing_mask) return PPEA_E_NULL;
  if (batch <= 0 || batch > 65535 || num_lookup < 0 || channels <= 0 || height < 2 || width < 2 || num_bins <= 0 || num_bins > kMatchMaxBins || ceil_div(num_bins, kMatchChunk) > 65535 ||
      ceil_div(num_bins, kMatchQChunk) > 65535 || (long long)height * width >= (1ll << 30))
    return PPEA_E_SHAPE;
  MatchArgs a;
  a.cur = current_feats;
  a.look = lookup_feats;
  a.poses = relative_poses;
  a.K = K;
  a.invK = inv_K;
  a.bins = depth_bins;
  a.cost = cost_volume;
  a.missing = missing_mask;
  a.B = batch;
  a.F = num_lookup;
  a.C = channels;
  a.h = height;
  a.w = width;
  a.D = num_bins;
  a.set_missing_to_max = set_missing_to_max;
  a.eps = eps;
  cudaError_t e;
  const size_t need = ppea_match_workspace_bytes(batch, num_lookup, channels, height, width);
  if (workspace && need > 0) {
    if (workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 15)) return PPEA_E_WORKSPACE;
    const unsigned plane = (unsigned)(height * width);
    float4* cur_q = reinterpret_cast<float4*>(workspace);
    float4* look_q = cur_q + (size_t)batch * (channels / 4) * plane;
    const size_t n_cur = (size_t)batch * (channels / 4) * plane, n_look = n_cur * num_lookup;
    match_repack_kernel<<<(unsigned)((n_cur + n_look + 255) / 256), 256, 0, (cudaStream_t)stream>>>(current_feats, lookup_feats, cur_q, plane, n_cur,
                                                                                                     n_cur + n_look);
    const dim3 grid((unsigned)(ceil_div(width, 32) * ceil_div(height, kMatchThreads / 32)), (unsigned)batch, (unsigned)ceil_div(num_bins, kMatchQChunk));
    match_features_quad_kernel<<<grid, kMatchThreads, 0, (cudaStream_t)stream>>>(a, cur_q, look_q);
  } else {
    const dim3 grid((unsigned)match_tiles(height, width), (unsigned)batch, (unsigned)ceil_div(num_bins, kMatchChunk));
    match_features_kernel<<<grid, kMatchThreads, 0, (cudaStream_t)stream>>>(a);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  if (set_missing_to_max) e = launch_match_fill_missing(cost_volume, batch, num_bins, height, width, (cudaStream_t)stream);
  return (int)e;
}

extern "C" int ppea_match_features(const float* current_feats, const float* lookup_feats, const float* relative_poses, const float* K,
                                   const float* inv_K, const float* depth_bins, float* cost_volume, float* missing_mask, int batch,
                                   int num_lookup, int channels, int height, int width, int num_bins, int set_missing_to_max, float eps,
                                   void* stream) {
  return ppea_match_features_ws(current_feats, lookup_feats, relative_poses, K, inv_K, depth_bins, cost_volume, missing_mask, batch, num_lookup,
                                channels, height, width, num_bins, set_missing_to_max, eps, nullptr, 0, stream);
}

}  // namespace ppea
