// vsl_fwd.cu -- fused forward of the view-synthesis loss, all pyramid scales in one launch.
//
// One launch replaces, for every scale, trainer.py:886-914 (upsample, depth, backproject,
// project, grid_sample for both sources) and trainer.py:1050-1141 (SSIM+L1 photometric loss of
// the warped and of the un-warped sources, min over sources, selec_reproj, identity automask /
// multi-frame mask, masked sums).
//
// A CTA owns a TW x TH tile of one image.  Shared memory holds the tile plus a one-pixel
// reflection halo of the target (3 planes, each value duplicated into both lanes of an f2) and
// of the two sources (3 planes of f2 = (source -1, source +1)): first the un-warped sources
// (identity loss, once for all scales), then, per scale, the warped ones.  All photometric
// arithmetic runs 2-wide (FFMA2/FADD2/FMUL2, lanes = sources).  Every 3x3 SSIM window is
// assembled by a thread that walks R consecutive rows of one column with a three-row sliding
// window of horizontal sums held in registers.
// HBM traffic per pixel: tgt 12 B + sources 24 B once per CTA, per scale the gathers (24 B
// compulsory, L1/L2-resident after the first scale) + noise 4 B + disp 4/4^s B in; depth 4 B +
// sel 1 B (+ loss 4 B on request) out.
#include "vsl_common.cuh"
#include "smooth.cuh"
#include "vsl_gather.cuh"

namespace ppea {

template <int TW, int TH>
struct FwdSmem {
  static constexpr int EW = TW + 2;
  static constexpr int EH = TH + 2;
  static constexpr int PLANE = EW * EH;
  float y[3][PLANE];   // target (broadcast into both lanes at use: FFMA2 takes a scalar .F32 operand)
  f2 x[3][PLANE];      // (source 0, source 1): un-warped, then warped
  f2 G[12];            // per-source geometry M[0..8], t[0..2] (vsl_math.cuh Geom), lanes = sources
  float red[3][8];
};

// Photometric loss 0.85*mean_c SSIM + 0.15*mean_c |y-x| (trainer.py:995-1007) of both sources
// against the target for the R pixels (rows row0..row0+R-1, column col) of this thread.
template <int R, int EW, int PLANE, bool WITH_CSUM>
__device__ __forceinline__ void photometric_pass(const f2* __restrict__ xs, const float* __restrict__ ys, int row0, int col,
                                                 bool no_ssim, f2 (&acc)[R], f2 (&cs)[R]) {
#pragma unroll
  for (int k = 0; k < R; ++k) {
    acc[k] = dup2(0.f);
    cs[k] = dup2(0.f);
  }
  const float w_l1 = no_ssim ? (1.f / 3.f) : PPEA_W_L1;
  const f2 w_ssim = dup2(PPEA_W_SSIM);
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
    const float* yp = ys + c * PLANE + row0 * EW + col;
    const f2* xp = xs + c * PLANE + row0 * EW + col;
    f2 hy[3], hyy[3], hx[3], hxx[3], hxy[3];
#pragma unroll
    for (int i = 0; i < R + 2; ++i) {
      const int s = i % 3;
      const f2 y0 = dup2(yp[i * EW]), y1 = dup2(yp[i * EW + 1]), y2 = dup2(yp[i * EW + 2]);
      const f2 x0 = xp[i * EW], x1 = xp[i * EW + 1], x2 = xp[i * EW + 2];
      row_sums_y<f2>(y0, y1, y2, hy[s], hyy[s]);
      row_sums_x<f2>(x0, x1, x2, y0, y1, y2, hx[s], hxx[s], hxy[s]);
      if (i >= 1 && i <= R) {
        const f2 d = vsub(y1, x1);
        acc[i - 1].x = fma_rn(w_l1, fabsf(d.x), acc[i - 1].x);
        acc[i - 1].y = fma_rn(w_l1, fabsf(d.y), acc[i - 1].y);
        if (WITH_CSUM) cs[i - 1] = vadd(cs[i - 1], x1);
      }
      if (i >= 2 && !no_ssim) {
        const SsimYT<f2> yst = ssim_y_stats<f2>(sum3(hy[0], hy[1], hy[2]), sum3(hyy[0], hyy[1], hyy[2]));
        const f2 S = ssim_from_sums<f2>(sum3(hx[0], hx[1], hx[2]), sum3(hxx[0], hxx[1], hxx[2]), sum3(hxy[0], hxy[1], hxy[2]), yst);
        acc[i - 2] = vfma(w_ssim, S, acc[i - 2]);
      }
    }
  }
}

// Gather pass of one scale: fills the (source 0, source 1) planes of the tile + 1-pixel halo and
// writes the tile's depth.  Warp w walks region rows w, w+nw, ... with lane == tile column; the two
// halo columns are a flat list of extra cells taken by the upper warps.
template <int TW, int TH, int NT, bool SAME_RES, class Smem>
__device__ __forceinline__ void fwd_gather(Smem& sm, const VslArgs& a, const ScaleArgs& sc, const float* __restrict__ disp_b,
                                           const SrcPlanes& sp, float* __restrict__ depth_b,
                                           int x0, int y0, int col, int wid, int tid, ColCtx cc, float wmax, float hmax) {
  constexpr int EW = Smem::EW, EH = Smem::EH, NW = NT / 32;
  const int H = a.H, W = a.W;
  const int gx_own = x0 + col;
  f2 val[3], unused0[3], unused1[3];
  if (!SAME_RES) cc.cx = up_coef(cc.px, sc.ws, sc.up_sx);
  // contiguous block of region rows per warp: a row's south corners are the next row's north corners (L1 reuse)
  constexpr int RPW = (EH + NW - 1) / NW;
  const int i_lo = wid * RPW, i_hi = (i_lo + RPW < EH) ? i_lo + RPW : EH;
  float dnext = 0.f;
  if (SAME_RES && i_lo < EH) dnext = __ldg(disp_b + ((unsigned)reflect_index(y0 - 1 + i_lo, H) * (unsigned)W + (unsigned)cc.px));
  for (int i = i_lo; i < i_hi; ++i) {
    const int gy = y0 - 1 + i, py = reflect_index(gy, H);
    UpCoef cy;
    float dep;
    if (SAME_RES) {
      const float dcur = dnext;
      if (i + 1 < i_hi) dnext = __ldg(disp_b + ((unsigned)reflect_index(gy + 1, H) * (unsigned)W + (unsigned)cc.px));
      dep = depth_from_disp_fast(dcur, a.disp_lo, a.disp_range);
    } else {
      cy = up_coef(py, sc.hs, sc.up_sy);
      dep = depth_of<false>(disp_b, W, sc.ws, py, cc, cy, a.disp_lo, a.disp_range);
    }
    if (i >= 1 && i <= TH && gy < H && gx_own < W) depth_b[(unsigned)gy * (unsigned)W + (unsigned)gx_own] = dep;
    f2 A[3];
    const ProjT<f2> pr = project_cell(sm.G, cc, py, dep, a.eps, wmax, hmax, A);
    sample_sources<false, true>(sp, W, pr, 0.f, 0.f, val, unused0, unused1);
    const int idx = i * EW + col + 1;
#pragma unroll
    for (int c = 0; c < 3; ++c) sm.x[c][idx] = val[c];
  }
  const int e = tid - (NT - 2 * 32);        // extra cells go to the last two warps (the last warp owns fewer rows)
  if (e >= 0 && e < 2 * EH) {
    const int i = e >> 1, j = (e & 1) ? EW - 1 : 0;
    const int py = reflect_index(y0 - 1 + i, H);
    ColCtx ce = make_col(sm.G, x0 - 1 + j, W);
    UpCoef cy;
    if (!SAME_RES) {
      ce.cx = up_coef(ce.px, sc.ws, sc.up_sx);
      cy = up_coef(py, sc.hs, sc.up_sy);
    }
    const float dep = depth_of<SAME_RES>(disp_b, W, sc.ws, py, ce, cy, a.disp_lo, a.disp_range);
    f2 A[3];
    const ProjT<f2> pr = project_cell(sm.G, ce, py, dep, a.eps, wmax, hmax, A);
    sample_sources<false, false>(sp, W, pr, 0.f, 0.f, val, unused0, unused1);
    const int idx = i * EW + j;
#pragma unroll
    for (int c = 0; c < 3; ++c) sm.x[c][idx] = val[c];
  }
}

#ifndef PPEA_FWD_CTAS
#define PPEA_FWD_CTAS 5
#endif
template <int TW, int TH, int NT>
__global__ void __launch_bounds__(NT, PPEA_FWD_CTAS) vsl_forward_kernel(const __grid_constant__ VslArgs a) {
  using Smem = FwdSmem<TW, TH>;
  constexpr int EW = Smem::EW, PLANE = Smem::PLANE;
  constexpr int R = (TW * TH) / NT;
  static_assert(NT % TW == 0 && (TW * TH) % NT == 0, "tile/thread mismatch");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

  // The first CTAs of the grid run the smoothness term of every scale (smooth.cuh): scheduled first, their
  // latency-bound work hides among the first wave of tile CTAs instead of stretching the tail.
  const int n_smooth = a.S * a.B * kSmoothChunks;
  if ((int)blockIdx.x < n_smooth) {
    smooth_forward_role(a, blockIdx.x, reinterpret_cast<float*>(smem_raw));
    return;
  }
  const int tid = threadIdx.x;
  int blk = blockIdx.x - n_smooth;
  const int tile_id = blk;
  const int tx = blk % a.tiles_x;
  blk /= a.tiles_x;
  const int ty = blk % a.tiles_y;
  const int b = blk / a.tiles_y;
  const int x0 = tx * TW, y0 = ty * TH;
  const int H = a.H, W = a.W;
  const size_t plane = (size_t)H * W;
  const bool multi = a.flags & PPEA_F_MULTI;
  const bool automask = (a.flags & PPEA_F_AUTOMASK) && !multi;
  const bool no_ssim = a.flags & PPEA_F_NO_SSIM;

  if (tid < 24) {
    const int f = tid / 12, e = tid % 12;
    const float v = geom_entry(a.K + b * 16, a.T[f] + b * 16, a.inv_K + b * 16, e);
    (f ? sm.G[e].y : sm.G[e].x) = v;
  }

  const float* tgt_b = a.tgt + (size_t)b * 3 * plane;
  const float* src_b[2] = {a.src[0] + (size_t)b * 3 * plane, a.src[1] + (size_t)b * 3 * plane};

  // ---- stage the target (+ un-warped sources for the identity loss) with the reflection halo
  for (int idx = tid; idx < PLANE; idx += NT) {
    const int i = idx / EW, j = idx - i * EW;
    const int py = reflect_index(y0 - 1 + i, H), px = reflect_index(x0 - 1 + j, W);
    const size_t o = (size_t)py * W + px;
#pragma unroll
    for (int c = 0; c < 3; ++c) sm.y[c][idx] = __ldg(tgt_b + c * plane + o);
    if (automask) {
#pragma unroll
      for (int c = 0; c < 3; ++c) sm.x[c][idx] = mk2(__ldg(src_b[0] + c * plane + o), __ldg(src_b[1] + c * plane + o));
    }
  }
  __syncthreads();

  const int col = tid % TW;
  const int row0 = (tid / TW) * R;
  float ident[R];
  f2 acc[R], cs[R];
  if (automask) {
    photometric_pass<R, EW, PLANE, false>(&sm.x[0][0], &sm.y[0][0], row0, col, no_ssim, acc, cs);
#pragma unroll
    for (int k = 0; k < R; ++k) ident[k] = fminf(acc[k].x, acc[k].y);   // trainer.py:1069
    __syncthreads();
  }

  const float wmax = coord_max(W), hmax = coord_max(H);
  const int gx_own = x0 + col;
  const float one_minus_aug = (multi && (a.flags & PPEA_F_MATCH_AUG)) ? 1.f - a.aug_mask[b] : 1.f;
  const int lane = tid & 31, wid = tid >> 5;
  const ColCtx col_own = make_col(sm.G, gx_own, W);     // this thread's tile column == region column col + 1
  const SrcPlanes sp = make_planes(src_b[0], src_b[1], plane);

#pragma unroll 1
  for (int s = 0; s < a.S; ++s) {
    const ScaleArgs& sc = a.sc[s];
    // ---- gather pass: depth, projection into both sources, bilinear samples -> x planes
    {
      const float* disp_b = sc.disp + (size_t)b * sc.hs * sc.ws;
      float* depth_b = sc.depth + (size_t)b * plane;
      if (sc.hs == H && sc.ws == W)
        fwd_gather<TW, TH, NT, true>(sm, a, sc, disp_b, sp, depth_b, x0, y0, col, wid, tid, col_own, wmax, hmax);
      else
        fwd_gather<TW, TH, NT, false>(sm, a, sc, disp_b, sp, depth_b, x0, y0, col, wid, tid, col_own, wmax, hmax);
    }
    __syncthreads();

    // epilogue inputs come from DRAM: issue their loads before the (long) photometric pass
    float pre[R], pre2[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int gy = y0 + row0 + k;
      pre[k] = pre2[k] = 0.f;
      if (gy < H && gx_own < W) {
        const unsigned o = (unsigned)gy * (unsigned)W + (unsigned)gx_own;
        if (multi) {
          pre[k] = (a.flags & PPEA_F_MOTION_MASK) ? __ldg(a.cons_mask + (size_t)b * plane + o) : 1.f;
          pre2[k] = __ldg(sc.mono_depth + (size_t)b * plane + o);
        } else if (automask && sc.noise) {
          pre[k] = __ldg(sc.noise + (size_t)b * plane + o);
        }
      }
    }

    photometric_pass<R, EW, PLANE, true>(&sm.x[0][0], &sm.y[0][0], row0, col, no_ssim, acc, cs);

    // ---- epilogue: min over sources, selec_reproj, mask, per-pixel outputs, block sums
    float s_rm = 0.f, s_m = 0.f, s_c = 0.f;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int gy = y0 + row0 + k;
      if (gy < H && gx_own < W) {
        const size_t o = (size_t)b * plane + (size_t)gy * W + gx_own;
        const Select sl = select_source(acc[k].x, acc[k].y, cs[k].x, cs[k].y, a.flags & PPEA_F_SELEC_REPROJ);
        unsigned bits = (unsigned)sl.src;
        float mask = 1.f;
        if (multi) {
          mask = pre[k] * one_minus_aug;
          bits |= PPEA_SEL_AUTOMASK;
          s_c += fabsf(sc.depth[o] - pre2[k]) * (1.f - mask);
        } else if (automask) {
          const float idl = sc.noise ? add_rn(ident[k], mul_rn(pre[k], 0.00001f)) : ident[k];   // trainer.py:1086-1087
          const bool on = sl.r <= idl;                                         // argmin([r, id]) == 0
          mask = on ? 1.f : 0.f;
          if (on) bits |= PPEA_SEL_AUTOMASK;
        } else {
          bits |= PPEA_SEL_AUTOMASK;
        }
        if (sc.loss_px) sc.loss_px[o] = sl.r;
        sc.sel[o] = (uint8_t)bits;
        s_rm += sl.r * mask;
        s_m += mask;
      }
    }
    s_rm = warp_sum(s_rm);
    s_m = warp_sum(s_m);
    s_c = warp_sum(s_c);
    if (lane == 0) {
      sm.red[0][wid] = s_rm;
      sm.red[1][wid] = s_m;
      sm.red[2][wid] = s_c;
    }
    __syncthreads();   // also fences the x planes before the next scale overwrites them
    if (tid < 3) {
      float t = 0.f;
      for (int w = 0; w < NT / 32; ++w) t += sm.red[tid][w];
      a.partials[((size_t)tile_id * a.S + s) * 4 + tid] = t;
    }
  }
}

cudaError_t launch_vsl_forward(const VslArgs& a, cudaStream_t stream) {
  using Smem = FwdSmem<kFwdTileW, kFwdTileH>;
  auto kern = vsl_forward_kernel<kFwdTileW, kFwdTileH, kFwdThreads>;
  static_assert(sizeof(Smem) <= 227 * 1024, "shared memory tile too large");
  cudaError_t e = ensure_dynamic_smem(kern, (int)sizeof(Smem));
  if (e != cudaSuccess) return e;
  const int nblk = a.B * a.tiles_x * a.tiles_y + a.S * a.B * kSmoothChunks;
  kern<<<nblk, kFwdThreads, sizeof(Smem), stream>>>(a);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// finish: fixed-order reduction of every scale's block partials, masked mean,
// consistency mean, smoothness and the multi-scale total (trainer.py:1113-1114,
// 1132, 1145-1158).  One CTA; warp w reduces the tile partials of scale w, then
// the per-image smoothness sums; thread 0 assembles the loss vector.
// ---------------------------------------------------------------------------
constexpr int kFinishWarps = 8;                          // warps per scale
constexpr int kFinishThreads = 32 * kFinishWarps * kMaxScales;

// Fused step: CTAs 1..B of the finish launch reduce the per-tile pose partials of one image each to per-scale sums
// (fixed order, double), so that the backward only has to weight S x 24 numbers per image (vsl_fused.cu pose_combine_role).
// Thread t < 768 owns entry t % 24 of the tiles t / 24, t / 24 + 32, ...: a warp reads 32 consecutive floats of the
// [tile][scale][24] array (coalesced), every load is independent; the 32 tile groups are then added in group order.
__device__ __forceinline__ void pose_sums_role(const VslArgs& a, int b, int tiles_per_image) {
  constexpr int kGroups = 32;
  __shared__ double red[kGroups][kMaxScales][24];
  const int tid = threadIdx.x, S = a.S;
  if (tid < kGroups * 24) {
    const int e = tid % 24, g = tid / 24;
    double t[kMaxScales] = {0, 0, 0, 0};
    const float* p = a.pose_partials + (size_t)b * tiles_per_image * S * 24 + e;
#pragma unroll 2
    for (int i = g; i < tiles_per_image; i += kGroups) {
#pragma unroll
      for (int s = 0; s < kMaxScales; ++s)
        if (s < S) t[s] += (double)__ldg(p + ((size_t)i * S + s) * 24);
    }
#pragma unroll
    for (int s = 0; s < kMaxScales; ++s) red[g][s][e] = t[s];
  }
  __syncthreads();
  if (tid < 24 * S) {
    const int e = tid % 24, s = tid / 24;
    double t = 0;
    for (int g = 0; g < kGroups; ++g) t += red[g][s][e];
    a.pose_sums[((size_t)b * S + s) * 24 + e] = t;
  }
}

__global__ void __launch_bounds__(kFinishThreads) vsl_finish_kernel(const __grid_constant__ VslArgs a, int nblk) {
  __shared__ double part[kMaxScales][kFinishWarps][3];
  __shared__ double tot[kMaxScales][4];     // sum r*mask, sum mask, sum cons, normalised smoothness
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int s = wid / kFinishWarps, w = wid % kFinishWarps;
  const int stride = sums_stride(a.B);
  grid_dependency_wait();        // the partial sums of the main launch
  grid_launch_dependents();      // few CTAs: the next kernel's CTAs may take the idle SMs now (they wait for our sums)
  if (blockIdx.x > 0) {
    pose_sums_role(a, blockIdx.x - 1, a.tiles_x * a.tiles_y);
    return;
  }
  if (s < a.S) {
    // tile partials of scale s: 256 threads, independent loads (4 in flight per thread)
    double v0 = 0, v1 = 0, v2 = 0;
    const int t = w * 32 + lane, nt = kFinishWarps * 32;
#pragma unroll 4
    for (int i = t; i < nblk; i += nt) {
      const float4 p = *reinterpret_cast<const float4*>(a.partials + ((size_t)i * a.S + s) * 4);
      v0 += (double)p.x;
      v1 += (double)p.y;
      v2 += (double)p.z;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, o);
      v1 += __shfl_xor_sync(0xffffffffu, v1, o);
      v2 += __shfl_xor_sync(0xffffffffu, v2, o);
    }
    if (lane == 0) {
      part[s][w][0] = v0;
      part[s][w][1] = v1;
      part[s][w][2] = v2;
    }
  }
  // per-image smoothness sums (kept for the backward): one warp per (scale, image), one lane per chunk
  static_assert(kSmoothChunks % 32 == 0, "a lane owns the chunks lane, lane + 32, ...");
  for (int pair = wid; pair < a.S * a.B; pair += kFinishThreads / 32) {
    const int ps = pair / a.B, b = pair - ps * a.B;
    const ScaleArgs& sc = a.sc[ps];
    const float* sws = a.smooth_ws + ((size_t)ps * a.B + b) * kSmoothChunks * 3;
    double ds = 0, sx = 0, sy = 0;
#pragma unroll
    for (int k = 0; k < kSmoothChunks / 32; ++k) {       // fixed order
      const float* q = sws + (k * 32 + lane) * 3;
      ds += (double)q[0], sx += (double)q[1], sy += (double)q[2];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ds += __shfl_xor_sync(0xffffffffu, ds, o);
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
    }
    if (lane == 0) {
      float* img = a.sums + (size_t)ps * stride + PPEA_SUMS_PER_SCALE + 4 * b;
      img[0] = (float)ds;
      img[1] = (float)sx;
      img[2] = (float)sy;
      const float n_sx = (float)a.B * sc.hs * (sc.ws - 1), n_sy = (float)a.B * (sc.hs - 1) * sc.ws;
      const float inv = 1.f / ((float)ds / (float)(sc.hs * sc.ws) + 1e-7f);
      img[3] = inv * ((float)sx / n_sx + (float)sy / n_sy);      // image b's share of the normalised smoothness
    }
  }
  __syncthreads();
  if (s < a.S && w == 0 && lane == 0) {
    float* row = a.sums + (size_t)s * stride;
    double v[3] = {0, 0, 0}, v3 = 0, rawx = 0, rawy = 0;
    for (int k = 0; k < kFinishWarps; ++k)
      for (int j = 0; j < 3; ++j) v[j] += part[s][k][j];
    for (int b = 0; b < a.B; ++b) {       // fixed order
      const float* img = row + PPEA_SUMS_PER_SCALE + 4 * b;
      rawx += (double)img[1];
      rawy += (double)img[2];
      v3 += (double)img[3];
    }
    tot[s][0] = v[0];
    tot[s][1] = v[1];
    tot[s][2] = v[2];
    tot[s][3] = v3;
    row[0] = (float)v[0];
    row[1] = (float)v[1];
    row[2] = (float)v[2];
    row[3] = (float)rawx;
    row[4] = (float)rawy;
    row[5] = row[6] = row[7] = 0.f;
  }
  __syncthreads();
  if (tid == 0) {
    float total = 0.f;
    const double n_px = (double)a.B * a.H * a.W;
    for (int k = 0; k < a.S; ++k) {
      const float reproj = (float)tot[k][0] / ((float)tot[k][1] + 1e-7f);
      const float cons = (a.flags & PPEA_F_MULTI) ? (float)(tot[k][2] / n_px) : 0.f;
      const float smooth = (float)tot[k][3];
      float loss = reproj + cons;
      loss += a.disparity_smoothness * smooth / (float)(1 << (a.first_scale + k));
      float* L = a.losses + 1 + k * PPEA_LOSSES_PER_SCALE;
      L[0] = loss;
      L[1] = reproj;
      L[2] = cons;
      L[3] = smooth;
      total += loss;
    }
    a.losses[0] = total / (float)a.total_scales;
    if (a.fmt_flag) *a.fmt_flag = 0u;      // streaming step: the main launch has consumed the format flag; clear it for the next step
  }
}

cudaError_t launch_vsl_finish(const VslArgs& a, int nblk_fwd, cudaStream_t stream, bool pose_sums) {
  return launch_pdl(vsl_finish_kernel, dim3(1 + (pose_sums ? a.B : 0)), dim3(kFinishThreads), 0, stream, a, nblk_fwd);
}

}  // namespace ppea
