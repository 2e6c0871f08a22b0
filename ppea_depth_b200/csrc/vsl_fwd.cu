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

namespace ppea {

template <int TW, int TH>
struct FwdSmem {
  static constexpr int EW = TW + 2;
  static constexpr int EH = TH + 2;
  static constexpr int PLANE = EW * EH;
  f2 y[3][PLANE];      // target, duplicated lanes
  f2 x[3][PLANE];      // (source 0, source 1): un-warped, then warped
  f2 G[12];            // per-source geometry M[0..8], t[0..2] (vsl_math.cuh Geom), lanes = sources
  float red[3][8];
};

// Photometric loss 0.85*mean_c SSIM + 0.15*mean_c |y-x| (trainer.py:995-1007) of both sources
// against the target for the R pixels (rows row0..row0+R-1, column col) of this thread.
template <int R, int EW, int PLANE, bool WITH_CSUM>
__device__ __forceinline__ void photometric_pass(const f2* __restrict__ xs, const f2* __restrict__ ys, int row0, int col,
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
    const f2* yp = ys + c * PLANE + row0 * EW + col;
    const f2* xp = xs + c * PLANE + row0 * EW + col;
    f2 hy[3], hyy[3], hx[3], hxx[3], hxy[3];
#pragma unroll
    for (int i = 0; i < R + 2; ++i) {
      const int s = i % 3;
      const f2 y0 = yp[i * EW], y1 = yp[i * EW + 1], y2 = yp[i * EW + 2];
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

template <int TW, int TH, int NT>
__global__ void __launch_bounds__(NT, 5) vsl_forward_kernel(const __grid_constant__ VslArgs a) {
  using Smem = FwdSmem<TW, TH>;
  constexpr int EW = Smem::EW, PLANE = Smem::PLANE;
  constexpr int R = (TW * TH) / NT;
  static_assert(NT % TW == 0 && (TW * TH) % NT == 0, "tile/thread mismatch");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

  const int tid = threadIdx.x;
  int blk = blockIdx.x;
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
    for (int c = 0; c < 3; ++c) sm.y[c][idx] = dup2(__ldg(tgt_b + c * plane + o));
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

#pragma unroll 1
  for (int s = 0; s < a.S; ++s) {
    const ScaleArgs& sc = a.sc[s];
    // ---- gather pass: depth, projection into both sources, bilinear samples -> x planes
    {
      const float* disp_b = sc.disp + (size_t)b * sc.hs * sc.ws;
      const bool same_res = (sc.hs == H && sc.ws == W);
      for (int idx = tid; idx < PLANE; idx += NT) {
        const int i = idx / EW, j = idx - i * EW;
        const int gy = y0 - 1 + i, gx = x0 - 1 + j;
        const int py = reflect_index(gy, H), px = reflect_index(gx, W);
        float dup;
        if (same_res) {
          dup = __ldg(disp_b + (size_t)py * W + px);
        } else {
          const UpCoef cy = up_coef(py, sc.hs, sc.up_sy), cx = up_coef(px, sc.ws, sc.up_sx);
          dup = up_sample(disp_b, sc.ws, cy, cx);
        }
        const float dep = depth_from_disp(dup, a.disp_lo, a.disp_range);
        if (i >= 1 && i <= TH && j >= 1 && j <= TW && gy < H && gx < W) sc.depth[(size_t)b * plane + (size_t)gy * W + gx] = dep;
        const f2 fx = dup2(int_to_float(px)), fy = dup2(int_to_float(py));
        const f2 A0 = vfma(sm.G[1], fy, vfma(sm.G[0], fx, sm.G[2]));
        const f2 A1 = vfma(sm.G[4], fy, vfma(sm.G[3], fx, sm.G[5]));
        const f2 A2 = vfma(sm.G[7], fy, vfma(sm.G[6], fx, sm.G[8]));
        const ProjT<f2> pr = project_fast(dep, A0, A1, A2, sm.G[9], sm.G[10], sm.G[11], a.eps, wmax, hmax);
        const Bilin b0 = bilin_setup(pr.ix.x, pr.iy.x, W), b1 = bilin_setup(pr.ix.y, pr.iy.y, W);
        const f2 wnw = mk2(b0.wnw, b1.wnw), wne = mk2(b0.wne, b1.wne), wsw = mk2(b0.wsw, b1.wsw), wse = mk2(b0.wse, b1.wse);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float* S0 = src_b[0] + c * plane + b0.o00;
          const float* S1 = src_b[1] + c * plane + b1.o00;
          const f2 nw = mk2(__ldg(S0), __ldg(S1)), ne = mk2(__ldg(S0 + 1), __ldg(S1 + 1));
          const f2 sw = mk2(__ldg(S0 + W), __ldg(S1 + W)), se = mk2(__ldg(S0 + W + 1), __ldg(S1 + W + 1));
          sm.x[c][idx] = vfma(se, wse, vfma(sw, wsw, vfma(ne, wne, vmul(nw, wnw))));
        }
      }
    }
    __syncthreads();

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
          if (a.flags & PPEA_F_MOTION_MASK) mask = a.cons_mask[o];
          mask *= one_minus_aug;
          bits |= PPEA_SEL_AUTOMASK;
          s_c += fabsf(sc.depth[o] - sc.mono_depth[o]) * (1.f - mask);
        } else if (automask) {
          const float idl = sc.noise ? add_rn(ident[k], mul_rn(sc.noise[o], 0.00001f)) : ident[k];   // trainer.py:1086-1087
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
      a.partials[((size_t)blockIdx.x * a.S + s) * 4 + tid] = t;
    }
  }
}

cudaError_t launch_vsl_forward(const VslArgs& a, cudaStream_t stream) {
  using Smem = FwdSmem<kFwdTileW, kFwdTileH>;
  auto kern = vsl_forward_kernel<kFwdTileW, kFwdTileH, kFwdThreads>;
  static_assert(sizeof(Smem) <= 227 * 1024, "shared memory tile too large");
  cudaError_t e = ensure_dynamic_smem(kern, (int)sizeof(Smem));
  if (e != cudaSuccess) return e;
  const int nblk = a.B * a.tiles_x * a.tiles_y;
  kern<<<nblk, kFwdThreads, sizeof(Smem), stream>>>(a);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// finish: fixed-order reduction of every scale's block partials, masked mean,
// consistency mean, smoothness and the multi-scale total (trainer.py:1113-1114,
// 1132, 1145-1158).  A single CTA; the partial lists are a few thousand floats.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vsl_finish_kernel(const __grid_constant__ VslArgs a, int nblk) {
  __shared__ double red[5][8];
  __shared__ float scale_loss[kMaxScales];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int stride = PPEA_SUMS_PER_SCALE + 4 * a.B;
  for (int s = 0; s < a.S; ++s) {
    const ScaleArgs& sc = a.sc[s];
    float* row = a.sums + (size_t)s * stride;
    const float* sws = a.smooth_ws + (size_t)s * a.B * kSmoothChunks * 3;
    double v[5] = {0, 0, 0, 0, 0};
    // per-image smoothness statistics (needed again by the backward), fixed order
    for (int b = tid; b < a.B; b += 256) {
      double ds = 0, sx = 0, sy = 0;
      for (int c = 0; c < kSmoothChunks; ++c) {
        ds += (double)sws[(b * kSmoothChunks + c) * 3 + 0];
        sx += (double)sws[(b * kSmoothChunks + c) * 3 + 1];
        sy += (double)sws[(b * kSmoothChunks + c) * 3 + 2];
      }
      float* img = row + PPEA_SUMS_PER_SCALE + 4 * b;
      img[0] = (float)ds;
      img[1] = (float)sx;
      img[2] = (float)sy;
      img[3] = 0.f;
      v[3] += sx;
      v[4] += sy;
    }
    for (int i = tid; i < nblk; i += 256) {
      const float* p = a.partials + ((size_t)i * a.S + s) * 4;
      v[0] += (double)p[0];
      v[1] += (double)p[1];
      v[2] += (double)p[2];
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
      if (lane == 0) red[k][wid] = v[k];
    }
    __syncthreads();
    if (tid == 0) {
      double t[5];
      for (int k = 0; k < 5; ++k) {
        t[k] = 0;
        for (int w = 0; w < 8; ++w) t[k] += red[k][w];
      }
      for (int k = 0; k < 5; ++k) row[k] = (float)t[k];
      row[5] = row[6] = row[7] = 0.f;
      const double n_px = (double)a.B * a.H * a.W;
      const double n_sx = (double)a.B * sc.hs * (sc.ws - 1), n_sy = (double)a.B * (sc.hs - 1) * sc.ws;
      const float reproj = (float)t[0] / ((float)t[1] + 1e-7f);
      const float cons = (a.flags & PPEA_F_MULTI) ? (float)(t[2] / n_px) : 0.f;
      const float smooth = (float)(t[3] / n_sx) + (float)(t[4] / n_sy);
      float loss = reproj + cons;
      loss += a.disparity_smoothness * smooth / (float)(1 << (a.first_scale + s));
      float* L = a.losses + 1 + s * PPEA_LOSSES_PER_SCALE;
      L[0] = loss;
      L[1] = reproj;
      L[2] = cons;
      L[3] = smooth;
      scale_loss[s] = loss;
    }
    __syncthreads();
  }
  if (tid == 0) {
    float total = 0.f;
    for (int k = 0; k < a.S; ++k) total += scale_loss[k];
    a.losses[0] = total / (float)a.total_scales;
  }
}

cudaError_t launch_vsl_finish(const VslArgs& a, int nblk_fwd, cudaStream_t stream) {
  vsl_finish_kernel<<<1, 256, 0, stream>>>(a, nblk_fwd);
  return cudaGetLastError();
}

}  // namespace ppea
