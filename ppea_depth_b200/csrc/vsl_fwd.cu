// vsl_fwd.cu -- fused forward of one pyramid scale of the view-synthesis loss.
//
// One launch replaces, for one scale, trainer.py:886-914 (upsample, depth,
// backproject, project, grid_sample for both sources) and trainer.py:1050-1141
// (SSIM+L1 photometric loss of the warped and of the un-warped sources, min over
// sources, selec_reproj, identity automask / multi-frame mask, masked sums).
//
// A CTA owns a TW x TH tile of one image.  Shared memory holds the tile plus a
// one-pixel reflection halo of: the target (3 planes) and both sources' colour
// (2 x 3 planes) -- first the un-warped sources (identity loss), then, in place,
// the warped ones.  Every 3x3 SSIM window is then assembled from shared memory
// by a thread that walks R consecutive rows of one column with a three-row
// sliding window of horizontal sums held in registers.
// HBM traffic per pixel: tgt 12 B + sources 24 B (+ gathers, L1/L2-resident) +
// noise 4 B + disp 4/4^s B in; depth 4 B + loss 4 B + sel 1 B out.
#include "vsl_common.cuh"

namespace ppea {

template <int TW, int TH>
struct FwdSmem {
  static constexpr int EW = TW + 2;
  static constexpr int EH = TH + 2;
  static constexpr int PLANE = EW * EH;
  float y[3][PLANE];
  float x[2][3][PLANE];
  float P[2][12];
  float iK[9];
  float red[3][32];
};

// Photometric loss 0.85*mean_c SSIM + 0.15*mean_c |y-x| (trainer.py:995-1007) of both
// x-planes against y for the R pixels (rows row0..row0+R-1, column col) of this thread.
template <int R, int EW, int PLANE, bool WITH_CSUM>
__device__ __forceinline__ void photometric_pass(const float* __restrict__ xs, const float* __restrict__ ys, int row0,
                                                 int col, bool no_ssim, float (&acc)[2][R], float (&cs)[2][R]) {
#pragma unroll
  for (int f = 0; f < 2; ++f)
#pragma unroll
    for (int k = 0; k < R; ++k) {
      acc[f][k] = 0.f;
      cs[f][k] = 0.f;
    }
  const float w_l1 = no_ssim ? (1.f / 3.f) : PPEA_W_L1;
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
    const float* yp = ys + c * PLANE + row0 * EW + col;
    const float* xp0 = xs + c * PLANE + row0 * EW + col;
    const float* xp1 = xs + (3 + c) * PLANE + row0 * EW + col;
    float hy[3], hyy[3], hx[2][3], hxx[2][3], hxy[2][3];
#pragma unroll
    for (int i = 0; i < R + 2; ++i) {
      const int s = i % 3;
      const float y0 = yp[i * EW], y1 = yp[i * EW + 1], y2 = yp[i * EW + 2];
      hy[s] = y0 + y1 + y2;
      hyy[s] = y0 * y0 + y1 * y1 + y2 * y2;
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        const float* xp = f ? xp1 : xp0;
        const float x0 = xp[i * EW], x1 = xp[i * EW + 1], x2 = xp[i * EW + 2];
        hx[f][s] = x0 + x1 + x2;
        hxx[f][s] = x0 * x0 + x1 * x1 + x2 * x2;
        hxy[f][s] = x0 * y0 + x1 * y1 + x2 * y2;
        if (i >= 1 && i <= R) {
          acc[f][i - 1] += w_l1 * fabsf(y1 - x1);
          if (WITH_CSUM) cs[f][i - 1] += x1;
        }
      }
      if (i >= 2 && !no_ssim) {
        const SsimY yst = ssim_y_stats(hy[0] + hy[1] + hy[2], hyy[0] + hyy[1] + hyy[2]);
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          const float S = ssim_from_sums(hx[f][0] + hx[f][1] + hx[f][2], hxx[f][0] + hxx[f][1] + hxx[f][2],
                                         hxy[f][0] + hxy[f][1] + hxy[f][2], yst);
          acc[f][i - 2] += PPEA_W_SSIM * S;
        }
      }
    }
  }
}

template <int TW, int TH, int NT>
__global__ void __launch_bounds__(NT) vsl_forward_kernel(const __grid_constant__ VslArgs a) {
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
    // P_f = (K @ T_f)[:3,:]: one thread per entry
    const int f = tid / 12, e = tid % 12, i = e / 4, j = e % 4;
    const float* K = a.K + b * 16;
    const float* T = a.T[f] + b * 16;
    float acc = mul_rn(K[i * 4 + 0], T[0 * 4 + j]);
    acc = add_rn(acc, mul_rn(K[i * 4 + 1], T[1 * 4 + j]));
    acc = add_rn(acc, mul_rn(K[i * 4 + 2], T[2 * 4 + j]));
    acc = add_rn(acc, mul_rn(K[i * 4 + 3], T[3 * 4 + j]));
    sm.P[f][e] = acc;
  } else if (tid >= 32 && tid < 41) {
    const int e = tid - 32;
    sm.iK[e] = a.inv_K[b * 16 + (e / 3) * 4 + (e % 3)];
  }

  const float* tgt_b = a.tgt + (size_t)b * 3 * plane;
  const float* src_b[2] = {a.src[0] + (size_t)b * 3 * plane, a.src[1] + (size_t)b * 3 * plane};

  // ---- pass 1: stage target (+ un-warped sources for the identity loss) with reflection halo
  for (int idx = tid; idx < PLANE; idx += NT) {
    const int i = idx / EW, j = idx - i * EW;
    const int py = reflect_index(y0 - 1 + i, H), px = reflect_index(x0 - 1 + j, W);
    const size_t o = (size_t)py * W + px;
#pragma unroll
    for (int c = 0; c < 3; ++c) sm.y[c][idx] = __ldg(tgt_b + c * plane + o);
    if (automask) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        sm.x[0][c][idx] = __ldg(src_b[0] + c * plane + o);
        sm.x[1][c][idx] = __ldg(src_b[1] + c * plane + o);
      }
    }
  }
  __syncthreads();

  const int col = tid % TW;
  const int row0 = (tid / TW) * R;
  float ident[R];
  float acc[2][R], cs[2][R];
  if (automask) {
    photometric_pass<R, EW, PLANE, false>(&sm.x[0][0][0], &sm.y[0][0], row0, col, no_ssim, acc, cs);
#pragma unroll
    for (int k = 0; k < R; ++k) ident[k] = fminf(acc[0][k], acc[1][k]);   // trainer.py:1069
    __syncthreads();
  }

  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  float iK[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) iK[e] = sm.iK[e];
  const int gx_own = x0 + col;
  const float one_minus_aug = (multi && (a.flags & PPEA_F_MATCH_AUG)) ? 1.f - a.aug_mask[b] : 1.f;
  const int lane = tid & 31, wid = tid >> 5;

#pragma unroll 1
  for (int s = 0; s < a.S; ++s) {
    const ScaleArgs& sc = a.sc[s];
    // ---- pass 2: depth, backproject, project, bilinear gather of both sources into the x planes
    {
      const float* disp_b = sc.disp + (size_t)b * sc.hs * sc.ws;
      for (int idx = tid; idx < PLANE; idx += NT) {
        const int i = idx / EW, j = idx - i * EW;
        const int gy = y0 - 1 + i, gx = x0 - 1 + j;
        const int py = reflect_index(gy, H), px = reflect_index(gx, W);
        const UpCoef cy = up_coef(py, sc.hs, sc.up_sy), cx = up_coef(px, sc.ws, sc.up_sx);
        const float dep = depth_from_disp(up_sample(disp_b, sc.ws, cy, cx), a.disp_lo, a.disp_range);
        if (i >= 1 && i <= TH && j >= 1 && j <= TW && gy < H && gx < W) sc.depth[(size_t)b * plane + (size_t)gy * W + gx] = dep;
        float ray[3], cam[3];
        pixel_ray(iK, (float)px, (float)py, ray);
#pragma unroll
        for (int e = 0; e < 3; ++e) cam[e] = mul_rn(dep, ray[e]);
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          const Proj pr = project_point(sm.P[f], cam, a.eps, wm1, hm1);
          const Bilin bl = bilin_setup(pr.ix, pr.iy, W, H);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float* S = src_b[f] + c * plane;
            sm.x[f][c][idx] = bilin_value(bl, __ldg(S + bl.o00), __ldg(S + bl.o01), __ldg(S + bl.o10), __ldg(S + bl.o11));
          }
        }
      }
    }
    __syncthreads();

    photometric_pass<R, EW, PLANE, true>(&sm.x[0][0][0], &sm.y[0][0], row0, col, no_ssim, acc, cs);

    // ---- epilogue: min over sources, selec_reproj, mask, per-pixel outputs, block sums
    float s_rm = 0.f, s_m = 0.f, s_c = 0.f;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int gy = y0 + row0 + k;
      if (gy < H && gx_own < W) {
        const size_t o = (size_t)b * plane + (size_t)gy * W + gx_own;
        const Select sl = select_source(acc[0][k], acc[1][k], cs[0][k], cs[1][k], a.flags & PPEA_F_SELEC_REPROJ);
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
