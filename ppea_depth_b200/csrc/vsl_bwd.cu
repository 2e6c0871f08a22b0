// vsl_bwd.cu -- fused backward of the view-synthesis loss, all pyramid scales in one launch.
//
// Autograd of trainer.py:886-914 + 1050-1141 wrt every disp_s and (mono path) the two poses,
// recomputing the forward instead of storing it: the only forward products read back are the
// selection map `sel` (1 B/px) and the reduced sums.  HBM traffic per pixel and scale: tgt 12 B
// (once per CTA, all scales) + source gathers (24 B compulsory) + sel 1 B (+ cons_mask /
// mono_depth on the multi path) in; grad_disp_s 4/4^s B out.
//
// A CTA owns a TW x TH tile of one image; all photometric arithmetic runs 2-wide
// (FFMA2/FADD2/FMUL2, lanes = the two sources).
//   phase 1  every cell of the tile + 2-pixel halo: depth -> projection -> bilinear gather of
//            both sources into shared memory (halo cells hold the value of the reflected pixel,
//            layers.py:238).  For its own R tile pixels a thread keeps d warped / d (u, v)
//            (clip-masked, GridSampler.h clip_coordinates_set_grad) in registers.
//   per channel c:
//   phase 2  every pixel q of the tile + 1-pixel halo: SSIM window sums -> adjoint coefficients
//            (cA,cB,cC), weighted per lane by  g_r * mask(q) * [sel(q) == source]  (zero for the
//            source whose loss was not taken at q).
//   phase 3  d L / d warped_c(p) = A + B x(p) + C y(p) + L1 term, where A,B,C are 3x3 box sums of
//            the coefficients with the reflection multiplicities (a border window counts its
//            mirrored tap twice), computed with a sliding three-row window in registers; folded
//            into d L / d (u, v) with the kept derivatives.
//   phase 4  chain through Project3D (layers.py:185-194) and BackprojectDepth (layers.py:164-166)
//            to d L / d depth and the pose partials; consistency term (multi path);
//            d depth / d disp (layers.py:21-22); adjoint of the bilinear upsample
//            (trainer.py:886-887): direct accumulate at scale 0, float atomics or
//            (deterministic) a full-res scratch + gather pass for coarser scales.
#include "vsl_common.cuh"
#include <type_traits>

#include "smooth.cuh"
#include "vsl_gather.cuh"

namespace ppea {

template <int TW, int TH>
struct BwdSmem {
  static constexpr int RW = TW + 4, RH = TH + 4, RP = RW * RH;   // value region (2-pixel halo)
  static constexpr int QW = TW + 2, QH = TH + 2, QP = QW * QH;   // coefficient region (1-pixel halo)
  float y[3][RP];       // target (broadcast into both lanes at use)
  f2 x[3][RP];          // warped (source 0, source 1)
  f2 cf[3][QP];         // cA, cB, cC of the current channel
  f2 wq[QP];            // g_r * mask(q) * [sel(q) == lane]
  f2 G[12];             // per-source geometry (vsl_math.cuh Geom), lanes = sources
  float red[24][8];
};

#ifndef PPEA_BWD_CTAS
#define PPEA_BWD_CTAS (kBwdThreads >= 256 ? 2 : 4)
#endif
template <int TW, int TH, int NT, bool POSE>
__global__ void __launch_bounds__(NT, PPEA_BWD_CTAS) vsl_backward_kernel(const __grid_constant__ VslArgs a) {
  using Smem = BwdSmem<TW, TH>;
  constexpr int RW = Smem::RW, RP = Smem::RP, QW = Smem::QW, QP = Smem::QP;
  constexpr int R = (TW * TH) / NT;
  static_assert(TW == 32 && (NT / 32) * R == TH && NT >= 128, "the gather/row mapping assumes lane == tile column and R rows per warp");
  static_assert(NT % TW == 0 && (TW * TH) % NT == 0, "tile/thread mismatch");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

  // non-deterministic mode: the first CTAs of the grid add the smoothness gradient of every scale (smooth.cuh)
  const int n_smooth = (a.flags & PPEA_F_DETERMINISTIC) ? 0 : a.S * a.B * kSmoothChunks;
  if ((int)blockIdx.x < n_smooth) {
    smooth_backward_role<1>(a, blockIdx.x);
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
  const bool no_ssim = a.flags & PPEA_F_NO_SSIM;
  const float l1w = no_ssim ? (1.f / 3.f) : PPEA_W_L1;

  if (tid < 24) {
    const int f = tid / 12, e = tid % 12;
    const float v = geom_entry(a.K + b * 16, a.T[f] + b * 16, a.inv_K + b * 16, e);
    (f ? sm.G[e].y : sm.G[e].x) = v;
  }

  const float* tgt_b = a.tgt + (size_t)b * 3 * plane;
  const float* src_b[2] = {a.src[0] + (size_t)b * 3 * plane, a.src[1] + (size_t)b * 3 * plane};

  // ---- stage the target with a 2-pixel reflection halo (shared by every scale)
  for (int idx = tid; idx < RP; idx += NT) {
    const int i = idx / RW, j = idx - i * RW;
    const int py = reflect_index(y0 - 2 + i, H), px = reflect_index(x0 - 2 + j, W);
    const size_t o = (size_t)py * W + px;
#pragma unroll
    for (int c = 0; c < 3; ++c) sm.y[c][idx] = __ldg(tgt_b + c * plane + o);
  }
  __syncthreads();

  const float wmax = coord_max(W), hmax = coord_max(H);
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  const int col = tid % TW;
  const int row0 = (tid / TW) * R;
  const int gx_own = x0 + col;
  const ColCtx col_own = make_col(sm.G, gx_own, W);              // this thread's tile column (reflect-clamped for partial tiles)
  const int px_own = col_own.px;
  const SrcPlanes sp = make_planes(src_b[0], src_b[1], plane);
  const float one_minus_aug = (multi && (a.flags & PPEA_F_MATCH_AUG)) ? 1.f - a.aug_mask[b] : 1.f;
  // reflection multiplicities of the taps left/right of this thread's column
  const f2 mL = dup2((gx_own == 1) ? 2.f : 1.f), mR = dup2((gx_own == W - 2) ? 2.f : 1.f);


#pragma unroll 1
  for (int s = 0; s < a.S; ++s) {
    const ScaleArgs& sc = a.sc[s];
    const float* disp_b = sc.disp + (size_t)b * sc.hs * sc.ws;
    const bool same_res = (sc.hs == H && sc.ws == W);
    const float* srow = a.sums + (size_t)s * sums_stride(a.B);
    const ScaleGrads sg = scale_grads(a, s);
    const float g_r = sg.reproj / (srow[1] + 1e-7f);           // d reproj_s / d (r*mask)(q)   trainer.py:1114
    const float g_c = sg.cons / ((float)a.B * (float)plane);   // d cons_s / d (|depth-mono|*(1-mask))(q)

    float dep[R];
    f2 ddx[R][3], ddy[R][3];     // d warped_c / d u, d warped_c / d v  (lanes = sources)
    f2 gu[R], gv[R];

    // ---- phase 1: gather.  Warp w walks its own R tile rows (lane == tile column, derivatives kept) plus
    // its share of the four halo rows; the four halo columns are a flat list of extra cells.
    ColCtx cc = col_own;
    if (!same_res) cc.cx = up_coef(cc.px, sc.ws, sc.up_sx);
    auto gather_row = [&](int i, auto want_deriv, f2 (&dx)[3], f2 (&dy)[3]) -> float {
      const int py = reflect_index(y0 - 2 + i, H);
      UpCoef cy;
      float d;
      if (same_res) {
        d = depth_of<true>(disp_b, W, sc.ws, py, cc, cy, a.disp_lo, a.disp_range);
      } else {
        cy = up_coef(py, sc.hs, sc.up_sy);
        d = depth_of<false>(disp_b, W, sc.ws, py, cc, cy, a.disp_lo, a.disp_range);
      }
      f2 A[3], val[3];
      const ProjT<f2> pr = project_cell(sm.G, cc, py, d, a.eps, wmax, hmax, A);
      sample_sources<decltype(want_deriv)::value, false>(sp, W, pr, wm1, hm1, val, dx, dy);   // (shuffle-sharing costs registers this kernel does not have)
      const int ridx = i * RW + col + 2;
#pragma unroll
      for (int c = 0; c < 3; ++c) sm.x[c][ridx] = val[c];
      return d;
    };
#pragma unroll
    for (int k = 0; k < R; ++k) {
      dep[k] = gather_row(row0 + k + 2, std::true_type{}, ddx[k], ddy[k]);
      gu[k] = gv[k] = dup2(0.f);
    }
    {
      f2 u0[3], u1[3];
      const int wid = tid >> 5;
      // halo rows 0,1 go to the first warp and RH-2,RH-1 to the last (adjacent to their own rows: L1 reuse);
      // the four halo columns of every region row are a flat list taken by the middle warps
      constexpr int NWB = NT / 32;
      if (wid == 0 || wid == NWB - 1) {
        const int base = (wid == 0) ? 0 : Smem::RH - 2;
        gather_row(base, std::false_type{}, u0, u1);
        gather_row(base + 1, std::false_type{}, u0, u1);
      }
      for (int e = tid - 32; e < 4 * Smem::RH && tid >= 32 && tid < NT - 32; e += NT - 64) {
        const int i = e >> 2, jj = e & 3, j = jj < 2 ? jj : RW - 4 + jj;
        const int py = reflect_index(y0 - 2 + i, H);
        ColCtx ce = make_col(sm.G, x0 - 2 + j, W);
        UpCoef cy;
        float d;
        if (same_res) {
          d = depth_of<true>(disp_b, W, sc.ws, py, ce, cy, a.disp_lo, a.disp_range);
        } else {
          ce.cx = up_coef(ce.px, sc.ws, sc.up_sx);
          cy = up_coef(py, sc.hs, sc.up_sy);
          d = depth_of<false>(disp_b, W, sc.ws, py, ce, cy, a.disp_lo, a.disp_range);
        }
        f2 A[3], val[3];
        const ProjT<f2> pr = project_cell(sm.G, ce, py, d, a.eps, wmax, hmax, A);
        sample_sources<false, false>(sp, W, pr, wm1, hm1, val, u0, u1);
        const int ridx = i * RW + j;
#pragma unroll
        for (int c = 0; c < 3; ++c) sm.x[c][ridx] = val[c];
      }
    }
    // ---- per-q weights  g_r * mask(q) * [sel(q) == lane]   (zero outside the image)
    for (int qi = tid; qi < QP; qi += NT) {
      const int i = qi / QW, j = qi - i * QW;
      const int qy = y0 - 1 + i, qx = x0 - 1 + j;
      f2 w = dup2(0.f);
      if (qy >= 0 && qy < H && qx >= 0 && qx < W) {
        const size_t o = (size_t)b * plane + (size_t)qy * W + qx;
        const unsigned bits = sc.sel[o];
        float mask;
        if (multi) {
          mask = (a.flags & PPEA_F_MOTION_MASK) ? a.cons_mask[o] : 1.f;
          mask *= one_minus_aug;
        } else {
          mask = (bits & PPEA_SEL_AUTOMASK) ? 1.f : 0.f;
        }
        const float wv = g_r * mask;
        const unsigned src = bits & PPEA_SEL_SRC_MASK;
        w = mk2(src == 0u ? wv : 0.f, src == 1u ? wv : 0.f);
      }
      sm.wq[qi] = w;
    }
    __syncthreads();

#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // ---- phase 2: adjoint coefficients of the SSIM windows centred in the tile + 1 halo
      if (!no_ssim) {
        for (int qi = tid; qi < QP; qi += NT) {
          const int i = qi / QW, j = qi - i * QW;
          const f2 w = sm.wq[qi];
          SsimAdjT<f2> ad;
          ad.cA = ad.cB = ad.cC = dup2(0.f);
          if (w.x != 0.f || w.y != 0.f) {
            const f2* xp = &sm.x[c][i * RW + j];
            const float* yp = &sm.y[c][i * RW + j];
            f2 hx[3], hxx[3], hxy[3], hy[3], hyy[3];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const f2 xa = xp[dy * RW], xb = xp[dy * RW + 1], xc = xp[dy * RW + 2];
              const f2 ya = dup2(yp[dy * RW]), yb = dup2(yp[dy * RW + 1]), yc = dup2(yp[dy * RW + 2]);
              row_sums_y<f2>(ya, yb, yc, hy[dy], hyy[dy]);
              row_sums_x<f2>(xa, xb, xc, ya, yb, yc, hx[dy], hxx[dy], hxy[dy]);
            }
            const SsimYT<f2> yst = ssim_y_stats<f2>(sum3(hy[0], hy[1], hy[2]), sum3(hyy[0], hyy[1], hyy[2]));
            ad = ssim_adjoint<f2>(sum3(hx[0], hx[1], hx[2]), sum3(hxx[0], hxx[1], hxx[2]), sum3(hxy[0], hxy[1], hxy[2]), yst,
                                  vmul(w, dup2(PPEA_W_SSIM)));
          }
          sm.cf[0][qi] = ad.cA;
          sm.cf[1][qi] = ad.cB;
          sm.cf[2][qi] = ad.cC;
        }
      }
      __syncthreads();

      // ---- phase 3: box-sum the coefficients (sliding window down this thread's column), fold into d L / d (u, v)
      {
        f2 h[3][3];
#pragma unroll
        for (int i = 0; i < R + 2; ++i) {
          const int sl = i % 3;
          if (!no_ssim) {
            const int qrow = (row0 + i) * QW + col;
#pragma unroll
            for (int e = 0; e < 3; ++e) {
              const f2* cp = &sm.cf[e][qrow];
              h[sl][e] = vfma(mL, cp[0], vfma(mR, cp[2], cp[1]));
            }
          }
          if (i >= 2) {
            const int k = i - 2;
            const int gy = y0 + row0 + k;
            const int ridx = (row0 + k + 2) * RW + col + 2;
            const f2 xv = sm.x[c][ridx], yv = dup2(sm.y[c][ridx]);
            const f2 wl = sm.wq[(row0 + k + 1) * QW + col + 1];
            const f2 d = vsub(yv, xv);
            // L1 term:  -w * l1w * sign(y - x)
            f2 G = mk2(-wl.x * l1w * sign_of(d.x), -wl.y * l1w * sign_of(d.y));
            if (!no_ssim) {
              const f2 mU = dup2((gy == 1) ? 2.f : 1.f), mD = dup2((gy == H - 2) ? 2.f : 1.f);
              const int su = (i - 2) % 3, smid = (i - 1) % 3;
              const f2 A = vfma(mU, h[su][0], vfma(mD, h[sl][0], h[smid][0]));
              const f2 Bc = vfma(mU, h[su][1], vfma(mD, h[sl][1], h[smid][1]));
              const f2 Cc = vfma(mU, h[su][2], vfma(mD, h[sl][2], h[smid][2]));
              G = vadd(G, vfma(Bc, xv, vfma(Cc, yv, A)));               // d L / d warped_c(p), both sources
            }
            gu[k] = vfma(G, ddx[k][c], gu[k]);
            gv[k] = vfma(G, ddy[k][c], gv[k]);
          }
        }
      }
      __syncthreads();   // coefficient planes are rewritten by the next channel / x planes by the next scale
    }

    // ---- phase 4: projection adjoint, consistency term, depth -> disp, adjoint of the bilinear upsample
    float* gd_b = sc.grad_disp + (size_t)b * sc.hs * sc.ws;
    // pose partials of this scale: Sw[r] = sum gc_r*depth, Swy[r] = sum gc_r*depth*y, Sg[r] = sum gc_r (x is this thread's
    // constant column).  Reduced and written per scale so that they are not live across the other phases.
    f2 Sw[POSE ? 3 : 1], Swy[POSE ? 3 : 1], Sg[POSE ? 3 : 1];
#pragma unroll
    for (int e = 0; e < (POSE ? 3 : 1); ++e) Sw[e] = Swy[e] = Sg[e] = dup2(0.f);
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int gy = y0 + row0 + k;
      if (gy < H && gx_own < W) {
        const size_t o = (size_t)b * plane + (size_t)gy * W + gx_own;
        f2 A[3];
        const ProjT<f2> pr = project_cell(sm.G, col_own, gy, dep[k], a.eps, wmax, hmax, A);
        const f2 gc0 = vmul(gu[k], pr.rz), gc1 = vmul(gv[k], pr.rz);
        const f2 gc2 = vneg(vmul(vfma(gu[k], pr.u, vmul(gv[k], pr.v)), pr.rz));
        const f2 gd2 = vfma(gc2, A[2], vfma(gc1, A[1], vmul(gc0, A[0])));
        float g = gd2.x + gd2.y;
        if (POSE) {
          const f2 dk = dup2(dep[k]), fy = dup2(int_to_float(gy));
          const f2 w0 = vmul(gc0, dk), w1 = vmul(gc1, dk), w2 = vmul(gc2, dk);
          Sw[0] = vadd(Sw[0], w0);
          Sw[1] = vadd(Sw[1], w1);
          Sw[2] = vadd(Sw[2], w2);
          Swy[0] = vfma(w0, fy, Swy[0]);
          Swy[1] = vfma(w1, fy, Swy[1]);
          Swy[2] = vfma(w2, fy, Swy[2]);
          Sg[0] = vadd(Sg[0], gc0);
          Sg[1] = vadd(Sg[1], gc1);
          Sg[2] = vadd(Sg[2], gc2);
        }
        if (multi) {
          float mask = (a.flags & PPEA_F_MOTION_MASK) ? a.cons_mask[o] : 1.f;
          mask *= one_minus_aug;
          g = fmaf(g_c * sign_of(dep[k] - sc.mono_depth[o]), 1.f - mask, g);
        }
        const float g_dup = g * ddepth_ddisp(dep[k], a.disp_range);
        if (same_res) {
          // deterministic mode: sole owner of this element, after smooth_backward's overwrite; otherwise the
          // smoothness CTAs of this very launch add into the same zero-initialised element (two commutative adds)
          if (a.flags & PPEA_F_DETERMINISTIC)
            gd_b[(unsigned)gy * (unsigned)W + (unsigned)gx_own] += g_dup;
          else
            atomicAdd(gd_b + ((unsigned)gy * (unsigned)W + (unsigned)gx_own), g_dup);
        } else if (sc.grad_dup) {
          sc.grad_dup[o] = g_dup;                           // deterministic mode: gathered by a second pass
        } else if (g_dup != 0.f) {
          const UpCoef cy = up_coef(gy, sc.hs, sc.up_sy), cx = up_coef(gx_own, sc.ws, sc.up_sx);
          atomicAdd(gd_b + cy.i0 * sc.ws + cx.i0, g_dup * cy.l0 * cx.l0);
          atomicAdd(gd_b + cy.i0 * sc.ws + cx.i1, g_dup * cy.l0 * cx.l1);
          atomicAdd(gd_b + cy.i1 * sc.ws + cx.i0, g_dup * cy.l1 * cx.l0);
          atomicAdd(gd_b + cy.i1 * sc.ws + cx.i1, g_dup * cy.l1 * cx.l1);
        }
      }
    }
    if (POSE) {
      // per source f and row r of dL/dP: (sum gc_r*depth*x, sum gc_r*depth*y, sum gc_r*depth, sum gc_r)
      const float fx = int_to_float(px_own);
      float v[24];
  #pragma unroll
      for (int r = 0; r < 3; ++r) {
        v[0 * 12 + r * 4 + 0] = Sw[r].x * fx;
        v[0 * 12 + r * 4 + 1] = Swy[r].x;
        v[0 * 12 + r * 4 + 2] = Sw[r].x;
        v[0 * 12 + r * 4 + 3] = Sg[r].x;
        v[1 * 12 + r * 4 + 0] = Sw[r].y * fx;
        v[1 * 12 + r * 4 + 1] = Swy[r].y;
        v[1 * 12 + r * 4 + 2] = Sw[r].y;
        v[1 * 12 + r * 4 + 3] = Sg[r].y;
      }
      const int lane = tid & 31, wid = tid >> 5;
      const float t = warp_sum24(v, lane);
      const int e = warp_sum24_index(lane);
      if (e < 24) sm.red[e][wid] = t;
      __syncthreads();
      if (tid < 24) {
        float t = 0.f;
        for (int w = 0; w < NT / 32; ++w) t += sm.red[tid][w];
        a.pose_partials[((size_t)tile_id * a.S + s) * 24 + tid] = t;
      }
    }
    if (POSE) __syncthreads();   // sm.red is reused by the next scale
  }
}

cudaError_t launch_vsl_backward(const VslArgs& a, cudaStream_t stream) {
  using Smem = BwdSmem<kBwdTileW, kBwdTileH>;
  static_assert(sizeof(Smem) <= 227 * 1024, "shared memory tile too large");
  static_assert(kBwdThreads / 32 <= 8, "red[] rows hold 8 warps");
  const int nblk = a.B * a.tiles_x * a.tiles_y + ((a.flags & PPEA_F_DETERMINISTIC) ? 0 : a.S * a.B * kSmoothChunks);
  cudaError_t e;
  if (a.flags & PPEA_F_GRAD_POSE) {
    auto kern = vsl_backward_kernel<kBwdTileW, kBwdTileH, kBwdThreads, true>;
    e = ensure_dynamic_smem(kern, (int)sizeof(Smem));
    if (e != cudaSuccess) return e;
    kern<<<nblk, kBwdThreads, sizeof(Smem), stream>>>(a);
  } else {
    auto kern = vsl_backward_kernel<kBwdTileW, kBwdTileH, kBwdThreads, false>;
    e = ensure_dynamic_smem(kern, (int)sizeof(Smem));
    if (e != cudaSuccess) return e;
    kern<<<nblk, kBwdThreads, sizeof(Smem), stream>>>(a);
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Deterministic mode, second pass: adjoint of the bilinear upsample in gather
// form.  One thread per coarse pixel sums, in a fixed order, the full-res
// gradients of every fine pixel whose interpolation footprint contains it.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample_gather_kernel(const __grid_constant__ VslArgs a, int s) {
  const ScaleArgs& sc = a.sc[s];
  const int n_s = sc.hs * sc.ws;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= a.B * n_s) return;
  const int b = idx / n_s, r = idx - b * n_s;
  const int jy = r / sc.ws, jx = r - jy * sc.ws;
  // fine pixels with source coordinate in (j-1, j+1): conservative bounds, exact test below
  const float inv_sy = 1.f / sc.up_sy, inv_sx = 1.f / sc.up_sx;
  const int ylo = max(0, (int)floorf(((float)jy - 0.5f) * inv_sy) - 2), yhi = min(a.H - 1, (int)ceilf(((float)jy + 1.5f) * inv_sy) + 2);
  const int xlo = max(0, (int)floorf(((float)jx - 0.5f) * inv_sx) - 2), xhi = min(a.W - 1, (int)ceilf(((float)jx + 1.5f) * inv_sx) + 2);
  const float* g = sc.grad_dup + (size_t)b * a.H * a.W;
  auto weight = [](const UpCoef& c, int j) {
    float w = 0.f;
    if (c.i0 == j) w += c.l0;
    if (c.i1 == j) w += c.l1;
    return w;
  };
  constexpr int kMaxTaps = 24;            // 2 * ratio + slack: covers ratios up to 8 (the reference's 4-scale pyramid)
  float acc = 0.f;
  if (xhi - xlo < kMaxTaps) {
    float wx[kMaxTaps];                   // the column weights do not depend on the row: hoisted out of the row loop
#pragma unroll
    for (int k = 0; k < kMaxTaps; ++k) wx[k] = (xlo + k <= xhi) ? weight(up_coef(xlo + k, sc.ws, sc.up_sx), jx) : 0.f;
    for (int y = ylo; y <= yhi; ++y) {
      const float wy = weight(up_coef(y, sc.hs, sc.up_sy), jy);
      if (wy == 0.f) continue;
      const float* row = g + (size_t)y * a.W + xlo;
      float racc = 0.f;
#pragma unroll
      for (int k = 0; k < kMaxTaps; ++k)
        if (wx[k] != 0.f) racc = fmaf(wx[k], row[k], racc);
      acc = fmaf(wy, racc, acc);
    }
  } else {                                // arbitrary ratios: same sums, same order, weights recomputed per tap
    for (int y = ylo; y <= yhi; ++y) {
      const float wy = weight(up_coef(y, sc.hs, sc.up_sy), jy);
      if (wy == 0.f) continue;
      float racc = 0.f;
      for (int x = xlo; x <= xhi; ++x) {
        const float w = weight(up_coef(x, sc.ws, sc.up_sx), jx);
        if (w != 0.f) racc = fmaf(w, g[(size_t)y * a.W + x], racc);
      }
      acc = fmaf(wy, racc, acc);
    }
  }
  sc.grad_disp[idx] += acc;
}

cudaError_t launch_upsample_gather(const VslArgs& a, cudaStream_t stream) {
  for (int s = 0; s < a.S; ++s) {
    if (!a.sc[s].grad_dup) continue;
    const int n = a.B * a.sc[s].hs * a.sc[s].ws;
    upsample_gather_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(a, s);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// ---------------------------------------------------------------------------
// Pose gradient: fixed-order reduction of the per-CTA d L / d P_f partials of each
// image, then d L / d T_f = K[:3,:]^T @ dL/dP_f  (autograd of layers.py:185).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * 24) pose_finish_kernel(const __grid_constant__ VslArgs a, int tiles) {
  __shared__ double Q[24];    // [f][r][(x, y, 1) moments of gc_r*depth, sum gc_r]
  __shared__ double gP[24];   // dL/dP_f, row-major 3x4
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, e = tid >> 5;   // warp e reduces entry e over the tiles
  {
    double t = 0;
    const float* p = a.pose_partials + (size_t)b * tiles * 24 + e;
    for (int i = lane; i < tiles; i += 32) t += (double)p[(size_t)i * 24];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) Q[e] = t;
  }
  __syncthreads();
  const float* K = a.K + b * 16;
  const float* iK = a.inv_K + b * 16;
  if (tid < 24) {
    // cam = depth * inv_K[:3,:3] (x,y,1)  =>  dL/dP[r][j] = sum_k Q[r][k] inv_K[j][k];  dL/dP[r][3] = sum gc_r
    const int f = tid / 12, ee = tid % 12, r = ee / 4, j = ee % 4;
    double t;
    if (j == 3) {
      t = Q[f * 12 + r * 4 + 3];
    } else {
      t = 0;
      for (int k = 0; k < 3; ++k) t += Q[f * 12 + r * 4 + k] * (double)iK[j * 4 + k];
    }
    gP[tid] = t;
  }
  __syncthreads();
  if (tid < 32) {
    const int f = tid / 16, ee = tid % 16, i = ee / 4, j = ee % 4;
    // (K3^T gP)[i][j] = sum_r K[r][i] * gP[r][j], r = 0..2
    double t = 0;
    for (int r = 0; r < 3; ++r) t += (double)K[r * 4 + i] * gP[f * 12 + r * 4 + j];
    a.grad_T[f][b * 16 + ee] = (float)t;
  }
}

cudaError_t launch_pose_finish(const VslArgs& a, int nblk_bwd, cudaStream_t stream) {
  pose_finish_kernel<<<a.B, 32 * 24, 0, stream>>>(a, nblk_bwd / a.B * a.S);
  return cudaGetLastError();
}

}  // namespace ppea
