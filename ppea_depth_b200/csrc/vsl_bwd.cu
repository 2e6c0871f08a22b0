// vsl_bwd.cu -- fused backward of the view-synthesis loss, all pyramid scales.
//
// Autograd of trainer.py:886-914 + 1050-1141 wrt every disp_s and (mono path)
// the two poses, recomputing the forward instead of storing it: the only
// forward products read back are the selection map `sel` (1 B/px) and the
// reduced sums.  HBM traffic per pixel and scale: tgt 12 B (once per CTA, all
// scales) + source gathers (24 B compulsory) + sel 1 B (+ cons_mask/mono_depth
// on the multi path) in; grad_disp_s 4/4^s B out.
//
// A CTA owns a TW x TH tile of one image.
//   phase 1  every cell of the tile + 2-pixel halo: depth -> backproject ->
//            project -> bilinear gather of both sources into shared memory
//            (halo cells hold the value of the reflected pixel, layers.py:238).
//            For its own R tile pixels a thread keeps d warped / d (ix, iy)
//            (clip-masked, GridSampler.h clip_coordinates_set_grad) in registers.
//   per source f:
//   phase 2  every pixel q of the tile + 1-pixel halo whose loss was taken from
//            source f (sel) and is unmasked: SSIM window sums -> adjoint
//            coefficients (cA,cB,cC) per channel; zero elsewhere.
//   phase 3  d L / d warped_f(p) = A + B x(p) + C y(p) + L1 term, where A,B,C
//            are 3x3 box sums of the coefficients with the reflection
//            multiplicities (a border window counts its mirrored tap twice),
//            computed with a sliding three-row window in registers; then the
//            chain through grid_sample, Project3D (layers.py:185-194) and
//            BackprojectDepth (layers.py:164-166) to d L / d depth and the
//            3x4 projection-matrix gradient partials.
//   phase 4  consistency term (multi path), d depth / d disp (layers.py:21-22),
//            adjoint of the bilinear upsample (trainer.py:886-887): direct
//            accumulate at scale 0, float atomics or (deterministic) a
//            full-res scratch + gather pass for coarser scales.
#include "vsl_common.cuh"

namespace ppea {

template <int TW, int TH>
struct BwdSmem {
  static constexpr int RW = TW + 4, RH = TH + 4, RP = RW * RH;   // value region (2-pixel halo)
  static constexpr int QW = TW + 2, QH = TH + 2, QP = QW * QH;   // coefficient region (1-pixel halo)
  float y[3][RP];
  float x[2][3][RP];
  float cf[9][QP];      // [channel*3 + {A,B,C}]
  float wl[QP];         // weight of the L1 term at q (0 unless q selected the current source)
  float P[2][12];
  float iK[9];
  float red[24][4];
};

struct WarpDeriv {
  float dx[3], dy[3];   // d warped_c / d u, d warped_c / d v (already multiplied by the clip masks)
};

// depth -> cam point -> projection + gather for one pixel and one source
template <bool WANT_DERIV>
__device__ __forceinline__ void warp_pixel(const float* __restrict__ P, const float* cam, float eps, float wm1, float hm1,
                                           int W, int H, const float* __restrict__ src_b, size_t plane, float* val,
                                           WarpDeriv* der) {
  const Proj pr = project_point(P, cam, eps, wm1, hm1);
  const Bilin bl = bilin_setup(pr.ix, pr.iy, W, H);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* S = src_b + c * plane;
    const float nw = __ldg(S + bl.o00), ne = __ldg(S + bl.o01), sw = __ldg(S + bl.o10), se = __ldg(S + bl.o11);
    val[c] = bilin_value(bl, nw, ne, sw, se);
    if (WANT_DERIV) {
      der->dx[c] = bilin_ddx(bl, nw, ne, sw, se) * pr.mx;
      der->dy[c] = bilin_ddy(bl, nw, ne, sw, se) * pr.my;
    }
  }
}

template <int TW, int TH, int NT, bool POSE>
__global__ void __launch_bounds__(NT) vsl_backward_kernel(const __grid_constant__ VslArgs a) {
  using Smem = BwdSmem<TW, TH>;
  constexpr int RW = Smem::RW, RP = Smem::RP, QW = Smem::QW, QP = Smem::QP;
  constexpr int R = (TW * TH) / NT;
  constexpr int HALO = RP - TW * TH;
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
  const bool no_ssim = a.flags & PPEA_F_NO_SSIM;
  const float l1w = no_ssim ? (1.f / 3.f) : PPEA_W_L1;

  if (tid < 24) {
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

  // ---- stage the target with a 2-pixel reflection halo (shared by every scale)
  for (int idx = tid; idx < RP; idx += NT) {
    const int i = idx / RW, j = idx - i * RW;
    const int py = reflect_index(y0 - 2 + i, H), px = reflect_index(x0 - 2 + j, W);
    const size_t o = (size_t)py * W + px;
#pragma unroll
    for (int c = 0; c < 3; ++c) sm.y[c][idx] = __ldg(tgt_b + c * plane + o);
  }
  __syncthreads();

  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  float iK[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) iK[e] = sm.iK[e];
  const int col = tid % TW;
  const int row0 = (tid / TW) * R;
  const int gx_own = x0 + col;
  const float one_minus_aug = (multi && (a.flags & PPEA_F_MATCH_AUG)) ? 1.f - a.aug_mask[b] : 1.f;
  // reflection multiplicities of the taps left/right of this thread's column
  const float mL = (gx_own == 1) ? 2.f : 1.f, mR = (gx_own == W - 2) ? 2.f : 1.f;

  float gp[POSE ? 24 : 1];
#pragma unroll
  for (int e = 0; e < (POSE ? 24 : 1); ++e) gp[e] = 0.f;

#pragma unroll 1
  for (int s = 0; s < a.S; ++s) {
    const ScaleArgs& sc = a.sc[s];
    const float* disp_b = sc.disp + (size_t)b * sc.hs * sc.ws;
    const float* srow = a.sums + (size_t)s * sums_stride(a.B);
    const ScaleGrads sg = scale_grads(a, s);
    const float g_r = sg.reproj / (srow[1] + 1e-7f);           // d reproj_s / d (r*mask)(q)   trainer.py:1114
    const float g_c = sg.cons / ((float)a.B * (float)plane);   // d cons_s / d (|depth-mono|*(1-mask))(q)

    float dep[R];
    WarpDeriv der[R][2];
    float gdep[R];

    // ---- phase 1a: own tile pixels (keeps derivatives in registers)
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int py = reflect_index(y0 + row0 + k, H), px = reflect_index(gx_own, W);   // clamp for partial tiles
      const UpCoef cy = up_coef(py, sc.hs, sc.up_sy), cx = up_coef(px, sc.ws, sc.up_sx);
      dep[k] = depth_from_disp(up_sample(disp_b, sc.ws, cy, cx), a.disp_lo, a.disp_range);
      gdep[k] = 0.f;
      float ray[3], cam[3];
      pixel_ray(iK, (float)px, (float)py, ray);
#pragma unroll
      for (int e = 0; e < 3; ++e) cam[e] = mul_rn(dep[k], ray[e]);
      const int ridx = (row0 + k + 2) * RW + col + 2;
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        float val[3];
        warp_pixel<true>(sm.P[f], cam, a.eps, wm1, hm1, W, H, src_b[f], plane, val, &der[k][f]);
#pragma unroll
        for (int c = 0; c < 3; ++c) sm.x[f][c][ridx] = val[c];
      }
    }
    // ---- phase 1b: halo ring (top 2 rows, bottom 2 rows, then left/right 2 columns of the tile rows)
    for (int hidx = tid; hidx < HALO; hidx += NT) {
      int i, j;
      if (hidx < 2 * RW) {
        i = hidx / RW;
        j = hidx - i * RW;
      } else if (hidx < 4 * RW) {
        const int t = hidx - 2 * RW;
        i = TH + 2 + t / RW;
        j = t % RW;
      } else {
        const int t = hidx - 4 * RW;
        i = 2 + t / 4;
        const int jj = t % 4;
        j = jj < 2 ? jj : TW + jj;
      }
      const int py = reflect_index(y0 - 2 + i, H), px = reflect_index(x0 - 2 + j, W);
      const UpCoef cy = up_coef(py, sc.hs, sc.up_sy), cx = up_coef(px, sc.ws, sc.up_sx);
      const float d = depth_from_disp(up_sample(disp_b, sc.ws, cy, cx), a.disp_lo, a.disp_range);
      float ray[3], cam[3];
      pixel_ray(iK, (float)px, (float)py, ray);
#pragma unroll
      for (int e = 0; e < 3; ++e) cam[e] = mul_rn(d, ray[e]);
      const int ridx = i * RW + j;
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        float val[3];
        warp_pixel<false>(sm.P[f], cam, a.eps, wm1, hm1, W, H, src_b[f], plane, val, nullptr);
#pragma unroll
        for (int c = 0; c < 3; ++c) sm.x[f][c][ridx] = val[c];
      }
    }
    __syncthreads();

#pragma unroll
    for (int f = 0; f < 2; ++f) {   // unrolled: der[k][f] / gp[f*12+e] must be statically indexed registers
      // ---- phase 2: adjoint coefficients of the SSIM windows centred in the tile + 1 halo
      for (int qi = tid; qi < QP; qi += NT) {
        const int i = qi / QW, j = qi - i * QW;
        const int qy = y0 - 1 + i, qx = x0 - 1 + j;
        float wq = 0.f;
        if (qy >= 0 && qy < H && qx >= 0 && qx < W) {
          const size_t o = (size_t)b * plane + (size_t)qy * W + qx;
          const unsigned bits = sc.sel[o];
          if ((int)(bits & PPEA_SEL_SRC_MASK) == f) {
            float mask;
            if (multi) {
              mask = (a.flags & PPEA_F_MOTION_MASK) ? a.cons_mask[o] : 1.f;
              mask *= one_minus_aug;
            } else {
              mask = (bits & PPEA_SEL_AUTOMASK) ? 1.f : 0.f;
            }
            wq = g_r * mask;
          }
        }
        sm.wl[qi] = wq;
        if (wq != 0.f && !no_ssim) {
          const float gs = wq * PPEA_W_SSIM;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float* xp = &sm.x[f][c][i * RW + j];
            const float* yp = &sm.y[c][i * RW + j];
            float Sx = 0.f, Sxx = 0.f, Sxy = 0.f, Sy = 0.f, Syy = 0.f;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const float xa = xp[dy * RW], xb = xp[dy * RW + 1], xc = xp[dy * RW + 2];
              const float ya = yp[dy * RW], yb = yp[dy * RW + 1], yc = yp[dy * RW + 2];
              // same association as the forward's sliding window: horizontal triples, then rows
              const float hx = xa + xb + xc, hy = ya + yb + yc;
              const float hxx = xa * xa + xb * xb + xc * xc, hyy = ya * ya + yb * yb + yc * yc;
              const float hxy = xa * ya + xb * yb + xc * yc;
              Sx = dy ? Sx + hx : hx;
              Sy = dy ? Sy + hy : hy;
              Sxx = dy ? Sxx + hxx : hxx;
              Syy = dy ? Syy + hyy : hyy;
              Sxy = dy ? Sxy + hxy : hxy;
            }
            const SsimAdj ad = ssim_adjoint(Sx, Sxx, Sxy, ssim_y_stats(Sy, Syy), gs);
            sm.cf[c * 3 + 0][qi] = ad.cA;
            sm.cf[c * 3 + 1][qi] = ad.cB;
            sm.cf[c * 3 + 2][qi] = ad.cC;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 9; ++e) sm.cf[e][qi] = 0.f;
        }
      }
      __syncthreads();

      // ---- phase 3: box-sum the coefficients (sliding window down this thread's column), chain rule
      {
        float h[3][9];
#pragma unroll
        for (int i = 0; i < R + 2; ++i) {
          const int sl = i % 3;
          const int qrow = (row0 + i) * QW + col;
#pragma unroll
          for (int e = 0; e < 9; ++e) {
            const float* cp = &sm.cf[e][qrow];
            h[sl][e] = fmaf(mL, cp[0], fmaf(mR, cp[2], cp[1]));
          }
          if (i >= 2) {
            const int k = i - 2;
            const int gy = y0 + row0 + k;
            const float mU = (gy == 1) ? 2.f : 1.f, mD = (gy == H - 2) ? 2.f : 1.f;
            const int su = (i - 2) % 3, smid = (i - 1) % 3;
            const int ridx = (row0 + k + 2) * RW + col + 2;
            const float wl = sm.wl[(row0 + k + 1) * QW + col + 1];
            float gu = 0.f, gv = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float A = fmaf(mU, h[su][c * 3 + 0], fmaf(mD, h[sl][c * 3 + 0], h[smid][c * 3 + 0]));
              const float Bc = fmaf(mU, h[su][c * 3 + 1], fmaf(mD, h[sl][c * 3 + 1], h[smid][c * 3 + 1]));
              const float Cc = fmaf(mU, h[su][c * 3 + 2], fmaf(mD, h[sl][c * 3 + 2], h[smid][c * 3 + 2]));
              const float xv = sm.x[f][c][ridx], yv = sm.y[c][ridx];
              const float d = yv - xv;
              const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
              const float G = fmaf(Bc, xv, fmaf(Cc, yv, A)) - wl * l1w * sgn;   // d L / d warped_{f,c}(p)
              gu = fmaf(G, der[k][f].dx[c], gu);
              gv = fmaf(G, der[k][f].dy[c], gv);
            }
            if (gy < H && gx_own < W && (gu != 0.f || gv != 0.f)) {
              // Project3D / BackprojectDepth adjoint (recomputes the cheap projection)
              float ray[3], cam[3];
              pixel_ray(iK, (float)gx_own, (float)gy, ray);
#pragma unroll
              for (int e = 0; e < 3; ++e) cam[e] = mul_rn(dep[k], ray[e]);
              const float* P = sm.P[f];
              const Proj pr = project_point(P, cam, a.eps, wm1, hm1);
              const float inv_z = 1.f / pr.z;
              const float gc0 = gu * inv_z, gc1 = gv * inv_z, gc2 = -(gu * pr.u + gv * pr.v) * inv_z;
              float gd = 0.f;
#pragma unroll
              for (int e = 0; e < 3; ++e) gd = fmaf(fmaf(P[e], gc0, fmaf(P[4 + e], gc1, P[8 + e] * gc2)), ray[e], gd);
              gdep[k] += gd;
              if (POSE) {
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                  gp[f * 12 + 0 + e] = fmaf(gc0, cam[e], gp[f * 12 + 0 + e]);
                  gp[f * 12 + 4 + e] = fmaf(gc1, cam[e], gp[f * 12 + 4 + e]);
                  gp[f * 12 + 8 + e] = fmaf(gc2, cam[e], gp[f * 12 + 8 + e]);
                }
                gp[f * 12 + 3] += gc0;
                gp[f * 12 + 7] += gc1;
                gp[f * 12 + 11] += gc2;
              }
            }
          }
        }
      }
      __syncthreads();   // coefficient planes are rewritten by the next source / x planes by the next scale
    }

    // ---- phase 4: consistency term, depth -> disp, adjoint of the bilinear upsample
    float* gd_b = sc.grad_disp + (size_t)b * sc.hs * sc.ws;
    const bool same_res = (sc.hs == H && sc.ws == W);
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int gy = y0 + row0 + k;
      if (gy < H && gx_own < W) {
        const size_t o = (size_t)b * plane + (size_t)gy * W + gx_own;
        float g = gdep[k];
        if (multi) {
          float mask = (a.flags & PPEA_F_MOTION_MASK) ? a.cons_mask[o] : 1.f;
          mask *= one_minus_aug;
          const float d = dep[k] - sc.mono_depth[o];
          const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
          g = fmaf(g_c * sgn, 1.f - mask, g);
        }
        const float g_dup = g * ddepth_ddisp(dep[k], a.disp_range);
        if (same_res) {
          gd_b[(size_t)gy * W + gx_own] += g_dup;           // sole owner of this element (after smooth_backward)
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
  }

  if (POSE) {
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int e = 0; e < 24; ++e) {
      const float v = warp_sum(gp[e]);
      if (lane == 0) sm.red[e][wid] = v;
    }
    __syncthreads();
    if (tid < 24) {
      float t = 0.f;
      for (int w = 0; w < NT / 32; ++w) t += sm.red[tid][w];
      a.pose_partials[(size_t)blockIdx.x * 24 + tid] = t;
    }
  }
}

cudaError_t launch_vsl_backward(const VslArgs& a, cudaStream_t stream) {
  using Smem = BwdSmem<kBwdTileW, kBwdTileH>;
  static_assert(sizeof(Smem) <= 227 * 1024, "shared memory tile too large");
  static_assert(kBwdThreads / 32 <= 4, "red[] rows hold 4 warps");
  const int nblk = a.B * a.tiles_x * a.tiles_y;
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
  float acc = 0.f;
  for (int y = ylo; y <= yhi; ++y) {
    const UpCoef cy = up_coef(y, sc.hs, sc.up_sy);
    float wy = 0.f;
    if (cy.i0 == jy) wy += cy.l0;
    if (cy.i1 == jy) wy += cy.l1;
    if (wy == 0.f) continue;
    float racc = 0.f;
    for (int x = xlo; x <= xhi; ++x) {
      const UpCoef cx = up_coef(x, sc.ws, sc.up_sx);
      float wx = 0.f;
      if (cx.i0 == jx) wx += cx.l0;
      if (cx.i1 == jx) wx += cx.l1;
      if (wx != 0.f) racc = fmaf(wx, g[(size_t)y * a.W + x], racc);
    }
    acc = fmaf(wy, racc, acc);
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
__global__ void __launch_bounds__(32) pose_finish_kernel(const __grid_constant__ VslArgs a, int tiles) {
  __shared__ double gP[24];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < 24) {
    double t = 0;
    const float* p = a.pose_partials + (size_t)b * tiles * 24 + tid;
    for (int i = 0; i < tiles; ++i) t += (double)p[(size_t)i * 24];
    gP[tid] = t;
  }
  __syncwarp();
  const float* K = a.K + b * 16;
  {
    const int f = tid / 16, e = tid % 16, i = e / 4, j = e % 4;
    // (K3^T gP)[i][j] = sum_r K[r][i] * gP[r][j], r = 0..2
    double t = 0;
    for (int r = 0; r < 3; ++r) t += (double)K[r * 4 + i] * gP[f * 12 + r * 4 + j];
    a.grad_T[f][b * 16 + e] = (float)t;
  }
}

cudaError_t launch_pose_finish(const VslArgs& a, int nblk_bwd, cudaStream_t stream) {
  pose_finish_kernel<<<a.B, 32, 0, stream>>>(a, nblk_bwd / a.B);
  return cudaGetLastError();
}

}  // namespace ppea
