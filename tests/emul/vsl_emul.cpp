// TEST TOOLING ONLY -- never part of the product path.
//
// Naive, whole-image, single-threaded CPU driver around the scalar functions
// of ppea_depth_b200/csrc/vsl_math.cuh.  The build container has no GPU, so
// tests/test_emul.py compiles this file with g++ and checks the *shared
// arithmetic* (forward values, selection, and above all the hand-derived
// backward formulas) against the oracle before any GPU time is spent.  The
// CUDA kernels call the same functions (their f2 instantiations run the two
// sources in lockstep) but tile/stage/reduce differently; their plumbing is
// validated by the `-m gpu` tests.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "vsl_math.cuh"

using namespace ppea;

namespace {

// reflect-padded 3x3 window sums of x, x*x, x*y at pixel (y, x) of channel planes
inline void window_sums(const float* X, const float* Y, int H, int W, int y, int x, float& Sx, float& Sxx, float& Sxy,
                        float& Sy, float& Syy) {
  float hx[3], hxx[3], hxy[3], hy[3], hyy[3];
  for (int dy = -1; dy <= 1; ++dy) {
    int yy = reflect_index(y + dy, H);
    float a[3], b[3];
    for (int dx = -1; dx <= 1; ++dx) {
      int xx = reflect_index(x + dx, W);
      a[dx + 1] = X[(size_t)yy * W + xx];
      b[dx + 1] = Y[(size_t)yy * W + xx];
    }
    row_sums_x<float>(a[0], a[1], a[2], b[0], b[1], b[2], hx[dy + 1], hxx[dy + 1], hxy[dy + 1]);
    row_sums_y<float>(b[0], b[1], b[2], hy[dy + 1], hyy[dy + 1]);
  }
  Sx = sum3(hx[0], hx[1], hx[2]);
  Sxx = sum3(hxx[0], hxx[1], hxx[2]);
  Sxy = sum3(hxy[0], hxy[1], hxy[2]);
  Sy = sum3(hy[0], hy[1], hy[2]);
  Syy = sum3(hyy[0], hyy[1], hyy[2]);
}

// photometric loss map 0.85*mean_c SSIM + 0.15*mean_c |y-x|  (trainer.py:995-1007)
void photometric_map(const float* X, const float* Y, int H, int W, bool no_ssim, float* out) {
  size_t plane = (size_t)H * W;
  const float w_l1 = no_ssim ? (1.f / 3.f) : PPEA_W_L1;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float acc = 0.f;
      for (int c = 0; c < 3; ++c) {
        const float* Xc = X + c * plane;
        const float* Yc = Y + c * plane;
        float l1 = fabsf(add_rn(Yc[(size_t)y * W + x], -Xc[(size_t)y * W + x]));
        acc = fma_rn(w_l1, l1, acc);
        if (!no_ssim) {
          float Sx, Sxx, Sxy, Sy, Syy;
          window_sums(Xc, Yc, H, W, y, x, Sx, Sxx, Sxy, Sy, Syy);
          acc = fma_rn(PPEA_W_SSIM, ssim_from_sums<float>(Sx, Sxx, Sxy, ssim_y_stats<float>(Sy, Syy)), acc);
        }
      }
      out[(size_t)y * W + x] = acc;
    }
}

struct PixelGeom {
  float A[2][3];
  ProjT<float> pr[2];
  Bilin bl[2];
};

inline PixelGeom pixel_geom(const Geom* g, float dep, int x, int y, float eps, int W, int H) {
  PixelGeom o;
  const float fx = (float)x, fy = (float)y;
  for (int f = 0; f < 2; ++f) {
    for (int i = 0; i < 3; ++i) o.A[f][i] = fma_rn(g[f].M[i * 3 + 1], fy, fma_rn(g[f].M[i * 3 + 0], fx, g[f].M[i * 3 + 2]));
    o.pr[f] = project_fast(dep, o.A[f][0], o.A[f][1], o.A[f][2], g[f].t[0], g[f].t[1], g[f].t[2], eps, coord_max(W), coord_max(H));
    o.bl[f] = bilin_setup(o.pr[f].ix, o.pr[f].iy, W);
  }
  return o;
}

}  // namespace

extern "C" {

// One scale, forward.  All arrays are host pointers with the layouts of include/ppea_vsl.h.
// out_sums: [sum(r*mask), sum(mask), sum(|depth-mono|*(1-mask))]
int emul_vsl_forward(int B, int H, int W, int h_s, int w_s, unsigned flags, float disp_lo, float disp_range, float eps,
                     const float* disp, const float* tgt, const float* src0, const float* src1, const float* K,
                     const float* inv_K, const float* T0, const float* T1, const float* noise, const float* cons_mask,
                     const float* aug_mask, const float* mono_depth, float* depth, float* loss_px, uint8_t* sel,
                     float* warped0, float* warped1, float* grid0, float* grid1, double* out_sums) {
  const bool multi = flags & 1u, automask = flags & 2u, selec = flags & 4u, no_ssim = flags & 8u;
  const bool motion = flags & 32u, aug = flags & 64u;
  const float* srcs[2] = {src0, src1};
  const float* Ts[2] = {T0, T1};
  float* warped[2] = {warped0, warped1};
  float* grids[2] = {grid0, grid1};
  size_t plane = (size_t)H * W;
  float sy = up_scale(h_s, H), sx = up_scale(w_s, W);
  double s_rm = 0, s_m = 0, s_c = 0;
  std::vector<float> L[2], Lid[2];
  for (int f = 0; f < 2; ++f) { L[f].resize(plane); Lid[f].resize(plane); }
  for (int b = 0; b < B; ++b) {
    Geom g[2];
    for (int f = 0; f < 2; ++f) compose_geom(K + b * 16, Ts[f] + b * 16, inv_K + b * 16, g[f]);
    const float* d_b = disp + (size_t)b * h_s * w_s;
    for (int y = 0; y < H; ++y) {
      UpCoef cy = up_coef(y, h_s, sy);
      for (int x = 0; x < W; ++x) {
        UpCoef cx = up_coef(x, w_s, sx);
        float dup = (h_s == H && w_s == W) ? d_b[(size_t)y * W + x] : up_sample(d_b, w_s, cy, cx);
        float dep = depth_from_disp(dup, disp_lo, disp_range);
        depth[b * plane + (size_t)y * W + x] = dep;
        PixelGeom pg = pixel_geom(g, dep, x, y, eps, W, H);
        for (int f = 0; f < 2; ++f) {
          if (grids[f]) {
            grids[f][(b * plane + (size_t)y * W + x) * 2 + 0] = (pg.pr[f].u / (float)(W - 1) - 0.5f) * 2.f;
            grids[f][(b * plane + (size_t)y * W + x) * 2 + 1] = (pg.pr[f].v / (float)(H - 1) - 0.5f) * 2.f;
          }
          const Bilin& bl = pg.bl[f];
          for (int c = 0; c < 3; ++c) {
            const float* S = srcs[f] + ((size_t)b * 3 + c) * plane;
            warped[f][((size_t)b * 3 + c) * plane + (size_t)y * W + x] =
                bilin_value(bl, S[bl.o00], S[bl.o00 + 1], S[bl.o00 + W], S[bl.o00 + W + 1]);
          }
        }
      }
    }
    const float* tg = tgt + (size_t)b * 3 * plane;
    for (int f = 0; f < 2; ++f) {
      photometric_map(warped[f] + (size_t)b * 3 * plane, tg, H, W, no_ssim, L[f].data());
      if (automask && !multi) photometric_map(srcs[f] + (size_t)b * 3 * plane, tg, H, W, no_ssim, Lid[f].data());
    }
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        size_t i = (size_t)y * W + x;
        float cs[2];
        for (int f = 0; f < 2; ++f) {
          const float* Wf = warped[f] + (size_t)b * 3 * plane;
          cs[f] = (Wf[i] + Wf[plane + i]) + Wf[2 * plane + i];
        }
        Select s = select_source(L[0][i], L[1][i], cs[0], cs[1], selec);
        float mask = 1.f;
        unsigned bits = (unsigned)s.src;
        if (multi) {
          if (motion) mask *= cons_mask[b * plane + i];
          if (aug) mask *= (1.f - aug_mask[b]);
          bits |= 4u;
          float dep = depth[b * plane + i];
          s_c += (double)(fabsf(dep - mono_depth[b * plane + i]) * (1.f - mask));
        } else if (automask) {
          float idl = fminf(Lid[0][i], Lid[1][i]);
          // min(dim=1) returns the first on ties; values identical either way
          if (noise) idl = idl + noise[b * plane + i] * 0.00001f;
          bool on = s.r <= idl;
          mask = on ? 1.f : 0.f;
          if (on) bits |= 4u;
        } else {
          bits |= 4u;
        }
        if (loss_px) loss_px[b * plane + i] = s.r;
        sel[b * plane + i] = (uint8_t)bits;
        s_rm += (double)(s.r * mask);
        s_m += (double)mask;
      }
  }
  out_sums[0] = s_rm;
  out_sums[1] = s_m;
  out_sums[2] = s_c;
  return 0;
}

// One scale, backward of
//   w_r * sum(r*mask)/(sum(mask)+1e-7)  +  w_c * mean(|depth-mono|*(1-mask))
// wrt disp_s and P_f = (K@T_f)[:3,:]; selection `sel` comes from the forward.
// g_r = w_r/(sum(mask)+1e-7), g_c = w_c/(B*H*W) are passed in.
int emul_vsl_backward(int B, int H, int W, int h_s, int w_s, unsigned flags, float disp_lo, float disp_range, float eps,
                      const float* disp, const float* tgt, const float* src0, const float* src1, const float* K,
                      const float* inv_K, const float* T0, const float* T1, const float* cons_mask,
                      const float* aug_mask, const float* mono_depth, const uint8_t* sel, float g_r, float g_c,
                      float* grad_disp, double* grad_P /* [B][2][12] */) {
  const bool multi = flags & 1u, no_ssim = flags & 8u, motion = flags & 32u, aug = flags & 64u;
  const float* srcs[2] = {src0, src1};
  const float* Ts[2] = {T0, T1};
  size_t plane = (size_t)H * W;
  float sy = up_scale(h_s, H), sx = up_scale(w_s, W);
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  const bool same = (h_s == H && w_s == W);
  memset(grad_disp, 0, sizeof(float) * (size_t)B * h_s * w_s);
  memset(grad_P, 0, sizeof(double) * (size_t)B * 24);
  std::vector<float> depth(plane), warped(2 * 3 * plane), gw(2 * 3 * plane);
  std::vector<float> cA(plane), cB(plane), cC(plane), wq(plane);
  for (int b = 0; b < B; ++b) {
    Geom g[2];
    for (int f = 0; f < 2; ++f) compose_geom(K + b * 16, Ts[f] + b * 16, inv_K + b * 16, g[f]);
    const float* d_b = disp + (size_t)b * h_s * w_s;
    // recompute depth + warped
    for (int y = 0; y < H; ++y) {
      UpCoef cy = up_coef(y, h_s, sy);
      for (int x = 0; x < W; ++x) {
        UpCoef cx = up_coef(x, w_s, sx);
        float dup = same ? d_b[(size_t)y * W + x] : up_sample(d_b, w_s, cy, cx);
        float dep = depth_from_disp(dup, disp_lo, disp_range);
        depth[(size_t)y * W + x] = dep;
        PixelGeom pg = pixel_geom(g, dep, x, y, eps, W, H);
        for (int f = 0; f < 2; ++f) {
          const Bilin& bl = pg.bl[f];
          for (int c = 0; c < 3; ++c) {
            const float* S = srcs[f] + ((size_t)b * 3 + c) * plane;
            warped[(f * 3 + c) * plane + (size_t)y * W + x] = bilin_value(bl, S[bl.o00], S[bl.o00 + 1], S[bl.o00 + W], S[bl.o00 + W + 1]);
          }
        }
      }
    }
    const float* tg = tgt + (size_t)b * 3 * plane;
    // dL/d warped_f,c(p)
    for (int f = 0; f < 2; ++f) {
      for (size_t i = 0; i < plane; ++i) {
        unsigned s = sel[b * plane + i];
        float mask;
        if (multi) {
          mask = 1.f;
          if (motion) mask *= cons_mask[b * plane + i];
          if (aug) mask *= (1.f - aug_mask[b]);
        } else {
          mask = (s & 4u) ? 1.f : 0.f;
        }
        wq[i] = ((int)(s & 3u) == f) ? g_r * mask : 0.f;
      }
      for (int c = 0; c < 3; ++c) {
        const float* Xc = warped.data() + (f * 3 + c) * plane;
        const float* Yc = tg + c * plane;
        for (int y = 0; y < H; ++y)
          for (int x = 0; x < W; ++x) {
            size_t i = (size_t)y * W + x;
            SsimAdjT<float> a = {0.f, 0.f, 0.f};
            if (!no_ssim && wq[i] != 0.f) {
              float Sx, Sxx, Sxy, Sy, Syy;
              window_sums(Xc, Yc, H, W, y, x, Sx, Sxx, Sxy, Sy, Syy);
              a = ssim_adjoint<float>(Sx, Sxx, Sxy, ssim_y_stats<float>(Sy, Syy), wq[i] * PPEA_W_SSIM);
            }
            cA[i] = a.cA; cB[i] = a.cB; cC[i] = a.cC;
          }
        // adjoint of reflect-pad + 3x3 box: tap (q+d) maps to reflect(q+d)
        float* G = gw.data() + (f * 3 + c) * plane;
        for (size_t i = 0; i < plane; ++i) {
          float l1w = no_ssim ? (1.f / 3.f) : PPEA_W_L1;
          G[i] = -wq[i] * l1w * sign_of(Yc[i] - Xc[i]);
        }
        if (!no_ssim)
          for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
              size_t q = (size_t)y * W + x;
              if (cA[q] == 0.f && cB[q] == 0.f && cC[q] == 0.f) continue;
              for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                  size_t p = (size_t)reflect_index(y + dy, H) * W + reflect_index(x + dx, W);
                  G[p] += cA[q] + cB[q] * Xc[p] + cC[q] * Yc[p];
                }
            }
      }
    }
    // chain through grid_sample, Project3D, BackprojectDepth, depth, upsample
    double Q[2][3][3];      // sum gc[r] * depth * (x, y, 1)
    double Q3[2][3];        // sum gc[r]
    memset(Q, 0, sizeof(Q));
    memset(Q3, 0, sizeof(Q3));
    for (int y = 0; y < H; ++y) {
      UpCoef cy = up_coef(y, h_s, sy);
      for (int x = 0; x < W; ++x) {
        UpCoef cx = up_coef(x, w_s, sx);
        size_t i = (size_t)y * W + x;
        float dep = depth[i];
        PixelGeom pg = pixel_geom(g, dep, x, y, eps, W, H);
        float g_depth = 0.f;
        for (int f = 0; f < 2; ++f) {
          const Bilin& bl = pg.bl[f];
          const ProjT<float>& pr = pg.pr[f];
          float gix = 0.f, giy = 0.f;
          for (int c = 0; c < 3; ++c) {
            const float* S = srcs[f] + ((size_t)b * 3 + c) * plane;
            float gv = gw[(f * 3 + c) * plane + i];
            gix += gv * bilin_ddx(bl, S[bl.o00], S[bl.o00 + 1], S[bl.o00 + W], S[bl.o00 + W + 1]);
            giy += gv * bilin_ddy(bl, S[bl.o00], S[bl.o00 + 1], S[bl.o00 + W], S[bl.o00 + W + 1]);
          }
          float gu = gix * clip_mask(pr.u, wm1), gv = giy * clip_mask(pr.v, hm1);
          float gc[3] = {gu * pr.rz, gv * pr.rz, -(gu * pr.u + gv * pr.v) * pr.rz};
          for (int r = 0; r < 3; ++r) {
            g_depth += gc[r] * pg.A[f][r];
            Q[f][r][0] += (double)(gc[r] * dep) * x;
            Q[f][r][1] += (double)(gc[r] * dep) * y;
            Q[f][r][2] += (double)(gc[r] * dep);
            Q3[f][r] += (double)gc[r];
          }
        }
        if (multi) {
          float mask = 1.f;
          if (motion) mask *= cons_mask[b * plane + i];
          if (aug) mask *= (1.f - aug_mask[b]);
          g_depth += g_c * sign_of(dep - mono_depth[b * plane + i]) * (1.f - mask);
        }
        float g_dup = g_depth * ddepth_ddisp(dep, disp_range);
        float* gd = grad_disp + (size_t)b * h_s * w_s;
        if (same) {
          gd[i] += g_dup;
        } else {
          gd[cy.i0 * w_s + cx.i0] += g_dup * cy.l0 * cx.l0;
          gd[cy.i0 * w_s + cx.i1] += g_dup * cy.l0 * cx.l1;
          gd[cy.i1 * w_s + cx.i0] += g_dup * cy.l1 * cx.l0;
          gd[cy.i1 * w_s + cx.i1] += g_dup * cy.l1 * cx.l1;
        }
      }
    }
    // dL/dP[r][j] = sum_j' Q[r][j'] * inv_K[j][j']  (cam = depth * inv_K3 (x,y,1)),  dL/dP[r][3] = sum gc[r]
    const float* iK = inv_K + b * 16;
    for (int f = 0; f < 2; ++f)
      for (int r = 0; r < 3; ++r) {
        for (int j = 0; j < 3; ++j) {
          double acc = 0;
          for (int k = 0; k < 3; ++k) acc += Q[f][r][k] * (double)iK[j * 4 + k];
          grad_P[(b * 2 + f) * 12 + r * 4 + j] = acc;
        }
        grad_P[(b * 2 + f) * 12 + r * 4 + 3] = Q3[f][r];
      }
  }
  return 0;
}

}  // extern "C"
