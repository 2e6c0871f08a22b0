// TEST TOOLING ONLY -- never part of the product path.
//
// Naive, whole-image, single-threaded CPU driver around the scalar functions
// of ppea_depth_b200/csrc/vsl_math.cuh.  The build container has no GPU, so
// tests/test_emul.py compiles this file with g++ and checks the *shared
// arithmetic* (forward values, selection, and above all the hand-derived
// backward formulas) against the oracle before any GPU time is spent.  The
// CUDA kernels call the same functions but tile/stage/reduce differently;
// their plumbing is validated by the `-m gpu` tests.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "vsl_math.cuh"

using namespace ppea;

namespace {

struct Img {
  int H, W;
  const float* p;
  float at(int c, int y, int x) const { return p[((size_t)c * H + y) * W + x]; }
};

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
    hx[dy + 1] = a[0] + a[1] + a[2];
    hxx[dy + 1] = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
    hxy[dy + 1] = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    hy[dy + 1] = b[0] + b[1] + b[2];
    hyy[dy + 1] = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
  }
  Sx = hx[0] + hx[1] + hx[2];
  Sxx = hxx[0] + hxx[1] + hxx[2];
  Sxy = hxy[0] + hxy[1] + hxy[2];
  Sy = hy[0] + hy[1] + hy[2];
  Syy = hyy[0] + hyy[1] + hyy[2];
}

// photometric loss map 0.85*mean_c SSIM + 0.15*mean_c |y-x|  (trainer.py:995-1007)
void photometric_map(const float* X, const float* Y, int H, int W, bool no_ssim, float* out) {
  size_t plane = (size_t)H * W;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float acc = 0.f;
      for (int c = 0; c < 3; ++c) {
        const float* Xc = X + c * plane;
        const float* Yc = Y + c * plane;
        float l1 = fabsf(Yc[(size_t)y * W + x] - Xc[(size_t)y * W + x]);
        if (no_ssim) {
          acc += l1 * (1.f / 3.f);
        } else {
          float Sx, Sxx, Sxy, Sy, Syy;
          window_sums(Xc, Yc, H, W, y, x, Sx, Sxx, Sxy, Sy, Syy);
          SsimY ys = ssim_y_stats(Sy, Syy);
          acc += PPEA_W_SSIM * ssim_from_sums(Sx, Sxx, Sxy, ys) + PPEA_W_L1 * l1;
        }
      }
      out[(size_t)y * W + x] = acc;
    }
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
  float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  double s_rm = 0, s_m = 0, s_c = 0;
  std::vector<float> L[2], Lid[2];
  for (int f = 0; f < 2; ++f) { L[f].resize(plane); Lid[f].resize(plane); }
  for (int b = 0; b < B; ++b) {
    float P[2][12], iK[9];
    for (int f = 0; f < 2; ++f) compose_P(K + b * 16, Ts[f] + b * 16, P[f]);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) iK[i * 3 + j] = inv_K[b * 16 + i * 4 + j];
    const float* d_b = disp + (size_t)b * h_s * w_s;
    for (int y = 0; y < H; ++y) {
      UpCoef cy = up_coef(y, h_s, sy);
      for (int x = 0; x < W; ++x) {
        UpCoef cx = up_coef(x, w_s, sx);
        float dup = up_sample(d_b, w_s, cy, cx);
        float dep = depth_from_disp(dup, disp_lo, disp_range);
        depth[b * plane + (size_t)y * W + x] = dep;
        float ray[3], cam[3];
        pixel_ray(iK, (float)x, (float)y, ray);
        for (int j = 0; j < 3; ++j) cam[j] = mul_rn(dep, ray[j]);
        for (int f = 0; f < 2; ++f) {
          Proj pr = project_point(P[f], cam, eps, wm1, hm1);
          if (grids[f]) {
            grids[f][(b * plane + (size_t)y * W + x) * 2 + 0] = pr.gx;
            grids[f][(b * plane + (size_t)y * W + x) * 2 + 1] = pr.gy;
          }
          Bilin bl = bilin_setup(pr.ix, pr.iy, W, H);
          for (int c = 0; c < 3; ++c) {
            const float* S = srcs[f] + ((size_t)b * 3 + c) * plane;
            warped[f][((size_t)b * 3 + c) * plane + (size_t)y * W + x] =
                bilin_value(bl, S[bl.o00], S[bl.o01], S[bl.o10], S[bl.o11]);
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
// wrt disp_s and P_f = (K@T_f)[:3,:]; selection `sel` and `warped` come from the forward.
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
  float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  memset(grad_disp, 0, sizeof(float) * (size_t)B * h_s * w_s);
  memset(grad_P, 0, sizeof(double) * (size_t)B * 24);
  std::vector<float> depth(plane), warped(2 * 3 * plane), gw(2 * 3 * plane);
  std::vector<float> cA(plane), cB(plane), cC(plane), wq(plane);
  std::vector<double> gdu(plane);
  for (int b = 0; b < B; ++b) {
    float P[2][12], iK[9];
    for (int f = 0; f < 2; ++f) compose_P(K + b * 16, Ts[f] + b * 16, P[f]);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) iK[i * 3 + j] = inv_K[b * 16 + i * 4 + j];
    const float* d_b = disp + (size_t)b * h_s * w_s;
    // recompute depth + warped
    for (int y = 0; y < H; ++y) {
      UpCoef cy = up_coef(y, h_s, sy);
      for (int x = 0; x < W; ++x) {
        UpCoef cx = up_coef(x, w_s, sx);
        float dep = depth_from_disp(up_sample(d_b, w_s, cy, cx), disp_lo, disp_range);
        depth[(size_t)y * W + x] = dep;
        float ray[3], cam[3];
        pixel_ray(iK, (float)x, (float)y, ray);
        for (int j = 0; j < 3; ++j) cam[j] = mul_rn(dep, ray[j]);
        for (int f = 0; f < 2; ++f) {
          Proj pr = project_point(P[f], cam, eps, wm1, hm1);
          Bilin bl = bilin_setup(pr.ix, pr.iy, W, H);
          for (int c = 0; c < 3; ++c) {
            const float* S = srcs[f] + ((size_t)b * 3 + c) * plane;
            warped[(f * 3 + c) * plane + (size_t)y * W + x] = bilin_value(bl, S[bl.o00], S[bl.o01], S[bl.o10], S[bl.o11]);
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
            SsimAdj a = {0.f, 0.f, 0.f};
            if (!no_ssim && wq[i] != 0.f) {
              float Sx, Sxx, Sxy, Sy, Syy;
              window_sums(Xc, Yc, H, W, y, x, Sx, Sxx, Sxy, Sy, Syy);
              a = ssim_adjoint(Sx, Sxx, Sxy, ssim_y_stats(Sy, Syy), wq[i] * PPEA_W_SSIM);
            }
            cA[i] = a.cA; cB[i] = a.cB; cC[i] = a.cC;
          }
        // adjoint of reflect-pad + 3x3 box: tap (q+d) maps to reflect(q+d)
        float* G = gw.data() + (f * 3 + c) * plane;
        for (size_t i = 0; i < plane; ++i) {
          float l1w = no_ssim ? (1.f / 3.f) : PPEA_W_L1;
          float d = Yc[i] - Xc[i];
          float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
          G[i] = -wq[i] * l1w * sgn;
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
    for (int y = 0; y < H; ++y) {
      UpCoef cy = up_coef(y, h_s, sy);
      for (int x = 0; x < W; ++x) {
        UpCoef cx = up_coef(x, w_s, sx);
        size_t i = (size_t)y * W + x;
        float dep = depth[i];
        float ray[3], cam[3];
        pixel_ray(iK, (float)x, (float)y, ray);
        for (int j = 0; j < 3; ++j) cam[j] = mul_rn(dep, ray[j]);
        float g_depth = 0.f;
        for (int f = 0; f < 2; ++f) {
          Proj pr = project_point(P[f], cam, eps, wm1, hm1);
          Bilin bl = bilin_setup(pr.ix, pr.iy, W, H);
          float gix = 0.f, giy = 0.f;
          for (int c = 0; c < 3; ++c) {
            const float* S = srcs[f] + ((size_t)b * 3 + c) * plane;
            float g = gw[(f * 3 + c) * plane + i];
            gix += g * bilin_ddx(bl, S[bl.o00], S[bl.o01], S[bl.o10], S[bl.o11]);
            giy += g * bilin_ddy(bl, S[bl.o00], S[bl.o01], S[bl.o10], S[bl.o11]);
          }
          float gu = gix * pr.mx, gv = giy * pr.my;
          float inv_z = 1.f / pr.z;
          float gc[3] = {gu * inv_z, gv * inv_z, -(gu * pr.u + gv * pr.v) * inv_z};
          for (int j = 0; j < 3; ++j) {
            float gcam = P[f][0 * 4 + j] * gc[0] + P[f][1 * 4 + j] * gc[1] + P[f][2 * 4 + j] * gc[2];
            g_depth += gcam * ray[j];
          }
          for (int r = 0; r < 3; ++r) {
            for (int j = 0; j < 3; ++j) grad_P[(b * 2 + f) * 12 + r * 4 + j] += (double)(gc[r] * cam[j]);
            grad_P[(b * 2 + f) * 12 + r * 4 + 3] += (double)gc[r];
          }
        }
        if (multi) {
          float mask = 1.f;
          if (motion) mask *= cons_mask[b * plane + i];
          if (aug) mask *= (1.f - aug_mask[b]);
          float d = dep - mono_depth[b * plane + i];
          float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
          g_depth += g_c * sgn * (1.f - mask);
        }
        float g_dup = g_depth * ddepth_ddisp(dep, disp_range);
        float* gd = grad_disp + (size_t)b * h_s * w_s;
        gd[cy.i0 * w_s + cx.i0] += g_dup * cy.l0 * cx.l0;
        gd[cy.i0 * w_s + cx.i1] += g_dup * cy.l0 * cx.l1;
        gd[cy.i1 * w_s + cx.i0] += g_dup * cy.l1 * cx.l0;
        gd[cy.i1 * w_s + cx.i1] += g_dup * cy.l1 * cx.l1;
      }
    }
  }
  return 0;
}

}  // extern "C"
