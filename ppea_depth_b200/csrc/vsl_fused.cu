// vsl_fused.cu -- single-pass training step of the view-synthesis loss (mono path): the forward
// products AND the un-normalised gradient fields of every pyramid scale in ONE launch.
//
// Why this is possible.  The only global quantity of the path is the masked-mean normaliser
// 1/(sum(mask_s) + 1e-7) (trainer.py:1113-1114) -- a scalar per scale that multiplies the whole
// photometric gradient field.  So the kernel that evaluates the loss of a tile can also push the
// adjoint of that tile through SSIM, grid_sample, Project3D, BackprojectDepth, disp_to_depth and the
// bilinear upsample with weight 1, and the backward proper shrinks to "multiply by
// upstream / (sum(mask_s) + 1e-7) and add the smoothness gradient" (vsl_grad_finish_kernel).  Compared
// with the forward + backward pair (vsl_fwd.cu / vsl_bwd.cu) the gather of both sources and the SSIM
// window sums are evaluated once per step instead of twice, and the per-channel phase structure
// (eight barriers per scale) becomes three phases per scale.
//
// A CTA owns a TW x TH tile of one image; lanes of an f2 are the two source frames.
//   once     stage the target (2-pixel reflection halo) and the un-warped sources; identity loss
//            (trainer.py:1060-1069) of every pixel q of the tile + 1-pixel halo.
//   per scale
//   gather   every cell of the tile + 2-pixel halo: upsample -> depth -> projection -> bilinear
//            samples of both sources into shared memory; own pixels keep d warped / d(u,v) in
//            registers and write depth.
//   pass Q   every q of the tile + 1-pixel halo: 3x3 window sums of the three channels -> SSIM + L1
//            loss of both sources -> min over sources / selec_reproj / automask (same arithmetic
//            in every CTA that sees q, so halo decisions agree with the owner's) -> the SSIM
//            adjoint coefficients (cA, cB, cC) of the SELECTED source only, as scalar planes, plus
//            the per-source indicator  mask(q) * [sel(q) == lane]; owners also write sel /
//            loss_px and the masked sums.
//   fold     d L / d warped_c(p) = sum_{q in 3x3(p)} ind(q) * (cA + cB x(p) + cC y(p))(q) + L1 term,
//            with the reflection multiplicities, by a sliding three-row window down the thread's
//            column; folded into d L / d (u, v) with the kept derivatives.
//   chain    projection / backprojection adjoint -> d L / d depth -> d L / d disp_up -> adjoint of
//            the upsample into the raw gradient of disp_s (plain store at scale 0, float atomics
//            above), pose partials.
#include "vsl_common.cuh"
#include <type_traits>

#include "smooth.cuh"
#include "vsl_gather.cuh"

namespace ppea {

template <int TW, int TH, int NT>
struct FusedSmem {
  static constexpr int RW = TW + 4, RH = TH + 4, RP = RW * RH;   // value region (2-pixel halo)
  static constexpr int QW = TW + 2, QH = TH + 2, QP = QW * QH;   // decision / coefficient region (1-pixel halo)
  // TMA needs a 16-byte aligned global start: the staged target / source tiles are YW = RW + 4 columns wide (two unused
  // columns on each side), region column j sits at tile column j + 2.
  static constexpr int YW = RW + 4, YP = YW * RH;
  f2 x[3][RP];          // warped samples (source 0, source 1).  Before the first gather the un-warped sources live here as
                        // two scalar plane triples [source][c][YP] (identity loss), spilling into cf (idle until pass Q)
  float cf[9][QP];      // [3 c + e]: cA, cB, cC of channel c for the source selected at q
  alignas(128) float y[3][YP];       // target (broadcast into both lanes at use)
  f2 ind[QP];           // mask(q) * [sel(q) == lane]
  float ident[QP];      // identity loss min_f photo(src_f, tgt)   (scale-invariant)
  f2 G[12];             // per-source geometry (vsl_math.cuh Geom), lanes = sources
  float redf[3][NT / 32];
  float red[24][NT / 32];
  unsigned long long mbar;                                   // arrival barrier of the TMA tile loads
  static constexpr int kSplit = 3 * YP;                      // second source's plane triple (floats)
  static_assert((kSplit * 4) % 128 == 0, "TMA destinations are 128-byte aligned");
  static_assert(2 * kSplit * 4 <= (int)(sizeof(f2) * 3 * RP + sizeof(float) * 9 * QP), "split source planes must fit in x + cf");
};

// ---- TMA / mbarrier primitives (PTX: cp.async.bulk.tensor, mbarrier)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// box at (x, y, plane) of a (planes, H, W) tensor -> dense [plane][row][col] tile in shared memory; out-of-range cells read 0
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int x, int y, int z, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<unsigned long long>(tm)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
               : "memory");
}

struct PhotoQ {
  f2 L;     // 0.85 mean_c SSIM + 0.15 mean_c |y - x|  (trainer.py:995-1007), both sources
  f2 cs;    // sum_c x (selec_reproj's darkness test, trainer.py:1078-1079)
};

// Photometric loss of both sources at one pixel q from its 3x3 window (xq / yq point at the window's
// top-left cell); with ADJ also the SSIM adjoint coefficients of the three channels (weight W_SSIM).
// Row / plane strides: XW, XP of the samples, YW, YP of the target.
// XSPLIT > 0: the two sources are scalar planes (xq viewed as float*, second source XSPLIT floats further).
template <bool ADJ, int XW, int XP, int YW, int YP, int XSPLIT = 0>
__device__ __forceinline__ PhotoQ photo_q(const f2* __restrict__ xq, const float* __restrict__ yq, bool no_ssim, float w_l1,
                                          f2 (&co)[9]) {
  PhotoQ o;
  o.L = dup2(0.f);
  o.cs = dup2(0.f);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const f2* xp = xq + c * XP;
    const float* x0p = reinterpret_cast<const float*>(xq) + c * XP;
    const float* yp = yq + c * YP;
    f2 Sx, Sxx, Sxy, Sy, Syy;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      f2 xa, xb, xc;
      if (XSPLIT > 0) {
        xa = mk2(x0p[dy * XW], x0p[XSPLIT + dy * XW]);
        xb = mk2(x0p[dy * XW + 1], x0p[XSPLIT + dy * XW + 1]);
        xc = mk2(x0p[dy * XW + 2], x0p[XSPLIT + dy * XW + 2]);
      } else {
        xa = xp[dy * XW], xb = xp[dy * XW + 1], xc = xp[dy * XW + 2];
      }
      const f2 ya = dup2(yp[dy * YW]), yb = dup2(yp[dy * YW + 1]), yc = dup2(yp[dy * YW + 2]);
      f2 hx, hxx, hxy, hy, hyy;
      row_sums_y<f2>(ya, yb, yc, hy, hyy);
      row_sums_x<f2>(xa, xb, xc, ya, yb, yc, hx, hxx, hxy);
      if (dy == 0) {
        Sx = hx, Sxx = hxx, Sxy = hxy, Sy = hy, Syy = hyy;
      } else {
        Sx = vadd(Sx, hx), Sxx = vadd(Sxx, hxx), Sxy = vadd(Sxy, hxy), Sy = vadd(Sy, hy), Syy = vadd(Syy, hyy);
      }
      if (dy == 1) {
        const f2 d = vsub(yb, xb);
        o.L.x = fma_rn(w_l1, fabsf(d.x), o.L.x);
        o.L.y = fma_rn(w_l1, fabsf(d.y), o.L.y);
        o.cs = vadd(o.cs, xb);
      }
    }
    if (!no_ssim) {
      const SsimYT<f2> yst = ssim_y_stats<f2>(Sy, Syy);
      const SsimTermsT<f2> t = ssim_terms<f2>(Sx, Sxx, Sxy, yst);
      const f2 inv_d = vrcp(vmul(t.d1, t.d2));
      const f2 R = vmul(vmul(t.n1, t.n2), inv_d);
      const f2 v = vfma(dup2(-0.5f), R, dup2(0.5f));
      o.L = vfma(dup2(PPEA_W_SSIM), mk2(clamp01(v.x), clamp01(v.y)), o.L);
      if (ADJ) {
        const f2 k = clamp_pass(v, vmul(dup2(-0.5f * PPEA_W_SSIM), inv_d));
        const f2 p = vmul(vmul(dup2(2.f), yst.s), vsub(t.n2, t.n1));
        const f2 q = vmul(vmul(vmul(dup2(2.f), R), Sx), vsub(t.d2, t.d1));
        co[c * 3 + 0] = vmul(k, vsub(p, q));
        co[c * 3 + 1] = vmul(k, vmul(vmul(dup2(-18.f), R), t.d1));
        co[c * 3 + 2] = vmul(k, vmul(dup2(18.f), t.n1));
      }
    } else if (ADJ) {
      co[c * 3 + 0] = co[c * 3 + 1] = co[c * 3 + 2] = dup2(0.f);
    }
  }
  return o;
}

// Deterministic mode: the coarse-scale gradient fields are accumulated as 64-bit fixed-point integers
// (2^-40 resolution, +-8.3e6 range): integer addition commutes, so the result does not depend on the
// order in which the atomics land, and no full-resolution scratch field / second pass is needed.
constexpr float kFixScale = 1099511627776.f;          // 2^40 (a power of two: the scaling itself is exact)
__device__ __forceinline__ void fixed_add(float* field, int idx, float v) {
  const long long q = __float2ll_rn(fminf(fmaxf(v, -8.3e6f), 8.3e6f) * kFixScale);
  atomicAdd(reinterpret_cast<unsigned long long*>(field) + idx, (unsigned long long)q);
}
__device__ __forceinline__ void scatter_fixed(float* field, float g, UpCoef cy, UpCoef cx, int ws) {
  if (g == 0.f) return;
  fixed_add(field, cy.i0 * ws + cx.i0, g * cy.l0 * cx.l0);
  fixed_add(field, cy.i0 * ws + cx.i1, g * cy.l0 * cx.l1);
  fixed_add(field, cy.i1 * ws + cx.i0, g * cy.l1 * cx.l0);
  fixed_add(field, cy.i1 * ws + cx.i1, g * cy.l1 * cx.l1);
}
__device__ __forceinline__ float fixed_value(unsigned long long acc) {
  return (float)((double)(long long)acc * (1.0 / (double)kFixScale));
}

#ifndef PPEA_FUSED_CTAS
#define PPEA_FUSED_CTAS 4
#endif
constexpr int kFusedTileW = 32;
constexpr int kFusedTileH = kFusedTileHc;
constexpr int kFusedThreads = PPEA_FUSED_THREADS;

template <int TW, int TH, int NT, bool POSE, bool MULTI, bool DET>
__global__ void __launch_bounds__(NT, PPEA_FUSED_CTAS) vsl_fused_kernel(const __grid_constant__ VslArgs a) {
  using Smem = FusedSmem<TW, TH, NT>;
  constexpr int RW = Smem::RW, RP = Smem::RP, QW = Smem::QW, YW = Smem::YW, YP = Smem::YP;
  constexpr int R = (TW * TH) / NT;
  static_assert(TW == 32 && (NT / 32) * R == TH && NT >= 128, "the gather/row mapping assumes lane == tile column and R rows per warp");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

  grid_launch_dependents();      // the dependent is the one-CTA finish kernel: let it take its place early
  // The smoothness roles are the LAST CTAs of the grid: they are short and fill the tail of the last
  // wave of tile CTAs (measured: -19 us against running them first).
  const int n_tiles_all = a.B * a.tiles_x * a.tiles_y;
  if ((int)blockIdx.x >= n_tiles_all) {
    smooth_fused_role(a, blockIdx.x - n_tiles_all, reinterpret_cast<float*>(smem_raw));
    return;
  }
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int blk = blockIdx.x;
  const int tile_id = blk;
  const int tx = blk % a.tiles_x;
  blk /= a.tiles_x;
  const int ty = blk % a.tiles_y;
  const int b = blk / a.tiles_y;
  const int x0 = tx * TW, y0 = ty * TH;
  const int H = a.H, W = a.W;
  const size_t plane = (size_t)H * W;
  const bool automask = !MULTI && (a.flags & PPEA_F_AUTOMASK);   // the multi path masks by consistency / augmentation instead
  constexpr bool det = DET;     // coarse-scale fields as 64-bit fixed-point accumulators (PPEA_F_DETERMINISTIC)
  // multi path (trainer.py:1101-1141): mask = consistency_mask * (1 - augmentation_mask[b])
  const bool motion = MULTI && (a.flags & PPEA_F_MOTION_MASK);
  const float one_minus_aug = (MULTI && (a.flags & PPEA_F_MATCH_AUG)) ? 1.f - a.aug_mask[b] : 1.f;
  const bool no_ssim = a.flags & PPEA_F_NO_SSIM;
  const bool selec = a.flags & PPEA_F_SELEC_REPROJ;
  const float l1w = no_ssim ? (1.f / 3.f) : PPEA_W_L1;

  if (tid < 24) {
    const int f = tid / 12, e = tid % 12;
    const float v = geom_entry(a.K + b * 16, a.T[f] + b * 16, a.inv_K + b * 16, e);
    (f ? sm.G[e].y : sm.G[e].x) = v;
  }

  const float* tgt_b = a.tgt + (size_t)b * 3 * plane;
  const float* src_b[2] = {a.src[0] + (size_t)b * 3 * plane, a.src[1] + (size_t)b * 3 * plane};

  // ---- stage the target (+ the un-warped sources for the identity loss) with a 2-pixel reflection halo.
  // Three TMA tile loads (cp.async.bulk.tensor.3d, box 3 planes x RH rows x YW columns starting at the 16-byte aligned
  // column x0 - 4) issued by one thread and awaited on an mbarrier; border tiles then patch their reflected halo.
  // Tensors TMA cannot describe (row pitch not a multiple of 16 bytes): the reflecting loop, into the same layout.
  float* const xs = reinterpret_cast<float*>(&sm.x[0][0]);       // identity-loss view of x (+ cf): [source][c][YP] scalars
  constexpr int kSplit = Smem::kSplit;
  const bool interior = x0 >= 2 && y0 >= 2 && x0 + TW + 2 <= W && y0 + TH + 2 <= H;      // (block-uniform)
  if (a.use_tma) {
    if (tid == 0) mbar_init(&sm.mbar, 1);
    __syncthreads();
    if (tid == 0) {
      constexpr unsigned kTileBytes = 3 * YP * sizeof(float);
      mbar_expect_tx(&sm.mbar, automask ? 3 * kTileBytes : kTileBytes);
      tma_load_3d(&sm.y[0][0], &a.tm_tgt, x0 - 4, y0 - 2, b * 3, &sm.mbar);
      if (automask) {
        tma_load_3d(xs, &a.tm_src[0], x0 - 4, y0 - 2, b * 3, &sm.mbar);
        tma_load_3d(xs + kSplit, &a.tm_src[1], x0 - 4, y0 - 2, b * 3, &sm.mbar);
      }
    }
    mbar_wait(&sm.mbar, 0);
    __syncthreads();             // (the geometry block above is read by every thread)
    if (!interior) {
      // TMA zero-fills what lies outside the image: the reflected halo (layers.py:238) of a border tile is copied from
      // the in-image cell it mirrors, which is part of the same tile (cells further out than the halo stay zero)
      for (int idx = tid; idx < RP; idx += NT) {
        const int i = idx / RW, j = idx - i * RW;
        const int gy = y0 - 2 + i, gx = x0 - 2 + j;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) continue;
        const int si = reflect_index(gy, H) - (y0 - 2), sj = reflect_index(gx, W) - (x0 - 2);
        if (si < 0 || si >= Smem::RH || sj < 0 || sj >= RW) continue;
        const int dst = i * YW + j + 2, src = si * YW + sj + 2;
#pragma unroll
        for (int c = 0; c < 3; ++c) sm.y[c][dst] = sm.y[c][src];
        if (automask) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            xs[c * YP + dst] = xs[c * YP + src];
            xs[kSplit + c * YP + dst] = xs[kSplit + c * YP + src];
          }
        }
      }
      __syncthreads();
    }
  } else {
    for (int idx = tid; idx < RP; idx += NT) {
      const int i = idx / RW, j = idx - i * RW;
      const int py = reflect_index(y0 - 2 + i, H), px = reflect_index(x0 - 2 + j, W);
      const size_t o = (size_t)py * W + px;
      const int yi = i * YW + j + 2;
#pragma unroll
      for (int c = 0; c < 3; ++c) sm.y[c][yi] = __ldg(tgt_b + c * plane + o);
      if (automask) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          xs[c * YP + yi] = __ldg(src_b[0] + c * plane + o);
          xs[kSplit + c * YP + yi] = __ldg(src_b[1] + c * plane + o);
        }
      }
    }
    __syncthreads();
  }

  // ---- identity loss of every q of the tile + 1 halo (trainer.py:1060-1069), once for all scales
  if (automask) {
    constexpr int QH = Smem::QH, NEX = (2 * QH + 31) / 32;      // same work items as pass Q below
#pragma unroll 1
    for (int item = wid; item < QH + NEX; item += NT / 32) {
      int i = item, j = lane;
      if (item >= QH) {
        const int e = (item - QH) * 32 + lane;
        if (e >= 2 * QH) break;
        i = e >> 1;
        j = 32 + (e & 1);
      }
      const int qi = i * QW + j;
      const int qy = y0 - 1 + i, qx = x0 - 1 + j;
      float idl = 0.f;
      if (qy >= 0 && qy < H && qx >= 0 && qx < W) {
        f2 unused[9];
        const PhotoQ ph = photo_q<false, YW, YP, YW, YP, kSplit>(reinterpret_cast<const f2*>(xs + i * YW + j + 2), &sm.y[0][i * YW + j + 2], no_ssim, l1w, unused);
        idl = fminf(ph.L.x, ph.L.y);
      }
      sm.ident[qi] = idl;
    }
    __syncthreads();   // the x planes are rewritten by the first gather
  }

  const float wmax = coord_max(W), hmax = coord_max(H);
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  const int col = tid % TW;
  const int row0 = (tid / TW) * R;
  const int gx_own = x0 + col;
  // this thread's tile column (reflect-clamped for partial tiles); the context (6 registers of folded projection) is
  // rebuilt from the geometry block where it is needed instead of living across the whole scale loop
  const int px_own = reflect_index(gx_own, W);
  // reflection multiplicities of the taps left/right of this thread's column

#pragma unroll 1
  for (int s = 0; s < a.S; ++s) {
    const ScaleArgs& sc = a.sc[s];
    const float* disp_b = sc.disp + (size_t)b * sc.hs * sc.ws;
    const bool same_res = (sc.hs == H && sc.ws == W);

    float dep[R];
    f2 ddx[R][3], ddy[R][3];     // d warped_c / d u, d warped_c / d v  (lanes = sources)
    f2 gu[R], gv[R];

    // ---- gather.  Warp w walks its own R tile rows (lane == tile column, derivatives kept) plus its share
    // of the four halo rows; the four halo columns are a flat list of extra cells.
    ColCtx cc = make_col(sm.G, gx_own, W);
    const SrcPlanes sp = make_planes(a.src[0] + (size_t)b * 3 * plane, a.src[1] + (size_t)b * 3 * plane, plane);   // (per scale: not kept live across pass Q / fold)
    if (!same_res) cc.cx = up_coef(cc.px, sc.ws, sc.up_sx);
    // the (upsampled) disparity of region row i at this thread's column: loads only, so the rows of a
    // thread can have them in flight together, ahead of the dependent source gathers
    auto load_disp = [&](int i) -> float {
      const int py = reflect_index(y0 - 2 + i, H);
      if (same_res) return __ldg(disp_b + ((unsigned)py * (unsigned)W + (unsigned)cc.px));
      return up_sample(disp_b, sc.ws, up_coef(py, sc.hs, sc.up_sy), cc.cx);
    };
    auto gather_row = [&](int i, float dup, auto want_deriv, f2 (&dx)[3], f2 (&dy)[3]) -> float {
      const int py = reflect_index(y0 - 2 + i, H);
      const float d = depth_from_disp_fast(dup, a.disp_lo, a.disp_range);
      f2 A[3], val[3];
      const ProjT<f2> pr = project_cell(sm.G, cc, py, d, a.eps, wmax, hmax, A);
      sample_sources<decltype(want_deriv)::value, false>(sp, W, pr, wm1, hm1, val, dx, dy);
      const int ridx = i * RW + col + 2;
#pragma unroll
      for (int c = 0; c < 3; ++c) sm.x[c][ridx] = val[c];
      return d;
    };
    {
      float* depth_b = sc.depth + (size_t)b * plane;
#pragma unroll
      for (int k = 0; k < R; ++k) dep[k] = load_disp(row0 + k + 2);
#pragma unroll
      for (int k = 0; k < R; ++k) {
        dep[k] = gather_row(row0 + k + 2, dep[k], std::true_type{}, ddx[k], ddy[k]);
        gu[k] = gv[k] = dup2(0.f);
        const int gy = y0 + row0 + k;
        if (gy < H && gx_own < W) depth_b[(unsigned)gy * (unsigned)W + (unsigned)gx_own] = dep[k];   // trainer.py:893
      }
    }
    {
      f2 u0[3], u1[3];
      constexpr int NWB = NT / 32;
      if (wid == 0 || wid == NWB - 1) {
        const int base = (wid == 0) ? 0 : Smem::RH - 2;
        const float da = load_disp(base), db = load_disp(base + 1);
        gather_row(base, da, std::false_type{}, u0, u1);
        gather_row(base + 1, db, std::false_type{}, u0, u1);
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
    __syncthreads();

    // ---- pass Q: loss, selection, mask and adjoint coefficients of every q of the tile + 1 halo
    float s_rm = 0.f, s_m = 0.f;
    // Work items: one Q row per warp (lane == column, so every shared-memory access of the warp is one
    // contiguous, bank-conflict-free run), then the two right-most Q columns as a flat list of cells.
    constexpr int QH = Smem::QH, NEX = (2 * QH + 31) / 32;
#pragma unroll 1
    for (int item = wid; item < QH + NEX; item += NT / 32) {
      int i = item, j = lane;
      if (item >= QH) {
        const int e = (item - QH) * 32 + lane;
        if (e >= 2 * QH) break;
        i = e >> 1;
        j = 32 + (e & 1);
      }
      const int qi = i * QW + j;
      const int qy = y0 - 1 + i, qx = x0 - 1 + j;
      f2 indv = dup2(0.f);
      float cfo[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) cfo[e] = 0.f;
      if (qy >= 0 && qy < H && qx >= 0 && qx < W) {
        const size_t o = (size_t)b * plane + (size_t)qy * W + qx;
        float nz = 0.f;
        if (automask && sc.noise) nz = __ldg(sc.noise + o);
        f2 co[9];
        const PhotoQ ph = photo_q<true, RW, RP, YW, YP>(&sm.x[0][i * RW + j], &sm.y[0][i * YW + j + 2], no_ssim, l1w, co);
        const Select sl = select_source(ph.L.x, ph.L.y, ph.cs.x, ph.cs.y, selec);
        bool on = true;
        float mq = 1.f;
        if (MULTI) {
          mq = (motion ? __ldg(a.cons_mask + o) : 1.f) * one_minus_aug;
        } else if (automask) {
          const float idl = sc.noise ? add_rn(sm.ident[qi], mul_rn(nz, 0.00001f)) : sm.ident[qi];   // trainer.py:1086-1087
          on = sl.r <= idl;                                                                       // argmin([r, id]) == 0
          mq = on ? 1.f : 0.f;
        }
        if (MULTI)
          indv = mk2(sl.src == 0 ? mq : 0.f, sl.src == 1 ? mq : 0.f);
        else if (on)
          indv = mk2(sl.src == 0 ? 1.f : 0.f, sl.src == 1 ? 1.f : 0.f);
#pragma unroll
        for (int e = 0; e < 9; ++e) cfo[e] = (sl.src == 1) ? co[e].y : co[e].x;
        if (i >= 1 && i <= TH && j >= 1 && j <= TW) {      // owner of q: forward products
          if (sc.loss_px) sc.loss_px[o] = sl.r;
          sc.sel[o] = (uint8_t)((unsigned)sl.src | (on ? PPEA_SEL_AUTOMASK : 0u));
          if (MULTI) {
            s_rm = fmaf(sl.r, mq, s_rm);
            s_m += mq;
          } else if (on) {
            s_rm += sl.r;
            s_m += 1.f;
          }
        }
      }
      sm.ind[qi] = indv;
#pragma unroll
      for (int e = 0; e < 9; ++e) sm.cf[e][qi] = cfo[e];
    }
    s_rm = warp_sum(s_rm);
    s_m = warp_sum(s_m);
    if (lane == 0) {
      sm.redf[0][wid] = s_rm;
      sm.redf[1][wid] = s_m;
    }
    __syncthreads();
    if (tid < (MULTI ? 2 : 3)) {          // (multi path: entry 2, the consistency sum, follows the chain phase)
      float t = 0.f;
      if (tid < 2)
        for (int w = 0; w < NT / 32; ++w) t += sm.redf[tid][w];
      a.partials[((size_t)tile_id * a.S + s) * 4 + tid] = t;
    }

    // ---- fold: box sums of the coefficients (sliding window down this thread's column) -> d L / d (u, v)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      f2 h[3][3];
#pragma unroll
      for (int i = 0; i < R + 2; ++i) {
        const int sl = i % 3;
        if (!no_ssim) {
          const int qrow = (row0 + i) * QW + col;
          const f2 mL = dup2((gx_own == 1) ? 2.f : 1.f), mR = dup2((gx_own == W - 2) ? 2.f : 1.f);
          const f2 i0 = vmul(mL, sm.ind[qrow]), i1 = sm.ind[qrow + 1], i2 = vmul(mR, sm.ind[qrow + 2]);
#pragma unroll
          for (int e = 0; e < 3; ++e) {
            const float* cp = &sm.cf[c * 3 + e][qrow];
            h[sl][e] = vfma(dup2(cp[2]), i2, vfma(dup2(cp[1]), i1, vmul(dup2(cp[0]), i0)));
          }
        }
        if (i >= 2) {
          const int k = i - 2;
          const int gy = y0 + row0 + k;
          const int ridx = (row0 + k + 2) * RW + col + 2;
          const f2 xv = sm.x[c][ridx], yv = dup2(sm.y[c][(row0 + k + 2) * YW + col + 4]);
          const f2 wl = sm.ind[(row0 + k + 1) * QW + col + 1];
          const f2 d = vsub(yv, xv);
          // L1 term:  -ind * l1w * sign(y - x)
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

    // ---- chain: projection adjoint, depth -> disp, adjoint of the bilinear upsample (weight 1: raw gradient)
    const size_t img_off = (size_t)b * sc.hs * sc.ws;
    float* gr_b = sc.grad_raw + (det && !same_res ? 2 * img_off : img_off);          // (fixed-point fields: 8 bytes per pixel)
    float* gc_b = MULTI ? sc.grad_raw2 + (det && !same_res ? 2 * img_off : img_off) : nullptr;
    float s_c = 0.f;
    f2 Sw[POSE ? 3 : 1], Swy[POSE ? 3 : 1], Sg[POSE ? 3 : 1];
#pragma unroll
    for (int e = 0; e < (POSE ? 3 : 1); ++e) Sw[e] = Swy[e] = Sg[e] = dup2(0.f);
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int gy = y0 + row0 + k;
      if (gy < H && gx_own < W) {
        f2 A[3];
        const float depk = dep[k];
        const ProjT<f2> pr = project_cell(sm.G, make_col(sm.G, gx_own, W), gy, depk, a.eps, wmax, hmax, A);
        const f2 gc0 = vmul(gu[k], pr.rz), gc1 = vmul(gv[k], pr.rz);
        const f2 gc2 = vneg(vmul(vfma(gu[k], pr.u, vmul(gv[k], pr.v)), pr.rz));
        const f2 gd2 = vfma(gc2, A[2], vfma(gc1, A[1], vmul(gc0, A[0])));
        const float g = gd2.x + gd2.y;
        if (POSE) {
          const f2 dk = dup2(depk), fy = dup2(int_to_float(gy));
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
        const float dd = ddepth_ddisp(depk, a.disp_range);
        const float g_dup = g * dd;
        float c_dup = 0.f;
        if (MULTI) {
          // consistency term mean(|depth - mono_depth| * (1 - mask))  (trainer.py:1128-1132): its un-normalised
          // gradient goes to a field of its own (its upstream weight differs from the photometric one)
          const size_t o = (size_t)b * plane + (size_t)gy * W + gx_own;
          const float om = 1.f - (motion ? __ldg(a.cons_mask + o) : 1.f) * one_minus_aug;
          const float dm = depk - __ldg(sc.mono_depth + o);
          s_c = fmaf(fabsf(dm), om, s_c);
          c_dup = sign_of(dm) * om * dd;
        }
        if (same_res) {
          const unsigned o0 = (unsigned)gy * (unsigned)W + (unsigned)gx_own;
          gr_b[o0] = g_dup;     // sole owner: plain store, no pre-zero needed
          if (MULTI) gc_b[o0] = c_dup;
        } else if (g_dup != 0.f || (MULTI && c_dup != 0.f)) {
          const UpCoef cy = up_coef(gy, sc.hs, sc.up_sy), cx = up_coef(gx_own, sc.ws, sc.up_sx);
          if (!det) {
            if (!MULTI || g_dup != 0.f) {
              atomicAdd(gr_b + cy.i0 * sc.ws + cx.i0, g_dup * cy.l0 * cx.l0);
              atomicAdd(gr_b + cy.i0 * sc.ws + cx.i1, g_dup * cy.l0 * cx.l1);
              atomicAdd(gr_b + cy.i1 * sc.ws + cx.i0, g_dup * cy.l1 * cx.l0);
              atomicAdd(gr_b + cy.i1 * sc.ws + cx.i1, g_dup * cy.l1 * cx.l1);
            }
            if (MULTI && c_dup != 0.f) {
              atomicAdd(gc_b + cy.i0 * sc.ws + cx.i0, c_dup * cy.l0 * cx.l0);
              atomicAdd(gc_b + cy.i0 * sc.ws + cx.i1, c_dup * cy.l0 * cx.l1);
              atomicAdd(gc_b + cy.i1 * sc.ws + cx.i0, c_dup * cy.l1 * cx.l0);
              atomicAdd(gc_b + cy.i1 * sc.ws + cx.i1, c_dup * cy.l1 * cx.l1);
            }
          } else {
            scatter_fixed(gr_b, g_dup, cy, cx, sc.ws);
            if (MULTI) scatter_fixed(gc_b, c_dup, cy, cx, sc.ws);
          }
        }
      }
    }
    if (MULTI) {
      s_c = warp_sum(s_c);
      if (lane == 0) sm.redf[2][wid] = s_c;
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
      const float t = warp_sum24(v, lane);
      const int e = warp_sum24_index(lane);
      if (e < 24) sm.red[e][wid] = t;
    }
    __syncthreads();   // fold is done with x / cf / ind: the next scale may overwrite them; red[] is complete
    if (POSE && tid < 24) {
      float t = 0.f;
      for (int w = 0; w < NT / 32; ++w) t += sm.red[tid][w];
      a.pose_partials[((size_t)tile_id * a.S + s) * 24 + tid] = t;
    }
    if (MULTI && tid == 32) {
      float t = 0.f;
      for (int w = 0; w < NT / 32; ++w) t += sm.redf[2][w];
      a.partials[((size_t)tile_id * a.S + s) * 4 + 2] = t;
    }
    // (red[] is next written after two more barriers of the following scale)
  }
}

template <bool POSE, bool MULTI, bool DET>
static cudaError_t launch_vsl_fused_as(const VslArgs& a, cudaStream_t stream) {
  using Smem = FusedSmem<kFusedTileW, kFusedTileH, kFusedThreads>;
  static_assert(sizeof(Smem) <= 227 * 1024, "shared memory tile too large");
  const int nblk = a.B * a.tiles_x * a.tiles_y + a.S * a.B * kSmoothChunks;
  auto kern = vsl_fused_kernel<kFusedTileW, kFusedTileH, kFusedThreads, POSE, MULTI, DET>;
  const cudaError_t e = ensure_dynamic_smem(kern, (int)sizeof(Smem));
  if (e != cudaSuccess) return e;
  kern<<<nblk, kFusedThreads, sizeof(Smem), stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_vsl_fused(const VslArgs& a, cudaStream_t stream) {
  const bool det = a.flags & PPEA_F_DETERMINISTIC;
  if (a.flags & PPEA_F_MULTI)      // T is detached on the multi path (trainer.py:900-902)
    return det ? launch_vsl_fused_as<false, true, true>(a, stream) : launch_vsl_fused_as<false, true, false>(a, stream);
  if (a.flags & PPEA_F_GRAD_POSE)
    return det ? launch_vsl_fused_as<true, false, true>(a, stream) : launch_vsl_fused_as<true, false, false>(a, stream);
  return det ? launch_vsl_fused_as<false, false, true>(a, stream) : launch_vsl_fused_as<false, false, false>(a, stream);
}

// ---------------------------------------------------------------------------
// Backward proper of the fused step, one thread per pixel of disp_s, every scale in one launch --
// elementwise, because both gradient fields were laid down un-normalised by the forward launch:
//   grad_disp_s = w_s * raw_s + g_s * inv_b * ( st_s - inv_b * (X_b / N_x + Y_b / N_y) / (h*w) ),
//   w_s = (upstream of reproj_s) / (sum(mask_s) + 1e-7)                    trainer.py:1113-1114
//   g_s = upstream of the smoothness term, inv_b = 1 / (mean(disp_s[b]) + 1e-7)   trainer.py:1147-1149
// (st: smooth.cuh smooth_fused_role).  With REZERO the raw field of the coarse scales (accumulated
// atomically by the next step) is cleared on the way out, so a replayed plan needs no memset.
// ---------------------------------------------------------------------------
constexpr int kGradFinishThreads = 256;

// Pose gradient of the fused step (a role of the gradient-finish launch, its last B CTAs, one per image): the per-scale
// sums of the raw d L / d P_f partials (finish launch, vsl_fwd.cu pose_sums_role) weighted by
// w_s = upstream(reproj_s) / (sum(mask_s) + 1e-7), then d L / d T_f = K[:3,:]^T @ dL/dP_f  (autograd of layers.py:185).
__device__ __forceinline__ void pose_combine_role(const VslArgs& a, int b) {
  __shared__ double Q[24];
  __shared__ double gP[24];
  const int tid = threadIdx.x;
  if (tid < 24) {
    double t = 0;
    for (int s = 0; s < a.S; ++s) {
      const float w = scale_grads(a, s).reproj / (a.sums[(size_t)s * sums_stride(a.B) + 1] + 1e-7f);
      t += (double)w * a.pose_sums[((size_t)b * a.S + s) * 24 + tid];
    }
    Q[tid] = t;
  }
  __syncthreads();
  const float* K = a.K + b * 16;
  const float* iK = a.inv_K + b * 16;
  if (tid < 24) {
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
    double t = 0;
    for (int r = 0; r < 3; ++r) t += (double)K[r * 4 + i] * gP[f * 12 + r * 4 + j];
    a.grad_T[f][b * 16 + ee] = (float)t;
  }
}

// VEC = 4: four consecutive pixels per thread through 128-bit accesses (needs h*w % 4 == 0 for every scale,
// so that the four share an image, and 16-byte aligned grad_disp); VEC = 1 otherwise.
template <int VEC>
struct RawVec {
  float v[VEC];
};
// one un-normalised field at pixels idx .. idx+VEC-1; `fixed`: 64-bit fixed-point accumulators (deterministic
// mode, coarse scales)
template <int VEC>
__device__ __forceinline__ RawVec<VEC> load_raw(const float* field, unsigned idx, bool fixed) {
  RawVec<VEC> r;
  if (fixed) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(field) + idx;
    if (VEC == 4) {
      const ulonglong2 lo = *reinterpret_cast<const ulonglong2*>(f), hi = *reinterpret_cast<const ulonglong2*>(f + 2);
      r.v[0] = fixed_value(lo.x), r.v[1 % VEC] = fixed_value(lo.y), r.v[2 % VEC] = fixed_value(hi.x), r.v[3 % VEC] = fixed_value(hi.y);
    } else {
      r.v[0] = fixed_value(*f);
    }
  } else if (VEC == 4) {
    const float4 t = *reinterpret_cast<const float4*>(field + idx);
    r.v[0] = t.x, r.v[1 % VEC] = t.y, r.v[2 % VEC] = t.z, r.v[3 % VEC] = t.w;
  } else {
    r.v[0] = field[idx];
  }
  return r;
}
// clear what was read: the next step of a replayed plan accumulates into it again (PPEA_F_RAW_PREZEROED)
template <int VEC>
__device__ __forceinline__ void zero_raw(float* field, unsigned idx, bool fixed) {
  if (fixed) {
    unsigned long long* f = reinterpret_cast<unsigned long long*>(field) + idx;
    if (VEC == 4) {
      *reinterpret_cast<ulonglong2*>(f) = make_ulonglong2(0ull, 0ull);
      *reinterpret_cast<ulonglong2*>(f + 2) = make_ulonglong2(0ull, 0ull);
    } else {
      *f = 0ull;
    }
  } else if (VEC == 4) {
    *reinterpret_cast<float4*>(field + idx) = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    field[idx] = 0.f;
  }
}

template <int VEC>
__global__ void __launch_bounds__(kGradFinishThreads) vsl_grad_finish_kernel(const __grid_constant__ VslArgs a, int4 blk_end) {
  grid_dependency_wait();        // sums of the finish kernel (and, transitively, the fields of the main launch)
  // scale of this CTA (blk_end.{x,y,z,w}: first block index past the blocks of scale 0..3)
  const int blk = blockIdx.x;
  if (blk >= blk_end.w) {        // the last B CTAs combine the pose gradient of one image each (tiny)
    pose_combine_role(a, blk - blk_end.w);
    return;
  }
  const int s = (blk >= blk_end.x) + (blk >= blk_end.y) + (blk >= blk_end.z);
  const int blk0 = s == 0 ? 0 : (s == 1 ? blk_end.x : (s == 2 ? blk_end.y : blk_end.z));
  const ScaleArgs& sc = a.sc[s];
  const int h = sc.hs, w = sc.ws;
  const unsigned n = (unsigned)(h * w);
  const unsigned idx = ((unsigned)(blk - blk0) * kGradFinishThreads + threadIdx.x) * VEC;
  if (idx >= (unsigned)a.B * n) return;
  const unsigned b = idx / n;
  const float* row = a.sums + (size_t)s * sums_stride(a.B);
  const float* img_sums = row + PPEA_SUMS_PER_SCALE + 4 * b;                                   // (sum d, X_b, Y_b)
  const float inv = 1.f / (img_sums[0] / (float)n + 1e-7f);
  const ScaleGrads sg = scale_grads(a, s);
  const float cx = w > 1 ? 1.f / ((float)a.B * h * (w - 1)) : 0.f, cy = h > 1 ? 1.f / ((float)a.B * (h - 1) * w) : 0.f;
  const float mean_term = inv * (cx * img_sums[1] + cy * img_sums[2]) / (float)n;
  const float w_st = sg.smooth * inv;
  const float w_raw = sg.reproj / (row[1] + 1e-7f);
  const bool multi = a.flags & PPEA_F_MULTI;
  const float w_cons = multi ? sg.cons / ((float)a.B * (float)a.H * (float)a.W) : 0.f;        // plain mean (trainer.py:1132)
  const bool coarse = (h != a.H || w != a.W);
  const bool fixed = coarse && (a.flags & PPEA_F_DETERMINISTIC);
  const bool rezero = coarse && (a.flags & PPEA_F_RAW_PREZEROED);
  // streaming step, fixed-point fields: some contribution of this step was not representable (NaN / Inf / beyond +-8.3e6,
  // vsl_stream.cu fixed_add_s) -> the coarse-scale gradients are NaN, as the float-atomic fields would be where it landed
  const bool poisoned = fixed && !(a.flags & PPEA_F_FUSED_TILES) && a.fmt_flag != nullptr && a.fmt_flag[1] != 0u;
  // every load is issued before the first use (the kernel is latency-bound: one DRAM round trip, not three)
  RawVec<VEC> o;
  if (VEC == 4) {
    const float4 st = __ldg(reinterpret_cast<const float4*>(sc.grad_st + idx));
    o.v[0] = st.x, o.v[1 % VEC] = st.y, o.v[2 % VEC] = st.z, o.v[3 % VEC] = st.w;
  } else {
    o.v[0] = __ldg(sc.grad_st + idx);
  }
  const RawVec<VEC> raw = load_raw<VEC>(sc.grad_raw, idx, fixed);
  RawVec<VEC> rc;
#pragma unroll
  for (int k = 0; k < VEC; ++k) rc.v[k] = 0.f;
  if (multi) rc = load_raw<VEC>(sc.grad_raw2, idx, fixed);
#pragma unroll
  for (int k = 0; k < VEC; ++k) o.v[k] = fmaf(w_cons, rc.v[k], fmaf(w_raw, raw.v[k], w_st * (o.v[k] - mean_term)));
  if (poisoned) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) o.v[k] = __int_as_float(0x7fc00000);
  }
  if (VEC == 4)
    *reinterpret_cast<float4*>(sc.grad_disp + idx) = make_float4(o.v[0], o.v[1 % VEC], o.v[2 % VEC], o.v[3 % VEC]);
  else
    sc.grad_disp[idx] = o.v[0];
  if (rezero) {
    zero_raw<VEC>(sc.grad_raw, idx, fixed);
    if (multi) zero_raw<VEC>(sc.grad_raw2, idx, fixed);
  }
}

cudaError_t launch_vsl_grad_finish(const VslArgs& a, cudaStream_t stream) {
  bool vec = true;
  for (int s = 0; s < a.S; ++s)
    vec = vec && (a.sc[s].hs * a.sc[s].ws) % 4 == 0 && (reinterpret_cast<uintptr_t>(a.sc[s].grad_disp) & 15) == 0;
  const int per = kGradFinishThreads * (vec ? 4 : 1);
  int end[4] = {0, 0, 0, 0};
  int acc = 0;
  for (int s = 0; s < 4; ++s) {
    if (s < a.S) acc += ceil_div(a.B * a.sc[s].hs * a.sc[s].ws, per);
    end[s] = acc;
  }
  const int4 be = make_int4(end[0], end[1], end[2], end[3]);
  const bool pose = (a.flags & PPEA_F_GRAD_POSE) && !(a.flags & PPEA_F_MULTI);
  const int nblk = acc + (pose ? a.B : 0);
  if (vec) return launch_pdl(vsl_grad_finish_kernel<4>, dim3(nblk), dim3(kGradFinishThreads), 0, stream, a, be);
  return launch_pdl(vsl_grad_finish_kernel<1>, dim3(nblk), dim3(kGradFinishThreads), 0, stream, a, be);
}

}  // namespace ppea
