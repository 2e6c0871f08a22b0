// vsl_gather.cuh -- the per-cell "depth -> projection -> bilinear gather of both sources" step
// shared by the forward and backward kernels.
//
// Both kernels walk their shared-memory region row by row with lane == tile column, so everything
// that depends only on the column (reflected x, the x-part of the folded projection, the upsample
// coefficients along x) is computed once per thread (ColCtx) and everything that depends only on
// the row is warp-uniform.  The few halo columns left and right of the tile are handled as a flat
// list of extra cells with their own ColCtx.
#pragma once

#include "vsl_common.cuh"

namespace ppea {

struct ColCtx {
  int px;        // reflected source column of this region column
  f2 ax[3];      // M[i][0]*x + M[i][2] for both sources (lanes), i = 0..2
  UpCoef cx;     // upsample coefficients along x (unused at full resolution)
};

__device__ __forceinline__ ColCtx make_col(const f2* __restrict__ G, int gx, int W) {
  ColCtx c;
  c.px = reflect_index(gx, W);
  const f2 fx = dup2(int_to_float(c.px));
  c.ax[0] = vfma(G[0], fx, G[2]);
  c.ax[1] = vfma(G[3], fx, G[5]);
  c.ax[2] = vfma(G[6], fx, G[8]);
  c.cx.i0 = c.cx.i1 = 0;
  c.cx.l0 = c.cx.l1 = 0.f;
  return c;
}

// depth of the full-resolution pixel (py, px): bilinear upsample of disp_s (trainer.py:886-887), layers.py:21-22
template <bool SAME_RES>
__device__ __forceinline__ float depth_of(const float* __restrict__ disp_b, int W, int ws, int py, const ColCtx& col,
                                          const UpCoef& cy, float disp_lo, float disp_range) {
  float dup;
  if (SAME_RES) {
    dup = __ldg(disp_b + ((unsigned)py * (unsigned)W + (unsigned)col.px));
  } else {
    dup = up_sample(disp_b, ws, cy, col.cx);
  }
  return depth_from_disp_fast(dup, disp_lo, disp_range);
}

// Bilinear setup of both sources in lockstep.
struct Bilin2 {
  int o0, o1;               // element offsets of the north-west corners (source 0, source 1)
  f2 wnw, wne, wsw, wse;    // corner weights
  f2 tx, ty, ex, ey;        // fractional position and complements
};
__device__ __forceinline__ Bilin2 bilin_setup2(const ProjT<f2>& pr, int W) {
  const FloorIF x0 = floor_if(pr.ix.x), x1 = floor_if(pr.ix.y), y0 = floor_if(pr.iy.x), y1 = floor_if(pr.iy.y);
  Bilin2 b;
  b.tx = vsub(pr.ix, mk2(x0.f, x1.f));
  b.ty = vsub(pr.iy, mk2(y0.f, y1.f));
  b.ex = vsub(dup2(1.f), b.tx);
  b.ey = vsub(dup2(1.f), b.ty);
  b.wnw = vmul(b.ey, b.ex);
  b.wne = vmul(b.ey, b.tx);
  b.wsw = vmul(b.ty, b.ex);
  b.wse = vmul(b.ty, b.tx);
  b.o0 = y0.i * W + x0.i;
  b.o1 = y1.i * W + x1.i;
  return b;
}

// Image base pointers of both sources.  Every gather address is base + a 32-bit element offset
// (3*H*W < 2^32), which the compiler turns into ONE IMAD.WIDE.U32 per distinct address.
struct SrcPlanes {
  const float* s0;
  const float* s1;
  unsigned plane;     // H*W
};
__device__ __forceinline__ SrcPlanes make_planes(const float* s0, const float* s1, size_t plane) {
  SrcPlanes o;
  // Opaque to the optimiser: otherwise it folds the (64-bit) image offset into every gather address and
  // each of the 24 loads of a cell gets its own 64-bit add chain instead of one IMAD.WIDE.U32.
  unsigned long long p0 = reinterpret_cast<unsigned long long>(s0), p1 = reinterpret_cast<unsigned long long>(s1);
  asm volatile("" : "+l"(p0), "+l"(p1));
  o.s0 = reinterpret_cast<const float*>(p0);
  o.s1 = reinterpret_cast<const float*>(p1);
  o.plane = (unsigned)plane;
  return o;
}

// Samples the three channels of both sources at `pr`; optionally also d value / d (u, v) with the
// border-clip masks folded in (GridSampler.h clip_coordinates_set_grad).
// SHARE: the warp's lanes are horizontally adjacent pixels of one row (and all 32 are active).  With a
// smooth flow lane j+1 samples one pixel to the right of lane j, so its west corners ARE lane j's east
// corners: those come by warp shuffle and only lanes where the footprints do not line up (flow
// discontinuities, lane 31) load them.  Halves the gather traffic through the L1 data pipe, the busiest
// unit of both kernels (ncu: l1tex data-pipe wavefronts ~60 % of peak).
template <bool DERIV, bool SHARE>
__device__ __forceinline__ void sample_sources(const SrcPlanes& sp, int W, const ProjT<f2>& pr, float wm1, float hm1,
                                               f2 (&val)[3], f2 (&ddx)[3], f2 (&ddy)[3]) {
  const Bilin2 b = bilin_setup2(pr, W);
  f2 eym, tym, exm, txm;
  if (DERIV) {
    const f2 mx = mk2(clip_mask(pr.u.x, wm1), clip_mask(pr.u.y, wm1)), my = mk2(clip_mask(pr.v.x, hm1), clip_mask(pr.v.y, hm1));
    eym = vmul(b.ey, mx);
    tym = vmul(b.ty, mx);
    exm = vmul(b.ex, my);
    txm = vmul(b.tx, my);
  }
  const unsigned uW = (unsigned)W;
  bool own0 = true, own1 = true;      // must this lane load its own east corners?
  if (SHARE) {
    const int lane = threadIdx.x & 31;
    const int n0 = __shfl_down_sync(0xffffffffu, b.o0, 1), n1 = __shfl_down_sync(0xffffffffu, b.o1, 1);
    own0 = (lane == 31) || (n0 != b.o0 + 1);
    own1 = (lane == 31) || (n1 != b.o1 + 1);
  }
#pragma unroll
  for (unsigned c = 0; c < 3; ++c) {
    const unsigned n0 = (unsigned)b.o0 + c * sp.plane, n1 = (unsigned)b.o1 + c * sp.plane;
    const f2 nw = mk2(__ldg(sp.s0 + n0), __ldg(sp.s1 + n1));
    const f2 sw = mk2(__ldg(sp.s0 + n0 + uW), __ldg(sp.s1 + n1 + uW));
    f2 ne, se;
    if (SHARE) {
      ne = mk2(__shfl_down_sync(0xffffffffu, nw.x, 1), __shfl_down_sync(0xffffffffu, nw.y, 1));
      se = mk2(__shfl_down_sync(0xffffffffu, sw.x, 1), __shfl_down_sync(0xffffffffu, sw.y, 1));
      if (own0) {
        ne.x = __ldg(sp.s0 + n0 + 1u);
        se.x = __ldg(sp.s0 + n0 + uW + 1u);
      }
      if (own1) {
        ne.y = __ldg(sp.s1 + n1 + 1u);
        se.y = __ldg(sp.s1 + n1 + uW + 1u);
      }
    } else {
      ne = mk2(__ldg(sp.s0 + n0 + 1u), __ldg(sp.s1 + n1 + 1u));
      se = mk2(__ldg(sp.s0 + n0 + uW + 1u), __ldg(sp.s1 + n1 + uW + 1u));
    }
    val[c] = vfma(se, b.wse, vfma(sw, b.wsw, vfma(ne, b.wne, vmul(nw, b.wnw))));
    if (DERIV) {
      ddx[c] = vfma(vsub(se, sw), tym, vmul(vsub(ne, nw), eym));   // bilin_ddx * clip mask
      ddy[c] = vfma(vsub(se, ne), txm, vmul(vsub(sw, nw), exm));   // bilin_ddy * clip mask
    }
  }
}

// projection of one pixel (row py, column context col) into both sources
__device__ __forceinline__ ProjT<f2> project_cell(const f2* __restrict__ G, const ColCtx& col, int py, float dep, float eps,
                                                  float wmax, float hmax, f2 (&A)[3]) {
  const f2 fy = dup2(int_to_float(py));
  A[0] = vfma(G[1], fy, col.ax[0]);
  A[1] = vfma(G[4], fy, col.ax[1]);
  A[2] = vfma(G[7], fy, col.ax[2]);
  return project_fast(dep, A[0], A[1], A[2], G[9], G[10], G[11], eps, wmax, hmax);
}

}  // namespace ppea
