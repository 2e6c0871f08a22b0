// vsl_stream.cu -- the training step of the view-synthesis loss as a WARP-STREAMING kernel: forward
// products and un-normalised gradient fields of every pyramid scale (same contract as vsl_fused.cu,
// same C entry points), organised so that no window tap is ever re-read and no CTA barrier exists.
//
// Why (round-1 ncu of the tile kernel, profiles/README.md r1j): 1 500 lane-instructions per pixel and
// scale, every SSIM tap a shared-memory load (nine per window and channel, the three row sums of a
// window recomputed for every q), eight barriers per scale, 4 x 57 KB of shared memory starving the
// L1 that serves the gathers, a 70 KB loop body.  Here:
//
//   * a WARP owns a strip of 32 image columns (lane == column) and walks DOWN the rows of its
//     segment; a warp task is (image, scale, strip, row segment).  Warps never talk to each other:
//     no __syncthreads, no shared tile.
//   * every quantity is produced exactly once per (row, lane) and carried in registers: the warped
//     samples of a row, their horizontal 3-tap sums (x, x^2, x y per channel, both sources packed in
//     one f32x2), the target's row sums.  The 3x3 window sum is (row r-2) + (row r-1) + (row r) of
//     register-resident row sums -- two adds per quantity instead of eight adds + nine loads.
//     Horizontal neighbours come by warp shuffle (lanes 0 / 31 are the strip's halo columns, so a
//     strip yields 30 decided and 28 output columns).
//   * the pipeline is skewed by one row per stage: iteration r gathers row r, decides row r-1
//     (SSIM + L1 of both sources, min / selec_reproj / automask, SSIM adjoint coefficients) and
//     folds / chains row r-2 (3x3 box sums of the coefficients -> dL/d warped -> dL/d(u,v) ->
//     projection adjoint -> dL/d disp -> upsample adjoint, pose partials).  The two-row register
//     rings alternate between two named slots (loop unrolled by two), so nothing is ever moved.
//   * what does not depend on the scale is computed once per step by a small preparation launch
//     (vsl_prep_kernel): the identity-reprojection loss map (trainer.py:1060-1069) and the source
//     frames packed to ONE 32-bit RGBA8 word per pixel -- the loss frames are ToTensor of uint8 images
//     (datasets/mono_dataset.py:62,106), i.e. exactly k/255, which the kernel verifies value by value
//     (fmt_flag); a bilinear corner is then one 32-bit load for three channels instead of three, and
//     the gather footprint shrinks 3x.  Frames that are not k/255 take the planar fp32 gathers
//     (same kernel, warp-uniform branch) -- results then match the tile kernel bit for bit.
//   * the adjoint of the bilinear upsample is aggregated in registers while the walk stays on the
//     same coarse row, so a coarse scale issues 1 / 0.5 / 0.25 atomics per pixel instead of 4.
//
// Arithmetic (vsl_math.cuh) and summation order of the window sums are those of the tile kernel:
// the per-pixel maps are bit-identical on the planar path.
#include "vsl_common.cuh"
#include <cstdlib>
#include <type_traits>

#include "smooth.cuh"
#include "vsl_gather.cuh"

namespace ppea {

// One warp per CTA: the task then derives from blockIdx alone, so everything per task (image / scale base pointers,
// segment bounds, flags) lives in UNIFORM registers instead of costing two vector registers per pointer.
#ifndef PPEA_STREAM_WARPS
#define PPEA_STREAM_WARPS 1
#endif
constexpr int kStreamWarps = PPEA_STREAM_WARPS;
constexpr int kStreamThreads = 32 * kStreamWarps;
#ifndef PPEA_PREP_SEG
#define PPEA_PREP_SEG 48
#endif
constexpr int kPrepSegRows = PPEA_PREP_SEG;      // rows per warp task of the preparation launch
constexpr int kPrepStripW = 30;       // decided columns per warp there
constexpr int kRingK = 12;            // float4 words per lane and ring slot of the main kernel (48 floats, see there)

__device__ __forceinline__ f2 shfl_up2(f2 v) {
  return mk2(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1));
}
__device__ __forceinline__ f2 shfl_down2(f2 v) {
  return mk2(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1));
}

// -(t) * sign(d) without conversions:  d > 0 -> -t,  d < 0 -> +t,  d == 0 -> 0   (one LOP3 + compare + select)
__device__ __forceinline__ float neg_signed(float t, float d) {
  const float r = __uint_as_float(__float_as_uint(t) ^ (~__float_as_uint(d) & 0x80000000u));
  return d == 0.f ? 0.f : r;
}

// SSIM + L1 of both sources (lanes of the f2) for ONE channel from its 3x3 window sums; xc / yc are the
// window centres.  With ADJ also the adjoint coefficients (cA, cB, cC) of this channel.  Same expressions, in
// the same order, as photo_q of the tile kernel; `w_ssim` is PPEA_W_SSIM, or 0 with opt.no_ssim (then the
// SSIM term adds an exact zero and the coefficients vanish: no branch).
template <bool ADJ>
__device__ __forceinline__ void photo_channel(f2 Sx, f2 Sxx, f2 Sxy, float Sy, float Syy, f2 xc, float yc, float w_ssim, float w_l1,
                                              f2& L, f2& cs, f2* __restrict__ co) {
  const f2 d = vsub(dup2(yc), xc);
  L.x = fma_rn(w_l1, fabsf(d.x), L.x);
  L.y = fma_rn(w_l1, fabsf(d.y), L.y);
  cs = vadd(cs, xc);
  const SsimYT<f2> yst = ssim_y_stats<f2>(dup2(Sy), dup2(Syy));
  const SsimTermsT<f2> t = ssim_terms<f2>(Sx, Sxx, Sxy, yst);
  const f2 inv_d = vrcp(vmul(t.d1, t.d2));
  const f2 R = vmul(vmul(t.n1, t.n2), inv_d);
  const f2 v = vfma(dup2(-0.5f), R, dup2(0.5f));
  L = vfma(dup2(w_ssim), mk2(clamp01(v.x), clamp01(v.y)), L);
  if (ADJ) {
    // torch.clamp passes the gradient where 0 <= v <= 1 (inclusive), i.e. |R| <= 1  (v = (1 - R) / 2 is exact enough:
    // every float R > 1 gives v <= -2^-24)
    f2 k = vmul(dup2(-0.5f * w_ssim), inv_d);
    k = mk2(fabsf(R.x) <= 1.f ? k.x : 0.f, fabsf(R.y) <= 1.f ? k.y : 0.f);
    const f2 p = vmul(vmul(dup2(2.f), yst.s), vsub(t.n2, t.n1));
    const f2 q = vmul(vmul(vmul(dup2(2.f), R), Sx), vsub(t.d2, t.d1));
    co[0] = vmul(k, vsub(p, q));
    co[1] = vmul(k, vmul(vmul(dup2(-18.f), R), t.d1));
    co[2] = vmul(k, vmul(dup2(18.f), t.n1));
  }
}

// ---------------------------------------------------------------------------------------------------
// Packed RGBA8 gathers.  A source pixel is one word r | g << 8 | b << 16 (k = round(255 x), exact for
// ToTensor frames).  A byte becomes a float through the mantissa: PRMT builds 0x4B0000kk = 2^23 + k,
// one packed add removes 2^23 for both sources.  The blend runs on k; the 1/255 is folded into the
// row weights, so  val = sum_i w_i k_i / 255  (the reference blends fl(k/255): same number to 1 ulp).
// The eight loads of a row are ISSUED one iteration ahead of their use (RowFetch), so the L2 latency of
// the gather is covered by a whole row of arithmetic.
// ---------------------------------------------------------------------------------------------------
template <int C>
__device__ __forceinline__ f2 unpack2(unsigned wa, unsigned wb, unsigned k4b) {
  unsigned pa, pb;      // (selector as an immediate, the 2^23 pattern in ONE register: ptxas otherwise spends a register per selector)
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(pa) : "r"(wa), "r"(k4b), "n"(0x7540 | C));
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(pb) : "r"(wb), "r"(k4b), "n"(0x7540 | C));
  return vadd(mk2(__uint_as_float(pa), __uint_as_float(pb)), dup2(-8388608.f));
}

struct RowFetch {
  unsigned a_nw, a_ne, a_sw, a_se, b_nw, b_ne, b_sw, b_se;   // corner words of source 0 / source 1
  f2 tx, ty;         // fractional sampling position
  unsigned clip;     // bit 0 / 1: u inside the image (source 0 / 1), bit 2 / 3: v inside  (GridSampler.h clip_coordinates_set_grad)
  float d;           // depth of the pixel
};

__device__ __forceinline__ void fetch_packed(const uint32_t* s0, const uint32_t* s1, int W, const ProjT<f2>& pr, float wm1, float hm1,
                                             float d, RowFetch& rf) {
  const FloorIF x0 = floor_if(pr.ix.x), x1 = floor_if(pr.ix.y), y0 = floor_if(pr.iy.x), y1 = floor_if(pr.iy.y);
  const unsigned o0 = (unsigned)(y0.i * W + x0.i), o1 = (unsigned)(y1.i * W + x1.i), uW = (unsigned)W;
  const uint32_t* p0 = s0 + o0;
  const uint32_t* p1 = s1 + o1;
  const uint32_t* q0 = s0 + (o0 + uW);
  const uint32_t* q1 = s1 + (o1 + uW);
  rf.a_nw = __ldg(p0), rf.a_ne = __ldg(p0 + 1), rf.a_sw = __ldg(q0), rf.a_se = __ldg(q0 + 1);
  rf.b_nw = __ldg(p1), rf.b_ne = __ldg(p1 + 1), rf.b_sw = __ldg(q1), rf.b_se = __ldg(q1 + 1);
  rf.tx = vsub(pr.ix, mk2(x0.f, x1.f));
  rf.ty = vsub(pr.iy, mk2(y0.f, y1.f));
  rf.clip = ((pr.u.x > 0.f && pr.u.x < wm1) ? 1u : 0u) | ((pr.u.y > 0.f && pr.u.y < wm1) ? 2u : 0u) |
            ((pr.v.x > 0.f && pr.v.x < hm1) ? 4u : 0u) | ((pr.v.y > 0.f && pr.v.y < hm1) ? 8u : 0u);
  rf.d = d;
}

__device__ __forceinline__ void blend_packed(const RowFetch& rf, unsigned k4b, f2 (&val)[3], f2 (&ddx)[3], f2 (&ddy)[3]) {
  constexpr float kInv255 = 1.f / 255.f;
  const f2 tx = rf.tx, ty = rf.ty;
  const f2 ex = vsub(dup2(1.f), tx);
  const f2 tys = vmul(ty, dup2(kInv255));                      // rows carry the 1/255
  const f2 eys = vfma(ty, dup2(-kInv255), dup2(kInv255));      // (1 - ty) / 255
  const f2 wnw = vmul(eys, ex), wne = vmul(eys, tx), wsw = vmul(tys, ex), wse = vmul(tys, tx);
  const f2 mx = mk2((rf.clip & 1u) ? 1.f : 0.f, (rf.clip & 2u) ? 1.f : 0.f);
  const f2 my = mk2((rf.clip & 4u) ? kInv255 : 0.f, (rf.clip & 8u) ? kInv255 : 0.f);
  const f2 eym = vmul(eys, mx), tym = vmul(tys, mx), exm = vmul(ex, my), txm = vmul(tx, my);
#define PPEA_PACKED_CHANNEL(C)                                                                                     \
  {                                                                                                                \
    const f2 nw = unpack2<C>(rf.a_nw, rf.b_nw, k4b), ne = unpack2<C>(rf.a_ne, rf.b_ne, k4b),                        \
             sw = unpack2<C>(rf.a_sw, rf.b_sw, k4b), se = unpack2<C>(rf.a_se, rf.b_se, k4b);                        \
    val[C] = vfma(se, wse, vfma(sw, wsw, vfma(ne, wne, vmul(nw, wnw))));                                            \
    ddx[C] = vfma(vsub(se, sw), tym, vmul(vsub(ne, nw), eym));                                                     \
    ddy[C] = vfma(vsub(se, ne), txm, vmul(vsub(sw, nw), exm));                                                     \
  }
  PPEA_PACKED_CHANNEL(0)
  PPEA_PACKED_CHANNEL(1)
  PPEA_PACKED_CHANNEL(2)
#undef PPEA_PACKED_CHANNEL
}

// ---------------------------------------------------------------------------------------------------
// Preparation launch, once per step (everything here is independent of the scale, of disp and of T):
//   * packs both source frames to RGBA8 words and verifies value by value that the packing is exact
//     (fmt_flag is raised otherwise; the main launch then gathers from the planar fp32 frames);
//   * identity loss  min_f photo(src_f, tgt)  (trainer.py:1060-1069) of every pixel, with the same
//     streaming window sums as the main kernel (forward only).
// A warp task = (image, strip of 30 decided columns, segment of kPrepSegRows rows); the nine loads of a
// row are issued one row ahead.
// ---------------------------------------------------------------------------------------------------
// k = round(255 v) and whether v is EXACTLY fl(k / 255), ToTensor's value for the byte k.  fl(k / 255) is
// formed without a division: q = k * fl(1/255), then one Newton correction fma(fma(-q, 255, k), fl(1/255), q),
// which equals the IEEE quotient for every k in 0..255 (checked exhaustively, tests/test_host.py).
__device__ __forceinline__ unsigned byte_of(float v, bool& exact) {
  constexpr float r = 1.f / 255.f, kMagic = 12582912.f;      // 1.5 * 2^23: adding it rounds to an integer in the low mantissa bits
  const float t = fma_rn(v, 255.f, kMagic);                  // (no conversion instructions: they run on the quarter-rate pipe)
  const unsigned ki = __float_as_uint(t) - 0x4B400000u;      // k = round(255 v); out of 0..255 (or NaN) fails the test below
  const float k = t - kMagic;
  const float q = mul_rn(k, r);
  const float q2 = fma_rn(fma_rn(-q, 255.f, k), r, q);
  exact = exact && (q2 == v) && (ki <= 255u);
  return ki & 255u;
}
__device__ __forceinline__ unsigned pack_rgb(float r, float g, float b, bool& exact) {
  return byte_of(r, exact) | (byte_of(g, exact) << 8) | (byte_of(b, exact) << 16);
}

struct PrepRow {
  float y[3];
  f2 x[3];
};

#ifndef PPEA_PREP_CTAS
#define PPEA_PREP_CTAS 3
#endif
__global__ void __launch_bounds__(kSmoothThreads, PPEA_PREP_CTAS) vsl_prep_kernel(const __grid_constant__ VslArgs a, int strips, int segs, int seg_rows, int n_task_ctas) {
  __shared__ float red[3 * kSmoothThreads / 32];
  grid_launch_dependents();      // the main launch may take idle SMs early (it waits for our results where it needs them)
  if ((int)blockIdx.x >= n_task_ctas) {
    // Smoothness term of every scale (smooth.cuh: sums + un-normalised stencil field) as extra CTAs of this launch: like
    // the preparation tasks they depend on nothing and are latency-bound column walks, so the two kinds of CTA share the
    // SMs instead of queueing behind each other.
    smooth_fused_role(a, blockIdx.x - n_task_ctas, red);
    return;
  }
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) a.fmt_flag[1] = 0u;      // sticky "non-finite contribution" word of the fixed-point fields (fixed_add_s)
  int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= a.B * strips * segs) return;
  const int strip = t % strips;
  t /= strips;
  const int seg = t % segs;
  const int b = t / segs;
  const int H = a.H, W = a.W;
  const unsigned plane = (unsigned)(H * W);
  const bool automask = !(a.flags & PPEA_F_MULTI) && (a.flags & PPEA_F_AUTOMASK);
  const bool no_ssim = a.flags & PPEA_F_NO_SSIM;
  const float l1w = no_ssim ? (1.f / 3.f) : PPEA_W_L1, w_ssim = no_ssim ? 0.f : PPEA_W_SSIM;
  const int x0 = strip * kPrepStripW - 1, gx = x0 + lane;
  const int px = reflect_index(gx, W);
  const bool own_col = lane >= 1 && lane <= 30 && gx < W;      // (gx >= 0 for lane >= 1)
  const int y0 = seg * seg_rows, y1 = min(y0 + seg_rows, H);
  const float* tgt_b = a.tgt + (size_t)b * 3 * plane;
  const float* s0_b = a.src[0] + (size_t)b * 3 * plane;
  const float* s1_b = a.src[1] + (size_t)b * 3 * plane;
  uint32_t* pk0 = a.pk[0] + (size_t)b * plane;
  uint32_t* pk1 = a.pk[1] + (size_t)b * plane;
  float* ident_b = a.ident + (size_t)b * plane;

  f2 hx[2][9], xr[2][3];
  float yh[2][6], ycr[2][3];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
#pragma unroll
    for (int e = 0; e < 9; ++e) hx[p][e] = dup2(0.f);
#pragma unroll
    for (int e = 0; e < 3; ++e) xr[p][e] = dup2(0.f), ycr[p][e] = 0.f;
#pragma unroll
    for (int e = 0; e < 6; ++e) yh[p][e] = 0.f;
  }
  bool exact = true;

  auto load_row = [&](int gi, PrepRow& r) {
    const unsigned o = (unsigned)reflect_index(gi, H) * (unsigned)W + (unsigned)px;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      r.y[c] = __ldg(tgt_b + (c * plane + o));
      r.x[c] = mk2(__ldg(s0_b + (c * plane + o)), __ldg(s1_b + (c * plane + o)));
    }
  };
  const int g_lo = y0 - 1, g_hi = y1;      // one halo row above and below: the window sums of rows y0 .. y1-1
#ifndef PPEA_PREP_PREFETCH
#define PPEA_PREP_PREFETCH 0
#endif
  // rows further ahead are pulled into L2 (no registers held): the tasks are bound by DRAM latency, not by bytes
  auto prefetch_row = [&](int gi) {
    if (PPEA_PREP_PREFETCH > 0 && gi <= g_hi) {
      const unsigned o = (unsigned)reflect_index(gi, H) * (unsigned)W + (unsigned)px;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(tgt_b + (c * plane + o)));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(s0_b + (c * plane + o)));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(s1_b + (c * plane + o)));
      }
    }
  };
  float2* ys_b = a.ystat + (size_t)b * plane;
  const size_t ys_plane = (size_t)a.B * plane;
#ifndef PPEA_PREP_DEPTH
#define PPEA_PREP_DEPTH 1
#endif
  // rows in flight ahead of the one being processed: the launch is bound by DRAM latency x loads in flight, not by bytes
  PrepRow nxt, nxt2;
  load_row(g_lo, nxt);
#pragma unroll
  for (int k = 1; k <= PPEA_PREP_PREFETCH; ++k) prefetch_row(g_lo + k);
  if (PPEA_PREP_DEPTH == 2) load_row(g_lo + 1 <= g_hi ? g_lo + 1 : g_lo, nxt2);

  auto step = [&](auto par, const int gi) {
    constexpr int P = decltype(par)::value, Q = 1 - P;
    const PrepRow cur = nxt;
    if (PPEA_PREP_DEPTH == 2) {
      nxt = nxt2;
      load_row(gi + 2 <= g_hi ? gi + 2 : g_hi, nxt2);
    } else {
      load_row(gi + 1 <= g_hi ? gi + 1 : gi, nxt);     // next row in flight while this one is processed
    }
    prefetch_row(gi + 1 + PPEA_PREP_PREFETCH);
    if (gi >= y0 && gi < y1) {                         // (then the row is not a reflected one)
      const unsigned w0 = pack_rgb(cur.x[0].x, cur.x[1].x, cur.x[2].x, exact), w1 = pack_rgb(cur.x[0].y, cur.x[1].y, cur.x[2].y, exact);
      if (own_col) {
        pk0[(unsigned)gi * (unsigned)W + (unsigned)gx] = w0;
        pk1[(unsigned)gi * (unsigned)W + (unsigned)gx] = w1;
      }
    }
    f2 hxn[9];
    float yhn[6];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float yl = __shfl_up_sync(0xffffffffu, cur.y[c], 1), yr = __shfl_down_sync(0xffffffffu, cur.y[c], 1);
      row_sums_y<float>(yl, cur.y[c], yr, yhn[2 * c], yhn[2 * c + 1]);
      if (automask) {
        const f2 xl = shfl_up2(cur.x[c]), xrt = shfl_down2(cur.x[c]);
        row_sums_x<f2>(xl, cur.x[c], xrt, dup2(yl), dup2(cur.y[c]), dup2(yr), hxn[3 * c], hxn[3 * c + 1], hxn[3 * c + 2]);
      } else {
        hxn[3 * c] = hxn[3 * c + 1] = hxn[3 * c + 2] = dup2(0.f);
      }
    }
    const int qi = gi - 1;
    if (qi >= y0) {                                    // (qi < y1 by the loop bounds)
      f2 L = dup2(0.f), cs = dup2(0.f);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        // target window sums: the streaming kernel reads them at every scale instead of re-forming them
        const float Sy = add_rn(add_rn(yh[P][2 * c], yh[Q][2 * c]), yhn[2 * c]);
        const float Syy = add_rn(add_rn(yh[P][2 * c + 1], yh[Q][2 * c + 1]), yhn[2 * c + 1]);
        if (own_col) ys_b[c * ys_plane + ((unsigned)qi * (unsigned)W + (unsigned)gx)] = make_float2(Sy, Syy);
        if (automask) {
          const f2 Sx = vadd(vadd(hx[P][3 * c], hx[Q][3 * c]), hxn[3 * c]);
          const f2 Sxx = vadd(vadd(hx[P][3 * c + 1], hx[Q][3 * c + 1]), hxn[3 * c + 1]);
          const f2 Sxy = vadd(vadd(hx[P][3 * c + 2], hx[Q][3 * c + 2]), hxn[3 * c + 2]);
          photo_channel<false>(Sx, Sxx, Sxy, Sy, Syy, xr[Q][c], ycr[Q][c], w_ssim, l1w, L, cs, nullptr);
        }
      }
      if (automask && own_col) ident_b[(unsigned)qi * (unsigned)W + (unsigned)gx] = fminf(L.x, L.y);
    }
#pragma unroll
    for (int e = 0; e < 9; ++e) hx[P][e] = hxn[e];
#pragma unroll
    for (int e = 0; e < 6; ++e) yh[P][e] = yhn[e];
#pragma unroll
    for (int c = 0; c < 3; ++c) xr[P][c] = cur.x[c], ycr[P][c] = cur.y[c];
  };
#pragma unroll 1
  for (int gi = g_lo; gi <= g_hi; gi += 2) {
    step(std::integral_constant<int, 0>{}, gi);
    if (gi + 1 <= g_hi) step(std::integral_constant<int, 1>{}, gi + 1);
  }
  if (!__all_sync(0xffffffffu, exact) && lane == 0) *a.fmt_flag = 1u;
}

// Smoothness term of every scale as a launch of its own that runs in the SHADOW of the streaming kernel: nothing the
// streaming kernel reads or writes is touched here (inputs: disp_s, colour_s; outputs: the chunk sums and the stencil
// field, read by the finish launches only), so the launch is made programmatically dependent on the streaming kernel
// and never waits for it BEFORE its work -- its CTAs (128 threads, few registers) take the places the streaming warps
// leave as the single wave drains.  Every CTA waits for the streaming grid at its END, so that "this grid complete"
// implies "streaming grid complete" for the finish launch that follows.
__global__ void __launch_bounds__(kSmoothThreads) vsl_smooth_tail_kernel(const __grid_constant__ VslArgs a) {
  __shared__ float red[3 * kSmoothThreads / 32];
  grid_launch_dependents();
  smooth_fused_role(a, blockIdx.x, red);
  grid_dependency_wait();
}

// Rows per warp task of the preparation launch: with the smoothness roles in a launch of their own the tasks alone
// must fill the device, so the segment is the longest one whose tasks still cover every resident warp slot of ONE
// wave (one halo row above and below each segment is recomputed: rows / (rows + 2) of the work is useful); batches
// that fill several waves anyway keep kPrepSegRows.
static int prep_seg_rows(const VslArgs& a, int strips) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("PPEA_PREP_SEG_ROWS");
    forced = e ? atoi(e) : 0;
  }
  if (forced > 0) return forced;
#ifdef PPEA_SMOOTH_IN_PREP
  return kPrepSegRows;
#else
  static int sm_count[64] = {};
  int dev = 0, sms = 0;
  (void)cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && sm_count[dev] > 0) {
    sms = sm_count[dev];
  } else {
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148, (void)cudaGetLastError();
    if (dev >= 0 && dev < 64) sm_count[dev] = sms;
  }
  const long long slots = (long long)sms * PPEA_PREP_CTAS * (kSmoothThreads / 32);
  const long long cols = (long long)a.B * strips;
  if (cols * ceil_div(a.H, kPrepSegRows) >= 2 * slots) return kPrepSegRows;
  const int segs = (int)(slots / cols) > 0 ? (int)(slots / cols) : 1;
  const int rows = ceil_div(a.H, segs);
  return rows < 16 ? 16 : rows;
#endif
}

cudaError_t launch_vsl_prep(const VslArgs& a, cudaStream_t stream) {
  const int strips = ceil_div(a.W, kPrepStripW), seg_rows = prep_seg_rows(a, strips), segs = ceil_div(a.H, seg_rows);
  const int n_task_ctas = ceil_div(a.B * strips * segs, kSmoothThreads / 32);
#ifdef PPEA_SMOOTH_IN_PREP
  const int n_smooth = a.S * a.B * kSmoothChunks;
#else
  const int n_smooth = 0;
#endif
#ifdef PPEA_PREP_MAX_CARVEOUT
  const cudaError_t e = ensure_max_carveout(vsl_prep_kernel);      // (the ring: 32 KB per CTA, five or six CTAs per SM)
  if (e != cudaSuccess) return e;
#endif
  vsl_prep_kernel<<<n_task_ctas + n_smooth, kSmoothThreads, 0, stream>>>(a, strips, segs, seg_rows, n_task_ctas);
  return cudaGetLastError();
}

cudaError_t launch_vsl_smooth_tail(const VslArgs& a, cudaStream_t stream) {
#ifdef PPEA_SMOOTH_IN_PREP
  (void)a, (void)stream;
  return cudaSuccess;
#else
  // (same carve-out as the streaming kernel: an SM need not drain to change its configuration before it can take these CTAs)
  const cudaError_t e = ensure_max_carveout(vsl_smooth_tail_kernel);
  if (e != cudaSuccess) return e;
  return launch_pdl(vsl_smooth_tail_kernel, dim3(a.S * a.B * kSmoothChunks), dim3(kSmoothThreads), 0, stream, a);
#endif
}

// ---------------------------------------------------------------------------------------------------
// Deterministic mode: 64-bit fixed-point accumulation of the coarse-scale fields (see vsl_fused.cu).
// ---------------------------------------------------------------------------------------------------
constexpr float kFixScaleS = 1099511627776.f;          // 2^40
// A contribution that is not a finite number within +-8.3e6 cannot be represented in the fixed-point field: it is clamped
// and noted in `bad`; the warp raises the sticky word fmt_flag[1] (cleared by the preparation launch) once per piece, and the
// gradient finish turns the coarse-scale gradients of the step into NaN -- what the float-atomic fields would carry in the
// cells such a contribution reaches.
__device__ __forceinline__ void fixed_add_s(float* field, unsigned idx, float v, bool& bad) {
  const float c = fminf(fmaxf(v, -8.3e6f), 8.3e6f);      // (NaN -> -8.3e6: c != v)
  bad = bad || (c != v);
  const long long q = __float2ll_rn(c * kFixScaleS);
  atomicAdd(reinterpret_cast<unsigned long long*>(field) + idx, (unsigned long long)q);
}
template <bool DET>
__device__ __forceinline__ void field_add(float* field, unsigned idx, float v, bool& bad) {
  if (v == 0.f) return;
  if (DET)
    fixed_add_s(field, idx, v, bad);
  else
    atomicAdd(field + idx, v);
}

// Adjoint of the bilinear upsample of one column, aggregated while the walk stays on the same pair of coarse
// rows (cur, cur + 1): a full-resolution pixel adds into four register accumulators, and a coarse row is
// flushed (two atomics) only when the walk leaves it.
template <bool DET>
struct UpAgg {
  float a00, a01, a10, a11;
  int cur;
  __device__ __forceinline__ void init(int i0) {
    a00 = a01 = a10 = a11 = 0.f;
    cur = i0;
  }
  __device__ __forceinline__ void flush_row0(float* field, int ws, const UpCoef& cx, bool own, bool& bad) {
    if (own) {
      field_add<DET>(field, (unsigned)(cur * ws + cx.i0), a00, bad);
      field_add<DET>(field, (unsigned)(cur * ws + cx.i1), a01, bad);
    }
  }
  // cy is warp-uniform (the row), cx this lane's column
  __device__ __forceinline__ void add(float* field, int ws, float g, const UpCoef& cy, const UpCoef& cx, bool own, bool& bad) {
    if (cy.i0 != cur) {          // (warp-uniform; the coarse row index grows by at most one per full-resolution row)
      flush_row0(field, ws, cx, own, bad);
      a00 = a10, a01 = a11;
      a10 = a11 = 0.f;
      cur = cy.i0;
    }
    const float t0 = g * cy.l0, t1 = g * cy.l1;
    a00 = fmaf(t0, cx.l0, a00);
    a01 = fmaf(t0, cx.l1, a01);
    if (cy.i1 != cy.i0) {
      a10 = fmaf(t1, cx.l0, a10);
      a11 = fmaf(t1, cx.l1, a11);
    } else {                     // bottom border: both taps are the last coarse row
      a00 = fmaf(t1, cx.l0, a00);
      a01 = fmaf(t1, cx.l1, a01);
    }
  }
  __device__ __forceinline__ void finish(float* field, int ws, int hs, const UpCoef& cx, bool own, bool& bad) {
    flush_row0(field, ws, cx, own, bad);
    if (own && cur + 1 < hs) {
      field_add<DET>(field, (unsigned)((cur + 1) * ws + cx.i0), a10, bad);
      field_add<DET>(field, (unsigned)((cur + 1) * ws + cx.i1), a11, bad);
    }
  }
};

// an image base pointer the optimiser cannot see through: every address formed from it is ONE IMAD.WIDE.U32 of a
// 32-bit element offset instead of a 64-bit add chain per access
template <class T>
__device__ __forceinline__ const T* opaque_base(const T* p) {
  unsigned long long v = reinterpret_cast<unsigned long long>(p);
  asm volatile("" : "+l"(v));
  return reinterpret_cast<const T*>(v);
}

#ifdef PPEA_STREAM_CLOCKS
// Diagnostic build only (scripts/stream_clocks.py): start / end time, SM and row range of every chunk of the last launch.
__device__ unsigned long long g_stream_clk[8192][2];
__device__ unsigned g_stream_sm[8192];
__device__ int g_stream_rows[8192][2];
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned sm_id() {
  unsigned v;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
  return v;
}
#endif

// ---------------------------------------------------------------------------------------------------
// Main launch.  grid = one CTA (one warp) per task.
// Iteration gi:  issue the loads of row gi+1 (gather corners, target, noise / identity loss of the row
// decided next) -> blend row gi from the words fetched one iteration ago -> row sums -> decide row
// gi-1 -> fold + chain row gi-2.
// ---------------------------------------------------------------------------------------------------
template <bool POSE, bool MULTI, bool DET>
__global__ void __launch_bounds__(kStreamThreads, PPEA_STREAM_CTAS) vsl_stream_kernel(const __grid_constant__ VslArgs a, int n_task_ctas) {
  // per warp: ring of three rows of bilinear derivatives (d warped_c / d(u,v), 12 floats per lane and row) and the
  // folded projection of the image (vsl_math.cuh Geom, lanes = sources)
  __shared__ float4 s_ring[kStreamWarps][2][kRingK][32];
  __shared__ float4 s_dd[kStreamWarps][3][3][32];
  __shared__ f2 s_G[kStreamWarps][12];

  grid_launch_dependents();      // the dependent is the small finish kernel: let it take its place early
  const int lane = threadIdx.x & 31, wid = kStreamWarps == 1 ? 0 : threadIdx.x >> 5;
  // Work = the rows of every (image, strip, scale) column laid end to end; warp j walks rows [j L, (j+1) L) of that line
  // (L = a.seg_rows), i.e. the tail of one column and the head of the next where its chunk crosses a column end: all warps
  // of the (single) wave carry the same number of rows.  A piece of a column is tile slot (image, piece index, strip).
  const int strips = a.tiles_x, pmax = a.tiles_y;
  const int n_total = a.B * strips * a.S * a.H;
  const int chunk = blockIdx.x * kStreamWarps + wid;
  const int g_begin = chunk * a.seg_rows, g_end = min(g_begin + a.seg_rows, n_total);
  if (g_begin >= n_total) return;          // (whole warp; nothing below synchronises across warps)
  grid_dependency_wait();        // packed sources, identity loss, window sums and the format flag come from the preparation launch
  const bool packed = (*reinterpret_cast<const volatile unsigned*>(a.fmt_flag) == 0u);
#ifdef PPEA_STREAM_CLOCKS
  if (lane == 0 && chunk < 8192)
    g_stream_clk[chunk][0] = global_ns(), g_stream_sm[chunk] = sm_id(), g_stream_rows[chunk][0] = g_begin, g_stream_rows[chunk][1] = g_end;
#endif
#pragma unroll 1
  for (int g_cur = g_begin; g_cur < g_end;) {
  const int col = g_cur / a.H;
  const int y0 = g_cur - col * a.H, y1 = min(a.H, y0 + (g_end - g_cur));
  g_cur += y1 - y0;
  const int s = col % a.S;
  const int strip = (col / a.S) % strips;
  const int b = col / (a.S * strips);
  const int piece = chunk - (col * a.H) / a.seg_rows;      // chunks that started inside this column before this one
  const int tile_id = (b * pmax + piece) * strips + strip;

  const int H = a.H, W = a.W;
  const unsigned plane = (unsigned)(H * W), uW = (unsigned)W;
  const ScaleArgs& sc = a.sc[s];
  const int hs = sc.hs, ws = sc.ws;
  const bool same_res = (hs == H && ws == W);
  const bool automask = !MULTI && (a.flags & PPEA_F_AUTOMASK);
  const bool motion = MULTI && (a.flags & PPEA_F_MOTION_MASK);
  const float one_minus_aug = (MULTI && (a.flags & PPEA_F_MATCH_AUG)) ? 1.f - a.aug_mask[b] : 1.f;
  const bool no_ssim = a.flags & PPEA_F_NO_SSIM;
  const bool selec = a.flags & PPEA_F_SELEC_REPROJ;
  const float l1w = no_ssim ? (1.f / 3.f) : PPEA_W_L1, w_ssim = no_ssim ? 0.f : PPEA_W_SSIM;
  const bool use_noise = automask && sc.noise != nullptr;

  f2* const G = s_G[wid];
  __syncwarp();                  // (the previous piece of this warp is done with G)
  if (lane < 24) {
    const int f = lane / 12, e = lane % 12;
    const float v = geom_entry(a.K + b * 16, a.T[f] + b * 16, a.inv_K + b * 16, e);
    (f ? G[e].y : G[e].x) = v;
  }
  __syncwarp();

  const int x0 = strip * kStripW - 2, gx = x0 + lane;
  const bool col_in = gx >= 0 && gx < W;
  const bool q_lane = col_in && lane >= 1 && lane <= 30;
  const bool own_col = lane >= 2 && lane <= 29 && gx < W;
  ColCtx cc = make_col(G, gx, W);
  if (!same_res) cc.cx = up_coef(cc.px, ws, sc.up_sx);
  const float wmax = coord_max(W), hmax = coord_max(H);
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  const unsigned upx = (unsigned)cc.px, ugx = (unsigned)(col_in ? gx : 0);
  const unsigned img0 = (unsigned)b * plane;          // first pixel of image b in the (B,1,H,W) maps (B*H*W < 2^31 is checked by the API)
  // this lane's column inside the per-pixel maps, as ONE register the optimiser cannot re-derive (it otherwise rebuilds the
  // clamped column from gx in five places of the row loop: 1 003 -> 987 instructions per row, step 0.3717 -> 0.3680 ms)
#ifndef PPEA_STREAM_NO_OCOL
  unsigned o_col = img0 + ugx;
  asm volatile("" : "+r"(o_col));
#else
  const unsigned o_col = img0 + ugx;
#endif

  // Per-piece base pointers the optimiser otherwise re-derives from the argument block inside the row loop can be pinned in
  // registers (opaque_base).  Whether that pays is decided by what it does to the register allocation of the loop, and was
  // measured per instantiation (profiles/README.md r2r): the target base helps the multi path (0.3850 -> 0.3778 ms) and costs
  // the mono path 0.3 %; the disparity base and the window-sum base lose on both.
#if defined(PPEA_STREAM_PIN) || defined(PPEA_STREAM_PIN_TGT)
  const float* tgt_b = opaque_base(a.tgt + (size_t)b * 3 * plane);
#elif defined(PPEA_STREAM_NO_PIN_TGT)
  const float* tgt_b = a.tgt + (size_t)b * 3 * plane;
#else
  const float* tgt_b = MULTI ? opaque_base(a.tgt + (size_t)b * 3 * plane) : a.tgt + (size_t)b * 3 * plane;
#endif
#if defined(PPEA_STREAM_PIN) || defined(PPEA_STREAM_PIN_DISP)
  const float* disp_b = opaque_base(sc.disp + (size_t)b * hs * ws);
#else
  const float* disp_b = sc.disp + (size_t)b * hs * ws;
#endif

  const uint32_t* pk0 = opaque_base(a.pk[0] + (size_t)b * plane);
  const uint32_t* pk1 = opaque_base(a.pk[1] + (size_t)b * plane);
  const SrcPlanes sp = make_planes(a.src[0] + (size_t)b * 3 * plane, a.src[1] + (size_t)b * 3 * plane, plane);
  // 2^23 as a bit pattern, in a register ptxas cannot fold (it would otherwise make it the immediate of every PRMT and
  // spend a register move per byte selector): B > 0, so the shifted term is zero
  const unsigned k4b = 0x4B000000u | ((unsigned)a.B >> 31);

  // ---- per-lane rings in shared memory.  Everything a row hands to the two iterations after it goes through them;
  // a lane only ever touches its own column, so program order is the only synchronisation.  Slot = parity of the
  // iteration that wrote it: iteration gi reads slot Q = (gi-1)&1 (previous iteration) and slot P = gi&1 (two
  // iterations ago), then overwrites slot P group by group once the group's old content has been read.
  //   k 0..3  hx[0..7]          horizontal 3-tap sums (x, x^2, x y per channel, both sources) of row gi
  //   k 4     hx[8], ycr[2], depth                                                              of row gi
  //   k 5..6  xr[0..2], ycr[0..1]   warped samples / target values                              of row gi
  //   k 7..10 hc[0..7]          horizontal (multiplicity-weighted) sums of the masked adjoint coefficients of row gi-1
  //   k 11    hc[8], ind        mask(q) * [sel(q) == lane]                                       of row gi-1
  float4* const ring = &s_ring[wid][0][0][lane];
  float4* const ddr = &s_dd[wid][0][0][lane];
#pragma unroll
  for (int i = 0; i < 2 * kRingK; ++i) ring[i * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 9; ++i) ddr[i * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
  const f2 mL = dup2((gx == 1) ? 2.f : 1.f), mR = dup2((gx == W - 2) ? 2.f : 1.f);   // reflection multiplicities (layers.py:238)

  float s_rm = 0.f, s_m = 0.f, s_c = 0.f;
  f2 Sw[POSE ? 3 : 1], Swy[POSE ? 3 : 1], Sg[POSE ? 3 : 1];
#pragma unroll
  for (int e = 0; e < (POSE ? 3 : 1); ++e) Sw[e] = Swy[e] = Sg[e] = dup2(0.f);

  const size_t img_off = (size_t)b * hs * ws;
  float* gr_b = sc.grad_raw + (DET && !same_res ? 2 * img_off : img_off);          // (fixed-point fields: 8 bytes per pixel)
  float* gc_b = MULTI ? sc.grad_raw2 + (DET && !same_res ? 2 * img_off : img_off) : nullptr;
  UpAgg<DET> agg, agg2;
  bool bad_add = false;          // fixed-point fields: some contribution of this piece was not representable (fixed_add_s)
  {
    const int i0 = same_res ? 0 : up_coef(y0, hs, sc.up_sy).i0;
    agg.init(i0);
    agg2.init(i0);
  }

  // the (upsampled) disparity of row gi at this lane's column: loads only, issued two rows ahead of the blend
  // (coarse scales: the four taps of the bilinear upsample are fetched raw and blended where the value is used, so that
  // the blend does not wait for them inside the iteration that issued the loads)
  struct DispTaps {
    float v00, v01, v10, v11;
  };
  auto load_disp = [&](int gi) -> DispTaps {
    const int py = reflect_index(gi, H);
    DispTaps t;
    if (same_res) {
      t.v00 = __ldg(disp_b + ((unsigned)py * uW + upx));
      t.v01 = t.v10 = t.v11 = 0.f;
    } else {
      const UpCoef cy = up_coef(py, hs, sc.up_sy);
      const unsigned r0 = (unsigned)cy.i0 * (unsigned)ws, r1 = (unsigned)cy.i1 * (unsigned)ws;
      t.v00 = __ldg(disp_b + (r0 + (unsigned)cc.cx.i0)), t.v01 = __ldg(disp_b + (r0 + (unsigned)cc.cx.i1));
      t.v10 = __ldg(disp_b + (r1 + (unsigned)cc.cx.i0)), t.v11 = __ldg(disp_b + (r1 + (unsigned)cc.cx.i1));
    }
    return t;
  };
  // same roundings as up_sample (vsl_math.cuh): every kernel gets the same bits
  auto disp_value = [&](int gi, const DispTaps& t) -> float {
    if (same_res) return t.v00;
    const UpCoef cy = up_coef(reflect_index(gi, H), hs, sc.up_sy);
    const float top = fma_rn(cc.cx.l1, t.v01, mul_rn(cc.cx.l0, t.v00));
    const float bot = fma_rn(cc.cx.l1, t.v11, mul_rn(cc.cx.l0, t.v10));
    return fma_rn(cy.l1, bot, mul_rn(cy.l0, top));
  };
  // projection of row gi and the loads of its eight packed corners (consumed by the next iteration)
  auto fetch_row = [&](int gi, float dup, RowFetch& r) {
    const int py = reflect_index(gi, H);
    const float d = depth_from_disp_fast(dup, a.disp_lo, a.disp_range);
    f2 A[3];
    const ProjT<f2> pr = project_cell(G, cc, py, d, a.eps, wmax, hmax, A);
    fetch_packed(pk0, pk1, W, pr, wm1, hm1, d, r);
  };
  auto load_tgt = [&](int gi, float (&y)[3]) {
    const unsigned o = (unsigned)reflect_index(gi, H) * uW + upx;
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = __ldg(tgt_b + (c * plane + o));
  };
  // what the NEXT iteration needs to decide row qi: the target's window sums (preparation launch; rows outside the
  // image are decided with mask 0, they only need finite numbers: the reflected row), identity loss (+ tie-break
  // noise) / consistency mask
  const float2* ys_b = a.ystat + (img0 + upx);
  const unsigned ys_plane = (unsigned)a.B * plane;
  auto load_decide = [&](int qi, float2 (&ys)[3], float& idv, float& nzv, float& cmv) {
    const unsigned orow = (unsigned)reflect_index(qi, H) * uW;
#pragma unroll
    for (int c = 0; c < 3; ++c) ys[c] = __ldg(ys_b + (c * ys_plane + orow));
    idv = nzv = 0.f;
    cmv = 1.f;
    if (q_lane && qi >= 0 && qi < H) {
      const unsigned o = o_col + (unsigned)qi * uW;
      if (automask) idv = __ldg(a.ident + o);
      if (use_noise) nzv = __ldg(sc.noise + o);
      if (motion) cmv = __ldg(a.cons_mask + o);
    }
  };

  // The row loop, specialised on the gather format: only the executed copy occupies the instruction cache.
  auto run_rows = [&](auto packed_c) {
  constexpr bool packed = decltype(packed_c)::value;
  DispTaps dup_pf = load_disp(y0 - 2);
  RowFetch rf = RowFetch{};       // gather of the row blended by the coming iteration
  if (packed) {                   // (the planar gathers are not prefetched: they keep the disparity of the row itself)
    fetch_row(y0 - 2, disp_value(y0 - 2, dup_pf), rf);
    dup_pf = load_disp(y0 - 1);
  }
  float y_pf[3];
  load_tgt(y0 - 2, y_pf);
  float2 ys_pf[3];
  float id_pf, nz_pf, cm_pf;
  load_decide(y0 - 3, ys_pf, id_pf, nz_pf, cm_pf);   // (the first two iterations decide nothing that is kept)
  int dslot = 0;                                     // ring slot the derivatives of the current row go to

  // One loop body for every row (no unrolling: the body must stay inside the instruction cache).
#pragma unroll 1
  for (int gi = y0 - 2; gi <= y1 + 1; ++gi) {
    float4* const sP = ring + (gi & 1) * (kRingK * 32);
    float4* const sQ = ring + ((gi & 1) ^ 1) * (kRingK * 32);
    float yv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) yv[c] = y_pf[c];

    // ---- blend row gi from the words fetched one iteration ago (bilinear samples of both sources, derivatives to
    // their ring), THEN issue the gather of row gi + 1 into the same registers: a whole row of arithmetic covers its
    // L2 latency and no fetch state is ever copied
    f2 xn[3];
    float d;
    {
      f2 dx[3], dy[3];
      if (packed) {
        blend_packed(rf, k4b, xn, dx, dy);
        d = rf.d;
        fetch_row(gi + 1, disp_value(gi + 1, dup_pf), rf);
        dup_pf = load_disp(gi + 2);
      } else {
        const float dup_cur = disp_value(gi, dup_pf);
        dup_pf = load_disp(gi + 1);
        const int py = reflect_index(gi, H);
        d = depth_from_disp_fast(dup_cur, a.disp_lo, a.disp_range);
        f2 A[3];
        const ProjT<f2> pr = project_cell(G, cc, py, d, a.eps, wmax, hmax, A);
        sample_sources<true, false>(sp, W, pr, wm1, hm1, xn, dx, dy);
      }
      float4* dst = ddr + dslot * (3 * 32);
      dst[0] = make_float4(dx[0].x, dx[0].y, dx[1].x, dx[1].y);
      dst[32] = make_float4(dx[2].x, dx[2].y, dy[0].x, dy[0].y);
      dst[64] = make_float4(dy[1].x, dy[1].y, dy[2].x, dy[2].y);
    }
    load_tgt(gi + 1, y_pf);
    if (gi >= y0 && gi < y1 && own_col) sc.depth[o_col + (unsigned)gi * uW] = d;   // trainer.py:893
    const int pi = gi - 2;
    float md = 0.f, cmf = 1.f;                        // fold row: mono depth / consistency mask (multi path)
    if (MULTI && own_col && pi >= y0) {
      const unsigned o = o_col + (unsigned)pi * uW;
      md = __ldg(sc.mono_depth + o);
      if (motion) cmf = __ldg(a.cons_mask + o);
    }

    // ---- horizontal 3-tap sums of this row
    f2 hxn[9];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float yl = __shfl_up_sync(0xffffffffu, yv[c], 1), yrt = __shfl_down_sync(0xffffffffu, yv[c], 1);
      const f2 xl = shfl_up2(xn[c]), xrt = shfl_down2(xn[c]);
      row_sums_x<f2>(xl, xn[c], xrt, dup2(yl), dup2(yv[c]), dup2(yrt), hxn[3 * c], hxn[3 * c + 1], hxn[3 * c + 2]);
    }

    // ---- window sums of row qi = gi - 1: rows gi-2 (slot P) + gi-1 (slot Q) from the ring, + this row
    f2 S[9];
    float ycP2, depP, ycQ[3];
    f2 xq[3];
    {
      const float4 p0 = sP[0], p1 = sP[32], p2 = sP[64], p3 = sP[96], p4 = sP[128];
      const float4 q0 = sQ[0], q1 = sQ[32], q2 = sQ[64], q3 = sQ[96], q4 = sQ[128];
      S[0] = vadd(vadd(mk2(p0.x, p0.y), mk2(q0.x, q0.y)), hxn[0]);
      S[1] = vadd(vadd(mk2(p0.z, p0.w), mk2(q0.z, q0.w)), hxn[1]);
      S[2] = vadd(vadd(mk2(p1.x, p1.y), mk2(q1.x, q1.y)), hxn[2]);
      S[3] = vadd(vadd(mk2(p1.z, p1.w), mk2(q1.z, q1.w)), hxn[3]);
      S[4] = vadd(vadd(mk2(p2.x, p2.y), mk2(q2.x, q2.y)), hxn[4]);
      S[5] = vadd(vadd(mk2(p2.z, p2.w), mk2(q2.z, q2.w)), hxn[5]);
      S[6] = vadd(vadd(mk2(p3.x, p3.y), mk2(q3.x, q3.y)), hxn[6]);
      S[7] = vadd(vadd(mk2(p3.z, p3.w), mk2(q3.z, q3.w)), hxn[7]);
      S[8] = vadd(vadd(mk2(p4.x, p4.y), mk2(q4.x, q4.y)), hxn[8]);
      ycP2 = p4.z, depP = p4.w, ycQ[2] = q4.z;
      // the sums of this row take the place of row gi-2
      sP[0] = make_float4(hxn[0].x, hxn[0].y, hxn[1].x, hxn[1].y);
      sP[32] = make_float4(hxn[2].x, hxn[2].y, hxn[3].x, hxn[3].y);
      sP[64] = make_float4(hxn[4].x, hxn[4].y, hxn[5].x, hxn[5].y);
      sP[96] = make_float4(hxn[6].x, hxn[6].y, hxn[7].x, hxn[7].y);
      sP[128] = make_float4(hxn[8].x, hxn[8].y, yv[2], d);
      const float4 q5 = sQ[160], q6 = sQ[192];
      xq[0] = mk2(q5.x, q5.y), xq[1] = mk2(q5.z, q5.w), xq[2] = mk2(q6.x, q6.y);
      ycQ[0] = q6.z, ycQ[1] = q6.w;
    }

    // ---- decide row qi = gi - 1: loss of both sources, selection, mask, masked adjoint coefficients.
    // (Runs in the two warm-up iterations too, on rows nobody owns: their products are never consumed.)
    const int qi = gi - 1;
    f2 hcn[9];
    f2 indn = dup2(0.f);
    {
      f2 L = dup2(0.f), cs = dup2(0.f);
      f2 co[9];
#pragma unroll
      for (int c = 0; c < 3; ++c)
        photo_channel<true>(S[3 * c], S[3 * c + 1], S[3 * c + 2], ys_pf[c].x, ys_pf[c].y, xq[c], ycQ[c], w_ssim, l1w, L, cs, &co[3 * c]);
      const Select sl = select_source(L.x, L.y, cs.x, cs.y, selec);
      bool on = true;
      float mq = 1.f;
      if (MULTI) {
        mq = cm_pf * one_minus_aug;
      } else if (automask) {
        const float idl = use_noise ? add_rn(id_pf, mul_rn(nz_pf, 0.00001f)) : id_pf;   // trainer.py:1086-1087
        on = sl.r <= idl;                                                               // argmin([r, id]) == 0
        mq = on ? 1.f : 0.f;
      }
      const bool q_ok = q_lane && qi >= 0 && qi < H;
      if (!q_ok) mq = 0.f;
      indn = mk2(sl.src == 0 ? mq : 0.f, sl.src == 1 ? mq : 0.f);
      if (own_col && qi >= y0 && qi < y1) {      // owner of q: forward products
        const unsigned o = o_col + (unsigned)qi * uW;
        if (sc.loss_px) sc.loss_px[o] = sl.r;
        sc.sel[o] = (uint8_t)((unsigned)sl.src | (on ? PPEA_SEL_AUTOMASK : 0u));
        s_rm = fmaf(sl.r, mq, s_rm);
        s_m += mq;
      }
      // (everything the decision read is dead: fetch what the next iteration decides with)
      load_decide(gi, ys_pf, id_pf, nz_pf, cm_pf);
      // horizontal box sums of the masked coefficients, with the reflection multiplicities of the columns
#pragma unroll
      for (int e = 0; e < 9; ++e) {
        const f2 cm = vmul(co[e], indn);
        hcn[e] = vfma(mL, shfl_up2(cm), vfma(mR, shfl_down2(cm), cm));
      }
    }

    // ---- fold + chain row pi = gi - 2
    if (gi >= y0 + 2 && pi < y1) {
      const int rslot = dslot == 2 ? 0 : dslot + 1;        // (dslot + 1) % 3 == (dslot - 2) % 3
      const float4* src = ddr + rslot * (3 * 32);
      const float4 r0 = src[0], r1 = src[32], r2 = src[64];
      const f2 ddx[3] = {mk2(r0.x, r0.y), mk2(r0.z, r0.w), mk2(r1.x, r1.y)};
      const f2 ddy[3] = {mk2(r1.z, r1.w), mk2(r2.x, r2.y), mk2(r2.z, r2.w)};
      const float4 p5 = sP[160], p6 = sP[192];
      const f2 xp[3] = {mk2(p5.x, p5.y), mk2(p5.z, p5.w), mk2(p6.x, p6.y)};
      const float ycP[3] = {p6.z, p6.w, ycP2};
      const float4 a0 = sP[224], a1 = sP[256], a2 = sP[288], a3 = sP[320], a4 = sP[352];      // coefficient row gi-3
      const float4 b0 = sQ[224], b1 = sQ[256], b2 = sQ[288], b3 = sQ[320], b4 = sQ[352];      // coefficient row gi-2 (+ its indicator)
      const f2 mU = dup2((pi == 1) ? 2.f : 1.f), mD = dup2((pi == H - 2) ? 2.f : 1.f);
      // 3x3 box sums of the coefficients around row pi:  mU * row(pi-1) + row(pi) + mD * row(pi+1)
      f2 V[9];
      V[0] = vfma(mD, hcn[0], vfma(mU, mk2(a0.x, a0.y), mk2(b0.x, b0.y)));
      V[1] = vfma(mD, hcn[1], vfma(mU, mk2(a0.z, a0.w), mk2(b0.z, b0.w)));
      V[2] = vfma(mD, hcn[2], vfma(mU, mk2(a1.x, a1.y), mk2(b1.x, b1.y)));
      V[3] = vfma(mD, hcn[3], vfma(mU, mk2(a1.z, a1.w), mk2(b1.z, b1.w)));
      V[4] = vfma(mD, hcn[4], vfma(mU, mk2(a2.x, a2.y), mk2(b2.x, b2.y)));
      V[5] = vfma(mD, hcn[5], vfma(mU, mk2(a2.z, a2.w), mk2(b2.z, b2.w)));
      V[6] = vfma(mD, hcn[6], vfma(mU, mk2(a3.x, a3.y), mk2(b3.x, b3.y)));
      V[7] = vfma(mD, hcn[7], vfma(mU, mk2(a3.z, a3.w), mk2(b3.z, b3.w)));
      V[8] = vfma(mD, hcn[8], vfma(mU, mk2(a4.x, a4.y), mk2(b4.x, b4.y)));
      const f2 wl = vmul(mk2(b4.z, b4.w), dup2(l1w));        // indicator of row pi (decided by the previous iteration)
      f2 gu = dup2(0.f), gv = dup2(0.f);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const f2 xv = xp[c], yv2 = dup2(ycP[c]);
        const f2 dd = vsub(yv2, xv);
        // L1 term:  -ind * l1w * sign(y - x);  SSIM term from the 3x3 box sums of the coefficients (zero with no_ssim)
        const f2 Gc = vadd(mk2(neg_signed(wl.x, dd.x), neg_signed(wl.y, dd.y)), vfma(V[3 * c + 1], xv, vfma(V[3 * c + 2], yv2, V[3 * c])));   // d L / d warped_c(p)
        gu = vfma(Gc, ddx[c], gu);
        gv = vfma(Gc, ddy[c], gv);
      }
      float g_dup = 0.f, c_dup = 0.f;
      if (own_col) {       // (pi in [y0, y1) by the loop bounds)
        f2 A[3];
        const float depk = depP;
        const ProjT<f2> pr = project_cell(G, cc, pi, depk, a.eps, wmax, hmax, A);
        const f2 gc0 = vmul(gu, pr.rz), gc1 = vmul(gv, pr.rz);
        const f2 gc2 = vneg(vmul(vfma(gu, pr.u, vmul(gv, pr.v)), pr.rz));
        const f2 gd2 = vfma(gc2, A[2], vfma(gc1, A[1], vmul(gc0, A[0])));
        const float g = gd2.x + gd2.y;
        if (POSE) {
          const f2 dk = dup2(depk), fy = dup2(int_to_float(pi));
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
        const float ddd = ddepth_ddisp(depk, a.disp_range);
        g_dup = g * ddd;
        if (MULTI) {
          // consistency term mean(|depth - mono_depth| * (1 - mask))  (trainer.py:1128-1132): its un-normalised
          // gradient goes to a field of its own (its upstream weight differs from the photometric one)
          const float om = 1.f - cmf * one_minus_aug;
          const float dm = depk - md;
          s_c = fmaf(fabsf(dm), om, s_c);
          c_dup = sign_of(dm) * om * ddd;
        }
        if (same_res) {
          const unsigned o0 = (unsigned)pi * uW + ugx;
          gr_b[o0] = g_dup;     // sole owner: plain store, no pre-zero needed
          if (MULTI) gc_b[o0] = c_dup;
        }
      }
      if (!same_res) {          // (warp-uniform: the aggregation decides its flushes on the row)
        const UpCoef cy = up_coef(pi, hs, sc.up_sy);
        agg.add(gr_b, ws, g_dup, cy, cc.cx, own_col, bad_add);
        if (MULTI) agg2.add(gc_b, ws, c_dup, cy, cc.cx, own_col, bad_add);
      }
    }

    // ---- the rest of this iteration's products take the place of what slot P held
    sP[160] = make_float4(xn[0].x, xn[0].y, xn[1].x, xn[1].y);
    sP[192] = make_float4(xn[2].x, xn[2].y, yv[0], yv[1]);
    sP[224] = make_float4(hcn[0].x, hcn[0].y, hcn[1].x, hcn[1].y);
    sP[256] = make_float4(hcn[2].x, hcn[2].y, hcn[3].x, hcn[3].y);
    sP[288] = make_float4(hcn[4].x, hcn[4].y, hcn[5].x, hcn[5].y);
    sP[320] = make_float4(hcn[6].x, hcn[6].y, hcn[7].x, hcn[7].y);
    sP[352] = make_float4(hcn[8].x, hcn[8].y, indn.x, indn.y);
    dslot = dslot == 2 ? 0 : dslot + 1;
  }
  if (!same_res) {
    agg.finish(gr_b, ws, hs, cc.cx, own_col, bad_add);
    if (MULTI) agg2.finish(gc_b, ws, hs, cc.cx, own_col, bad_add);
  }
  };
  if (packed)
    run_rows(std::true_type{});
  else
    run_rows(std::false_type{});

  if (DET && __any_sync(0xffffffffu, bad_add) && lane == 0) atomicOr(a.fmt_flag + 1, 1u);
  // ---- task sums: masked loss sums, consistency sum, pose partials (fixed order)
  s_rm = warp_sum(s_rm);
  s_m = warp_sum(s_m);
  if (MULTI) s_c = warp_sum(s_c);
  if (lane == 0)
    *reinterpret_cast<float4*>(a.partials + ((size_t)tile_id * a.S + s) * 4) = make_float4(s_rm, s_m, s_c, 0.f);
  if (POSE) {
    // per source f and row r of dL/dP: (sum gc_r*depth*x, sum gc_r*depth*y, sum gc_r*depth, sum gc_r)
    const float fx = int_to_float(cc.px);
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
    const float tsum = warp_sum24(v, lane);
    const int e = warp_sum24_index(lane);
    if (e < 24) a.pose_partials[((size_t)tile_id * a.S + s) * 24 + e] = tsum;
  }
  // the last piece of a column clears the column's unused tile slots (the finish launches sum every slot)
  if (y1 == a.H) {
    for (int k = piece + 1; k < pmax; ++k) {
      const size_t slot = (size_t)((b * pmax + k) * strips + strip) * a.S + s;
      if (lane == 0) *reinterpret_cast<float4*>(a.partials + slot * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      if (POSE && lane < 24) a.pose_partials[slot * 24 + lane] = 0.f;
    }
  }
  }      // pieces of this warp's chunk
#ifdef PPEA_STREAM_CLOCKS
  if (lane == 0 && chunk < 8192) g_stream_clk[chunk][1] = global_ns();
#endif
}

template <bool POSE, bool MULTI, bool DET>
static cudaError_t launch_vsl_stream_as(const VslArgs& a, cudaStream_t stream) {
  const int tasks = ceil_div(a.B * a.tiles_x * a.S * a.H, a.seg_rows);      // chunks of seg_rows rows of the column line
  const int n_task_ctas = ceil_div(tasks, kStreamWarps);
  const cudaError_t e = ensure_max_carveout(vsl_stream_kernel<POSE, MULTI, DET>);
  if (e != cudaSuccess) return e;
  return launch_pdl(vsl_stream_kernel<POSE, MULTI, DET>, dim3(n_task_ctas), dim3(kStreamThreads), 0, stream, a, n_task_ctas);
}

cudaError_t launch_vsl_stream(const VslArgs& a, cudaStream_t stream) {
  const bool det = a.flags & PPEA_F_DETERMINISTIC;
  if (a.flags & PPEA_F_MULTI)      // T is detached on the multi path (trainer.py:900-902)
    return det ? launch_vsl_stream_as<false, true, true>(a, stream) : launch_vsl_stream_as<false, true, false>(a, stream);
  if (a.flags & PPEA_F_GRAD_POSE)
    return det ? launch_vsl_stream_as<true, false, true>(a, stream) : launch_vsl_stream_as<true, false, false>(a, stream);
  return det ? launch_vsl_stream_as<false, false, true>(a, stream) : launch_vsl_stream_as<false, false, false>(a, stream);
}

}  // namespace ppea

#ifdef PPEA_STREAM_CLOCKS
extern "C" int ppea_debug_stream_clocks(unsigned long long* clk, unsigned* sm, int* rows, int n) {
  if (n > 8192) n = 8192;
  if (cudaMemcpyFromSymbol(rows, ppea::g_stream_rows, sizeof(int) * 2 * n) != cudaSuccess) return 1;
  if (cudaMemcpyFromSymbol(clk, ppea::g_stream_clk, sizeof(unsigned long long) * 2 * n) != cudaSuccess) return 1;
  if (cudaMemcpyFromSymbol(sm, ppea::g_stream_sm, sizeof(unsigned) * n) != cudaSuccess) return 1;
  return 0;
}
#endif
