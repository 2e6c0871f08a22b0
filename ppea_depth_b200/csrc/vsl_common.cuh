// vsl_common.cuh -- launch-side structures shared by the view-synthesis-loss kernels.
#pragma once

#include <cuda.h>            // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "../../include/ppea_vsl.h"
#include "vsl_math.cuh"

namespace ppea {

constexpr int kMaxScales = PPEA_MAX_SCALES;

struct ScaleArgs {
  int hs, ws;           // disp_s resolution
  float up_sy, up_sx;   // upsample source scales hs/H, ws/W (ATen area_pixel_compute_scale)
  const float* disp;
  const float* color;   // smoothness image (B,3,hs,ws)
  const float* noise;
  const float* mono_depth;
  float* depth;
  float* loss_px;
  uint8_t* sel;
  float* grad_disp;
  float* grad_dup;      // deterministic mode: full-res dL/d disp_up scratch (scales with hs != H only)
  float* grad_raw;      // fused step: un-normalised photometric gradient of disp_s (vsl_fused.cu)
  float* grad_raw2;     // fused step, multi path: un-normalised gradient of the consistency term
  float* grad_st;       // fused step: un-normalised smoothness stencil field of disp_s (smooth.cuh)
};

// Kernel argument block (passed by value, lives in the constant bank).
struct VslArgs {
  int B, H, W;
  int S;                // scales in this launch
  int first_scale, total_scales;
  unsigned flags;
  float disp_lo, disp_range, eps, disparity_smoothness;
  int tiles_x, tiles_y; // tiles per image (of the kernel being launched)
  int seg_rows;         // streaming step: rows per warp chunk (tiles_y = stream_pieces(H, seg_rows) tile slots per column)
  const float* tgt;
  const float* src[2];
  const float* K;
  const float* inv_K;
  const float* T[2];
  const float* cons_mask;
  const float* aug_mask;
  ScaleArgs sc[kMaxScales];
  float* partials;      // forward: [nblk][S][4] block partial sums
  float* sums;          // [S][8 + 4B]
  float* losses;        // [1 + 4S]
  float* smooth_ws;     // [S][B][kSmoothChunks][3]: disp sum, smooth_x sum, smooth_y sum
  // backward only
  const float* grad_losses;
  float* pose_partials; // [nblk_bwd][S][24]
  double* pose_sums;    // fused step: [B][S][24] per-image, per-scale sums of the pose partials (finish launch -> gradient finish)
  float* grad_T[2];
  // fused step: TMA descriptors of the colour frames viewed as (B*3, H, W) fp32 tensors, box 3 x (TH+4) x (TW+8)
  // (valid when use_tma; interior tiles are staged by cp.async.bulk.tensor, border tiles by the reflecting loop)
  int use_tma;
  // streaming step (vsl_stream.cu): per-step products of the preparation launch
  uint32_t* pk[2];      // (B,H,W) source frames packed to one RGBA8 word per pixel (exact when the frames are k/255)
  float* ident;         // (B,H,W) identity loss min_f photo(src_f, tgt) (trainer.py:1060-1069), scale-invariant
  unsigned* fmt_flag;   // != 0 after the preparation launch: some source value is not exactly k/255 -> planar fp32 gathers
  float2* ystat;        // [3][B*H*W] 3x3 window sums (Sy, Syy) of the target per channel (layers.py:243-247), scale-invariant
  alignas(64) CUtensorMap tm_tgt;
  alignas(64) CUtensorMap tm_src[2];
};

constexpr int kFwdTileW = 32;
constexpr int kFwdTileH = 16;
constexpr int kFwdThreads = 128;
constexpr int kBwdTileW = 32;
#ifndef PPEA_BWD_TILE_H
#define PPEA_BWD_TILE_H 16
#endif
constexpr int kBwdTileH = PPEA_BWD_TILE_H;
#ifndef PPEA_BWD_THREADS
#define PPEA_BWD_THREADS 128
#endif
constexpr int kBwdThreads = PPEA_BWD_THREADS;
#ifndef PPEA_SMOOTH_CHUNKS
#define PPEA_SMOOTH_CHUNKS 64
#endif
constexpr int kSmoothChunks = PPEA_SMOOTH_CHUNKS;     // role CTAs per image and scale in the smoothness kernels (column walks: more chunks = shorter walks)
constexpr int kSmoothThreads = 128;

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Streaming step (vsl_stream.cu): a warp walks a chunk of rows of the 28-column strips.  The chunk length is chosen per launch
// (stream_chunk_rows below); workspaces are sized for the shortest one.
#define PPEA_STREAM_MIN_SEG_ROWS 16
inline int fwd_blocks(int B, int H, int W) { return B * ceil_div(W, kFwdTileW) * ceil_div(H, kFwdTileH); }
inline int bwd_blocks(int B, int H, int W) { return B * ceil_div(W, kBwdTileW) * ceil_div(H, kBwdTileH); }
__host__ __device__ inline int sums_stride(int B) { return PPEA_SUMS_PER_SCALE + 4 * B; }

// Forward workspace layout (floats): [partials: nblk*S*4][smooth_ws: S*B*kSmoothChunks*3]
struct FwdWorkspace {
  size_t off_partials, off_smooth, total_floats;
};
inline FwdWorkspace fwd_workspace(int B, int H, int W, int S) {
  FwdWorkspace w;
  w.off_partials = 0;
  const int stream = B * ceil_div(W, 28) * ((H - 1) / PPEA_STREAM_MIN_SEG_ROWS + 2);     // tile slots of the streaming step (vsl_stream.cu)
  const int nblk = fwd_blocks(B, H, W) > stream ? fwd_blocks(B, H, W) : stream;
  w.off_smooth = align_up((size_t)nblk * S * 4, 4);
  w.total_floats = w.off_smooth + (size_t)S * B * kSmoothChunks * 3;
  return w;
}

// Backward workspace layout (floats): [pose partials: nblk*S*24][grad_dup: S*B*H*W when deterministic]
struct BwdWorkspace {
  size_t off_pose, off_dup, total_floats;
};
inline BwdWorkspace bwd_workspace(int B, int H, int W, int S, unsigned flags) {
  BwdWorkspace w;
  w.off_pose = 0;
  w.off_dup = align_up((size_t)bwd_blocks(B, H, W) * S * 24, 4);
  w.total_floats = w.off_dup + ((flags & PPEA_F_DETERMINISTIC) ? (size_t)S * B * H * W : 0);
  return w;
}

// Fused-step workspace layout (floats): [pose partials: nblk*S*24][raw photometric gradient of disp_s, s = 0..S-1][smoothness stencil field of disp_s, s = 0..S-1]
#ifndef PPEA_FUSED_TILE_H
#define PPEA_FUSED_TILE_H 16
#endif
#ifndef PPEA_FUSED_THREADS
#define PPEA_FUSED_THREADS 128
#endif
constexpr int kFusedTileWc = 32, kFusedTileHc = PPEA_FUSED_TILE_H;
inline int fused_blocks(int B, int H, int W) { return B * ceil_div(W, kFusedTileWc) * ceil_div(H, kFusedTileHc); }

cudaError_t launch_vsl_forward(const VslArgs& a, cudaStream_t stream);
cudaError_t launch_vsl_backward(const VslArgs& a, cudaStream_t stream);
cudaError_t launch_vsl_finish(const VslArgs& a, int nblk_fwd, cudaStream_t stream, bool pose_sums = false);
cudaError_t launch_smooth_backward(const VslArgs& a, cudaStream_t stream);
cudaError_t launch_upsample_gather(const VslArgs& a, cudaStream_t stream);
cudaError_t launch_pose_finish(const VslArgs& a, int nblk_bwd, cudaStream_t stream);
cudaError_t launch_vsl_fused(const VslArgs& a, cudaStream_t stream);
// streaming step (vsl_stream.cu): warp-per-strip row walk, no CTA barriers; tiles_x/tiles_y = strips / row segments
constexpr int kStripW = 28;           // output columns per warp (32 gathered, 30 decided)
#ifndef PPEA_STREAM_CTAS
#define PPEA_STREAM_CTAS 12
#endif
constexpr int kStreamCtasPerSm = PPEA_STREAM_CTAS;   // resident one-warp CTAs per SM (168 registers, 17 KB of shared memory each)
inline int stream_strips(int W) { return ceil_div(W, kStripW); }
// Rows per warp of the streaming step.  The rows of all (image, strip, scale) columns form one line that is cut into equal
// chunks, one per resident warp slot of the device (sm_count * kStreamCtasPerSm one-warp CTAs): a single, evenly loaded wave;
// a chunk that crosses a column end becomes two pieces (each pays four rows of run-in / run-out).  `forced` > 0 (environment
// PPEA_STREAM_SEG_ROWS, tuning only) overrides the choice.
inline int stream_chunk_rows(int B, int H, int W, int S, int sm_count, int forced) {
  if (forced > 0) return forced < PPEA_STREAM_MIN_SEG_ROWS ? PPEA_STREAM_MIN_SEG_ROWS : forced;
  const long long total = (long long)B * stream_strips(W) * S * H;
  const long long slots = (long long)(sm_count > 0 ? sm_count : 148) * kStreamCtasPerSm;
  const long long rows = (total + slots - 1) / slots;
  return rows < PPEA_STREAM_MIN_SEG_ROWS ? PPEA_STREAM_MIN_SEG_ROWS : (int)rows;
}
// tile slots per column: pieces a column of H rows can be cut into by chunks of `rows` rows
inline int stream_pieces(int H, int rows) { return (H - 1) / rows + 2; }
inline int stream_tiles_max(int B, int H, int W) { return B * stream_strips(W) * stream_pieces(H, PPEA_STREAM_MIN_SEG_ROWS); }
cudaError_t launch_vsl_prep(const VslArgs& a, cudaStream_t stream);
cudaError_t launch_vsl_stream(const VslArgs& a, cudaStream_t stream);
cudaError_t launch_vsl_smooth_tail(const VslArgs& a, cudaStream_t stream);
cudaError_t launch_vsl_grad_finish(const VslArgs& a, cudaStream_t stream);

// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl may become resident while its
// predecessor in the stream is still running; it must call grid_dependency_wait() before touching anything
// the predecessor writes.  A predecessor that calls grid_launch_dependents() early lets the dependent's CTAs
// take their place ahead of time (only worth it when those CTAs are few or the predecessor leaves SMs idle).
// Captured by stream capture as a programmatic graph edge.  Used for the small tail kernels of a step, whose
// cost is launch latency and ramp, not work.
#if defined(__CUDACC__)
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}
#endif

// Opt a kernel into > 48 KB of dynamic shared memory once per device (the attribute is sticky per
// function and device); keeping it out of the steady state also keeps it out of CUDA-graph capture.
struct SmemAttrCache {
  const void* fn[16];
  unsigned long long devmask[16];
  int n;
};
inline cudaError_t ensure_func_attr_impl(SmemAttrCache& cache, std::mutex& mu, const void* fn, cudaFuncAttribute attr, int value);
inline cudaError_t ensure_dynamic_smem_impl(const void* fn, int bytes) {
  static SmemAttrCache cache = {};
  static std::mutex mu;
  return ensure_func_attr_impl(cache, mu, fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
// Ask for the largest shared-memory carve-out once per device: a kernel whose residency is limited by its STATIC shared
// memory (the streaming step: 12 one-warp CTAs x 17 KB per SM) must not be given a smaller carve-out by the driver's
// heuristic.
inline cudaError_t ensure_max_carveout_impl(const void* fn) {
  static SmemAttrCache cache = {};
  static std::mutex mu;
  return ensure_func_attr_impl(cache, mu, fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}
template <typename Kern>
inline cudaError_t ensure_max_carveout(Kern kern) {
  return ensure_max_carveout_impl(reinterpret_cast<const void*>(kern));
}
inline cudaError_t ensure_func_attr_impl(SmemAttrCache& cache, std::mutex& mu, const void* fn, cudaFuncAttribute attr, int bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  int slot = -1;
  for (int i = 0; i < cache.n; ++i)
    if (cache.fn[i] == fn) slot = i;
  if (slot >= 0 && dev < 64 && (cache.devmask[slot] >> dev) & 1ull) return cudaSuccess;
  e = cudaFuncSetAttribute(fn, attr, bytes);
  if (e != cudaSuccess) return e;
  if (slot < 0 && cache.n < 16) {
    slot = cache.n++;
    cache.fn[slot] = fn;
    cache.devmask[slot] = 0;
  }
  if (slot >= 0 && dev < 64) cache.devmask[slot] |= 1ull << dev;
  return cudaSuccess;
}
template <typename Kern>
inline cudaError_t ensure_dynamic_smem(Kern kern, int bytes) {
  return ensure_dynamic_smem_impl(reinterpret_cast<const void*>(kern), bytes);
}

// ---- device helpers
#if defined(__CUDACC__)
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Cross-lane sums of 24 per-lane values with 24 shuffles instead of 24 x 5: every step halves the
// number of values a lane carries (the lane keeps the half selected by one bit of its id and hands the
// other half to its partner), 24 -> 12 -> 6 -> 3(+1 pad) -> 2 -> 1.  Returns the total of entry
// `warp_sum24_index(lane)`; lanes whose index is 24 or more hold padding.  Fixed summation order.
template <int N>
__device__ __forceinline__ void warp_halve(float (&v)[24], int lane, int o) {
  const bool hi = lane & o;
#pragma unroll
  for (int k = 0; k < N / 2; ++k) {
    const float send = hi ? v[k] : v[k + N / 2];
    const float keep = hi ? v[k + N / 2] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
  }
}
__device__ __forceinline__ float warp_sum24(float (&v)[24], int lane) {
  warp_halve<24>(v, lane, 16);
  warp_halve<12>(v, lane, 8);
  warp_halve<6>(v, lane, 4);
  v[3] = 0.f;
  warp_halve<4>(v, lane, 2);
  warp_halve<2>(v, lane, 1);
  return v[0];
}
__device__ __forceinline__ int warp_sum24_index(int lane) {
  const int sub = lane & 3;                                  // 3 is the padding slot
  return sub == 3 ? 24 : ((lane & 16) ? 12 : 0) + ((lane & 8) ? 6 : 0) + ((lane & 4) ? 3 : 0) + sub;
}

// Effective upstream weights of the three per-scale terms, from the gradient of the
// `losses` vector (layout in ppea_vsl.h): losses[0] = sum_s loss_s / total_scales,
// loss_s = reproj_s + cons_s + (disparity_smoothness / 2^s) * smooth_s.
struct ScaleGrads {
  float reproj, cons, smooth;
};
__device__ __forceinline__ ScaleGrads scale_grads(const VslArgs& a, int s) {
  const float* g = a.grad_losses;
  const float g_loss = g[0] / (float)a.total_scales + g[1 + s * PPEA_LOSSES_PER_SCALE + 0];
  ScaleGrads o;
  o.reproj = g_loss + g[1 + s * PPEA_LOSSES_PER_SCALE + 1];
  o.cons = g_loss + g[1 + s * PPEA_LOSSES_PER_SCALE + 2];
  o.smooth = g_loss * (a.disparity_smoothness / (float)(1 << (a.first_scale + s))) + g[1 + s * PPEA_LOSSES_PER_SCALE + 3];
  return o;
}
#endif

}  // namespace ppea
