// api.cu -- the C ABI of libppea_vsl.so (include/ppea_vsl.h): argument validation
// and launch sequencing of the fused view-synthesis loss.  No allocation, no
// host synchronisation, no global state: every call only enqueues on `stream`.
#include "vsl_common.cuh"

#include <algorithm>
#include <cstdlib>

using namespace ppea;

namespace {

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

int check_params(const PpeaVslParams* p, bool backward) {
  if (!p) return PPEA_E_NULL;
  if (p->struct_size != sizeof(PpeaVslParams)) return PPEA_E_VERSION;
  if (p->batch <= 0 || p->height < 3 || p->width < 3) return PPEA_E_SHAPE;   // reflection pad 1 + a 3x3 window
  if (p->num_scales < 1 || p->num_scales > PPEA_MAX_SCALES || p->total_scales < 1 || p->first_scale < 0 ||
      p->first_scale + p->num_scales > 16)
    return PPEA_E_SHAPE;
  if ((size_t)p->batch * p->height * p->width >= (size_t)1 << 31) return PPEA_E_SHAPE;
  if (p->width > 32 * 120) return PPEA_E_SHAPE;   // smoothness strips (smooth.cuh): at most 32 strips of 120 columns
  const bool multi = p->flags & PPEA_F_MULTI;
  if (multi && (p->flags & PPEA_F_GRAD_POSE)) return PPEA_E_FLAGS;   // T is detached on the multi path (trainer.py:900-902)
  if (!p->tgt || !p->src[0] || !p->src[1] || !p->K || !p->inv_K || !p->T[0] || !p->T[1] || !p->sums || !p->losses)
    return PPEA_E_NULL;
  if (multi && (p->flags & PPEA_F_MOTION_MASK) && !p->cons_mask) return PPEA_E_FLAGS;
  if (multi && (p->flags & PPEA_F_MATCH_AUG) && !p->aug_mask) return PPEA_E_FLAGS;
  const void* fl[] = {p->tgt, p->src[0], p->src[1], p->K, p->inv_K, p->T[0], p->T[1], p->cons_mask, p->aug_mask, p->sums, p->losses};
  for (const void* q : fl)
    if (!aligned(q, 4)) return PPEA_E_ALIGN;
  for (int s = 0; s < p->num_scales; ++s) {
    const PpeaVslScale& sc = p->scales[s];
    if (sc.disp_h < 2 || sc.disp_w < 2 || sc.disp_h > p->height || sc.disp_w > p->width) return PPEA_E_SHAPE;
    if (!sc.disp || !sc.color || !sc.depth || !sc.sel) return PPEA_E_NULL;
    if (multi && !sc.mono_depth) return PPEA_E_FLAGS;
    if (backward && !sc.grad_disp) return PPEA_E_NULL;
    const void* fs[] = {sc.disp, sc.color, sc.noise, sc.mono_depth, sc.depth, sc.loss_px, sc.grad_disp};
    for (const void* q : fs)
      if (!aligned(q, 4)) return PPEA_E_ALIGN;
  }
  return PPEA_OK;
}

void fill_args(const PpeaVslParams* p, VslArgs& a) {
  a = VslArgs{};
  a.B = p->batch;
  a.H = p->height;
  a.W = p->width;
  a.S = p->num_scales;
  a.first_scale = p->first_scale;
  a.total_scales = p->total_scales;
  a.flags = p->flags;
  a.disp_lo = p->disp_lo;
  a.disp_range = p->disp_range;
  a.eps = p->eps;
  a.disparity_smoothness = p->disparity_smoothness;
  a.tgt = p->tgt;
  a.src[0] = p->src[0];
  a.src[1] = p->src[1];
  a.K = p->K;
  a.inv_K = p->inv_K;
  a.T[0] = p->T[0];
  a.T[1] = p->T[1];
  a.cons_mask = p->cons_mask;
  a.aug_mask = p->aug_mask;
  for (int s = 0; s < p->num_scales; ++s) {
    const PpeaVslScale& in = p->scales[s];
    ScaleArgs& sc = a.sc[s];
    sc.hs = in.disp_h;
    sc.ws = in.disp_w;
    sc.up_sy = up_scale(in.disp_h, p->height);
    sc.up_sx = up_scale(in.disp_w, p->width);
    sc.disp = in.disp;
    sc.color = in.color;
    sc.noise = in.noise;
    sc.mono_depth = in.mono_depth;
    sc.depth = in.depth;
    sc.loss_px = in.loss_px;
    sc.sel = in.sel;
    sc.grad_disp = in.grad_disp;
    sc.grad_dup = nullptr;
    sc.grad_raw = nullptr;
    sc.grad_raw2 = nullptr;
    sc.grad_st = nullptr;
  }
  a.sums = p->sums;
  a.losses = p->losses;
}

#define PPEA_TRACE(p, i)                                                       \
  do {                                                                         \
    if ((p)->trace_events) {                                                   \
      cudaError_t e__ = cudaEventRecord((cudaEvent_t)(p)->trace_events[i], stream); \
      if (e__ != cudaSuccess) return (int)e__;                                 \
    }                                                                          \
  } while (0)

#define PPEA_TRY(expr)                \
  do {                                \
    cudaError_t e__ = (expr);         \
    if (e__ != cudaSuccess) return (int)e__; \
  } while (0)

}  // namespace

extern "C" {

int ppea_abi_version(void) { return PPEA_ABI_VERSION; }

const char* ppea_strerror(int code) {
  switch (code) {
    case PPEA_OK: return "success";
    case PPEA_E_NULL: return "ppea: required pointer is NULL";
    case PPEA_E_SHAPE: return "ppea: non-positive, inconsistent or unsupported shape";
    case PPEA_E_ALIGN: return "ppea: misaligned pointer";
    case PPEA_E_FLAGS: return "ppea: contradictory flags or missing optional input";
    case PPEA_E_VERSION: return "ppea: struct_size / ABI version mismatch";
    case PPEA_E_WORKSPACE: return "ppea: workspace too small or misaligned";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "ppea: unknown error";
}

void* ppea_event_create(void) {
  cudaEvent_t ev = nullptr;
  if (cudaEventCreate(&ev) != cudaSuccess) return nullptr;
  return (void*)ev;
}
void ppea_event_destroy(void* event) {
  if (event) cudaEventDestroy((cudaEvent_t)event);
}
int ppea_event_record(void* event, void* stream) {
  if (!event) return PPEA_E_NULL;
  return (int)cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream);
}
int ppea_event_elapsed_ms(void* start, void* stop, float* ms) {
  if (!start || !stop || !ms) return PPEA_E_NULL;
  cudaError_t e = cudaEventSynchronize((cudaEvent_t)stop);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop);
}

size_t ppea_vsl_workspace_bytes(int batch, int height, int width, int num_scales) {
  if (batch <= 0 || height <= 0 || width <= 0 || num_scales <= 0) return 0;
  return fwd_workspace(batch, height, width, num_scales).total_floats * sizeof(float);
}

size_t ppea_vsl_backward_workspace_bytes(int batch, int height, int width, int num_scales, uint32_t flags) {
  if (batch <= 0 || height <= 0 || width <= 0 || num_scales <= 0) return 0;
  return bwd_workspace(batch, height, width, num_scales, flags).total_floats * sizeof(float);
}

size_t ppea_vsl_sums_floats(int batch, int num_scales) {
  if (batch <= 0 || num_scales <= 0) return 0;
  return (size_t)num_scales * sums_stride(batch);
}

int ppea_vsl_forward(const PpeaVslParams* p, void* stream_) {
  const int rc = check_params(p, false);
  if (rc != PPEA_OK) return rc;
  const FwdWorkspace ws = fwd_workspace(p->batch, p->height, p->width, p->num_scales);
  if (!p->workspace) return PPEA_E_NULL;
  if (!aligned(p->workspace, 16) || p->workspace_bytes < ws.total_floats * sizeof(float)) return PPEA_E_WORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  VslArgs a;
  fill_args(p, a);
  a.tiles_x = ceil_div(a.W, kFwdTileW);
  a.tiles_y = ceil_div(a.H, kFwdTileH);
  a.partials = (float*)p->workspace + ws.off_partials;
  a.smooth_ws = (float*)p->workspace + ws.off_smooth;
  PPEA_TRACE(p, 0);
  PPEA_TRACE(p, 1);
  PPEA_TRY(launch_vsl_forward(a, stream));      // tiles of every scale + the smoothness CTAs
  PPEA_TRACE(p, 2);
  PPEA_TRACE(p, 3);
  PPEA_TRY(launch_vsl_finish(a, fwd_blocks(a.B, a.H, a.W), stream));
  PPEA_TRACE(p, 4);
  return PPEA_OK;
}

int ppea_vsl_backward(const PpeaVslParams* p, const PpeaVslGrads* g, void* stream_) {
  const int rc = check_params(p, true);
  if (rc != PPEA_OK) return rc;
  if (!g) return PPEA_E_NULL;
  if (g->struct_size != sizeof(PpeaVslGrads)) return PPEA_E_VERSION;
  if (!g->grad_losses || !g->workspace) return PPEA_E_NULL;
  const bool pose = p->flags & PPEA_F_GRAD_POSE;
  if (pose && (!g->grad_T[0] || !g->grad_T[1])) return PPEA_E_NULL;
  const BwdWorkspace ws = bwd_workspace(p->batch, p->height, p->width, p->num_scales, p->flags);
  if (!aligned(g->workspace, 16) || g->workspace_bytes < ws.total_floats * sizeof(float)) return PPEA_E_WORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  VslArgs a;
  fill_args(p, a);
  a.tiles_x = ceil_div(a.W, kBwdTileW);
  a.tiles_y = ceil_div(a.H, kBwdTileH);
  a.grad_losses = g->grad_losses;
  a.pose_partials = (float*)g->workspace + ws.off_pose;
  a.grad_T[0] = g->grad_T[0];
  a.grad_T[1] = g->grad_T[1];
  bool any_dup = false;
  if (p->flags & PPEA_F_DETERMINISTIC) {
    const size_t n = (size_t)a.B * a.H * a.W;
    for (int s = 0; s < a.S; ++s)
      if (a.sc[s].hs != a.H || a.sc[s].ws != a.W) {
        a.sc[s].grad_dup = (float*)g->workspace + ws.off_dup + (size_t)s * n;
        any_dup = true;
      }
  }
  PPEA_TRACE(p, 0);
  if (p->flags & PPEA_F_DETERMINISTIC) {
    PPEA_TRY(launch_smooth_backward(a, stream));          // overwrite, then fixed-order adds on top
  } else if (!(p->flags & PPEA_F_GRAD_PREZEROED)) {
    // every contribution (tile CTAs and smoothness CTAs of the fused launch) is accumulated atomically
    for (int s = 0; s < a.S; ++s)
      PPEA_TRY(cudaMemsetAsync(a.sc[s].grad_disp, 0, sizeof(float) * (size_t)a.B * a.sc[s].hs * a.sc[s].ws, stream));
  }
  PPEA_TRACE(p, 1);
  PPEA_TRY(launch_vsl_backward(a, stream));
  PPEA_TRACE(p, 2);
  if (any_dup) PPEA_TRY(launch_upsample_gather(a, stream));
  PPEA_TRACE(p, 3);
  if (pose) PPEA_TRY(launch_pose_finish(a, bwd_blocks(a.B, a.H, a.W), stream));
  PPEA_TRACE(p, 4);
  return PPEA_OK;
}

// ---- fused training step (mono path): vsl_fused.cu -------------------------------------------------
static_assert(kFusedTileWc == kFwdTileW && kFusedTileHc >= kFwdTileH, "the fused step reuses the forward workspace layout (never more tiles than the forward)");

struct FusedWorkspace {
  size_t off_pose, off_pose_sums, off_raw[kMaxScales], off_raw2[kMaxScales], off_st[kMaxScales], raw_floats[kMaxScales], total_floats;
  size_t off_flag, off_ident, off_pk[2], off_ystat;       // streaming step: format flag (4 floats), identity-loss map, packed sources, target window sums
};
// Layout (floats): [pose partials: nblk*S*24][per-image pose sums: B*S*24 doubles][raw photometric gradient of disp_s][multi path: raw consistency
// gradient of disp_s][smoothness stencil field of disp_s].  A coarse-scale raw field takes 2 floats per pixel:
// with PPEA_F_DETERMINISTIC it is an array of 64-bit fixed-point accumulators.
static FusedWorkspace fused_workspace(const PpeaVslParams* p) {
  FusedWorkspace w;
  w.off_pose = 0;
  const size_t n_tiles = (size_t)std::max(fused_blocks(p->batch, p->height, p->width), stream_tiles_max(p->batch, p->height, p->width));
  size_t off = align_up(n_tiles * p->num_scales * 24, 4);
  w.off_pose_sums = off;                                             // [B][S][24] doubles
  off += align_up((size_t)p->batch * p->num_scales * 24 * 2, 4);
  for (int s = 0; s < kMaxScales; ++s) {
    w.raw_floats[s] = 0;
    if (s < p->num_scales) {
      const size_t n = (size_t)p->batch * p->scales[s].disp_h * p->scales[s].disp_w;
      const bool coarse = p->scales[s].disp_h != p->height || p->scales[s].disp_w != p->width;
      w.raw_floats[s] = align_up(coarse ? 2 * n : n, 4);
    }
  }
  for (int s = 0; s < kMaxScales; ++s) {
    w.off_raw[s] = off;
    off += w.raw_floats[s];
  }
  for (int s = 0; s < kMaxScales; ++s) {
    w.off_raw2[s] = off;
    if (p->flags & PPEA_F_MULTI) off += w.raw_floats[s];
  }
  for (int s = 0; s < kMaxScales; ++s) {
    w.off_st[s] = off;
    if (s < p->num_scales) off += align_up((size_t)p->batch * p->scales[s].disp_h * p->scales[s].disp_w, 4);
  }
  const size_t n_px = align_up((size_t)p->batch * p->height * p->width, 4);
  w.off_flag = off;
  off += 4;
  w.off_ident = off;
  off += n_px;
  w.off_pk[0] = off;
  off += n_px;
  w.off_pk[1] = off;
  off += n_px;
  w.off_ystat = off;
  off += 6 * n_px;
  w.total_floats = off;
  return w;
}

static int check_fused(const PpeaVslParams* p, const PpeaVslFused* f, bool backward) {
  const int rc = check_params(p, backward);
  if (rc != PPEA_OK) return rc;
  if (!f) return PPEA_E_NULL;
  if (f->struct_size != sizeof(PpeaVslFused)) return PPEA_E_VERSION;
  if (!f->workspace) return PPEA_E_NULL;
  if (!aligned(f->workspace, 16) || f->workspace_bytes < fused_workspace(p).total_floats * sizeof(float)) return PPEA_E_WORKSPACE;
  return PPEA_OK;
}

// TMA descriptor of one colour frame batch viewed as a (B*3, H, W) fp32 tensor, box = 3 planes x (TH+4) rows x (TW+8)
// columns (the fused kernel's staging tile; TMA wants a 16-byte aligned start, so two unused columns on each side).  The encoder is a driver entry point (no libcuda link: fetched through the
// statically linked runtime).  Returns false when the tensor cannot be described (row pitch not a multiple of 16 bytes,
// unaligned base, driver without the entry point): the kernel then stages every tile with its reflecting loop.
typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeTiledFn tensor_map_encoder() {
  static std::once_flag once;
  static TensorMapEncodeTiledFn fn = nullptr;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TensorMapEncodeTiledFn>(sym);
    (void)cudaGetLastError();
  });
  return fn;
}
static bool frame_tensor_map(CUtensorMap* tm, const float* base, int B, int H, int W) {
  TensorMapEncodeTiledFn enc = tensor_map_encoder();
  if (!enc || (W % 4) != 0 || !aligned(base, 16)) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * 3};
  const cuuint64_t strides[2] = {(cuuint64_t)W * sizeof(float), (cuuint64_t)W * H * sizeof(float)};
  const cuuint32_t box[3] = {(cuuint32_t)kFusedTileWc + 8, (cuuint32_t)kFusedTileHc + 4, 3};   // starts at the aligned column x0 - 4
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool use_tiles(const PpeaVslParams* p) { return p->flags & PPEA_F_FUSED_TILES; }
// rows per warp chunk of the streaming step on the current device (vsl_common.cuh stream_chunk_rows)
static int stream_rows_for(const PpeaVslParams* p) {
  static int sm_count[64] = {};
  static int forced = -1;
  int dev = 0;
  (void)cudaGetDevice(&dev);
  if (forced < 0) {
    const char* e = getenv("PPEA_STREAM_SEG_ROWS");
    forced = e ? atoi(e) : 0;
  }
  int sms = 0;
  if (dev >= 0 && dev < 64 && sm_count[dev] > 0) {
    sms = sm_count[dev];
  } else {
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148, (void)cudaGetLastError();
    if (dev >= 0 && dev < 64) sm_count[dev] = sms;
  }
  return stream_chunk_rows(p->batch, p->height, p->width, p->num_scales, sms, forced);
}
static int fused_tiles(const PpeaVslParams* p) {
  return use_tiles(p) ? fused_blocks(p->batch, p->height, p->width)
                      : p->batch * stream_strips(p->width) * stream_pieces(p->height, stream_rows_for(p));
}

static void fused_args(const PpeaVslParams* p, const PpeaVslFused* f, VslArgs& a, bool forward) {
  fill_args(p, a);
  const bool tiles = use_tiles(p);
  a.use_tma = (tiles && forward && frame_tensor_map(&a.tm_tgt, a.tgt, a.B, a.H, a.W) && frame_tensor_map(&a.tm_src[0], a.src[0], a.B, a.H, a.W) &&
               frame_tensor_map(&a.tm_src[1], a.src[1], a.B, a.H, a.W))
                  ? 1
                  : 0;
  const FusedWorkspace fw = fused_workspace(p);
  const FwdWorkspace ws = fwd_workspace(p->batch, p->height, p->width, p->num_scales);
  a.tiles_x = tiles ? ceil_div(a.W, kFusedTileWc) : stream_strips(a.W);
  a.seg_rows = stream_rows_for(p);
  a.tiles_y = tiles ? ceil_div(a.H, kFusedTileHc) : stream_pieces(a.H, a.seg_rows);
  a.fmt_flag = reinterpret_cast<unsigned*>((float*)f->workspace + fw.off_flag);
  a.ident = (float*)f->workspace + fw.off_ident;
  a.pk[0] = reinterpret_cast<uint32_t*>((float*)f->workspace + fw.off_pk[0]);
  a.pk[1] = reinterpret_cast<uint32_t*>((float*)f->workspace + fw.off_pk[1]);
  a.ystat = reinterpret_cast<float2*>((float*)f->workspace + fw.off_ystat);
  a.partials = (float*)p->workspace + ws.off_partials;
  a.smooth_ws = (float*)p->workspace + ws.off_smooth;
  a.pose_partials = (float*)f->workspace + fw.off_pose;
  a.pose_sums = reinterpret_cast<double*>((float*)f->workspace + fw.off_pose_sums);
  for (int s = 0; s < a.S; ++s) {
    a.sc[s].grad_raw = (float*)f->workspace + fw.off_raw[s];
    a.sc[s].grad_raw2 = (p->flags & PPEA_F_MULTI) ? (float*)f->workspace + fw.off_raw2[s] : nullptr;
    a.sc[s].grad_st = (float*)f->workspace + fw.off_st[s];
  }
}

size_t ppea_vsl_fused_workspace_bytes(const PpeaVslParams* p) {
  if (!p || p->struct_size != sizeof(PpeaVslParams) || p->batch <= 0 || p->height <= 0 || p->width <= 0 ||
      p->num_scales < 1 || p->num_scales > PPEA_MAX_SCALES)
    return 0;
  return fused_workspace(p).total_floats * sizeof(float);
}

int ppea_vsl_fused_forward(const PpeaVslParams* p, const PpeaVslFused* f, void* stream_) {
  const int rc = check_fused(p, f, false);
  if (rc != PPEA_OK) return rc;
  const FwdWorkspace ws = fwd_workspace(p->batch, p->height, p->width, p->num_scales);
  if (!p->workspace) return PPEA_E_NULL;
  if (!aligned(p->workspace, 16) || p->workspace_bytes < ws.total_floats * sizeof(float)) return PPEA_E_WORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  VslArgs a;
  fused_args(p, f, a, true);
  for (int s = 0; s < a.S; ++s) a.sc[s].grad_disp = nullptr;     // (the smoothness role would pre-zero it)
  PPEA_TRACE(p, 0);
  if (!(p->flags & PPEA_F_RAW_PREZEROED)) {                      // coarse scales accumulate atomically
    const FusedWorkspace fw = fused_workspace(p);
    PPEA_TRY(cudaMemsetAsync(a.fmt_flag, 0, 16, stream));        // (the finish launch leaves it cleared for the next step)
    for (int s = 0; s < a.S; ++s)
      if (a.sc[s].hs != a.H || a.sc[s].ws != a.W) {
        PPEA_TRY(cudaMemsetAsync(a.sc[s].grad_raw, 0, sizeof(float) * fw.raw_floats[s], stream));
        if (a.sc[s].grad_raw2) PPEA_TRY(cudaMemsetAsync(a.sc[s].grad_raw2, 0, sizeof(float) * fw.raw_floats[s], stream));
      }
  }
  PPEA_TRACE(p, 1);
  if (use_tiles(p)) {
    a.fmt_flag = nullptr;
    PPEA_TRY(launch_vsl_fused(a, stream));
    PPEA_TRACE(p, 2);
  } else {
    PPEA_TRY(launch_vsl_prep(a, stream));       // packed sources + identity loss, once for all scales
    PPEA_TRACE(p, 2);
    PPEA_TRY(launch_vsl_stream(a, stream));
    PPEA_TRY(launch_vsl_smooth_tail(a, stream));   // smoothness term, in the shadow of the streaming kernel's last warps
  }
  PPEA_TRACE(p, 3);
  PPEA_TRY(launch_vsl_finish(a, fused_tiles(p), stream, (p->flags & PPEA_F_GRAD_POSE) && !(p->flags & PPEA_F_MULTI)));
  PPEA_TRACE(p, 4);
  return PPEA_OK;
}

int ppea_vsl_fused_backward(const PpeaVslParams* p, const PpeaVslGrads* g, const PpeaVslFused* f, void* stream_) {
  const int rc = check_fused(p, f, true);
  if (rc != PPEA_OK) return rc;
  if (!g) return PPEA_E_NULL;
  if (g->struct_size != sizeof(PpeaVslGrads)) return PPEA_E_VERSION;
  if (!g->grad_losses) return PPEA_E_NULL;
  const bool pose = (p->flags & PPEA_F_GRAD_POSE) && !(p->flags & PPEA_F_MULTI);   // T is detached on the multi path
  if (pose && (!g->grad_T[0] || !g->grad_T[1])) return PPEA_E_NULL;
  cudaStream_t stream = (cudaStream_t)stream_;
  VslArgs a;
  fused_args(p, f, a, false);
  a.grad_losses = g->grad_losses;
  a.grad_T[0] = g->grad_T[0];
  a.grad_T[1] = g->grad_T[1];
  PPEA_TRACE(p, 0);
  PPEA_TRACE(p, 1);
  PPEA_TRY(launch_vsl_grad_finish(a, stream));
  PPEA_TRACE(p, 2);
  PPEA_TRACE(p, 3);
  PPEA_TRACE(p, 4);
  return PPEA_OK;
}

}  // extern "C"
