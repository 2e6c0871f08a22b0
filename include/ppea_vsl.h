/*
 * ppea_vsl.h -- C ABI of libppea_vsl.so: the B200 (sm_100a) implementation of
 * PPEA-Depth's self-supervised view-synthesis loss path.
 *
 * The reference has no native layer and no plugin registry; the boundary it
 * offers is Python name binding.  Every entry point below
 * therefore names the reference Python call it replaces:
 *
 *   ppea_vsl_forward / ppea_vsl_backward      one pyramid scale of
 *       Trainer.generate_images_pred   /root/reference/ppeadepth/trainer.py:871-918
 *     + Trainer.compute_losses         trainer.py:1032-1160  (loop body, per scale)
 *       which inline  F.interpolate :886, disp_to_depth layers.py:14-23,
 *       BackprojectDepth layers.py:163-168, Project3D layers.py:184-199,
 *       F.grid_sample trainer.py:911-914, SSIM layers.py:243-257,
 *       compute_reprojection_loss trainer.py:995-1007, the min over sources and
 *       the selec_reproj dark-pixel rule :1076-1083, the identity loss + noise
 *       :1060-1069/:1084-1087, compute_loss_masks :1009-1027, the multi-frame
 *       mask and consistency term :1101-1141 and get_smooth_loss
 *       layers.py:210-223 with the mean-normalisation of trainer.py:1147-1149.
 *   ppea_vsl_finish / ppea_vsl_backward_finish  the scalar reductions and the
 *       multi-scale accumulation trainer.py:1113-1114, :1145-1158; dL/dT from
 *       dL/d(K@T) (autograd of layers.py:185).
 *   ppea_ssim_forward/backward                 SSIM.forward            layers.py:243-257
 *   ppea_backproject_forward/backward          BackprojectDepth.forward layers.py:163-168
 *   ppea_project3d_forward/backward            Project3D.forward       layers.py:184-199
 *   ppea_smooth_forward/backward               get_smooth_loss         layers.py:210-223
 *
 * Conventions
 *   - plain C, no torch / C++ types; all tensors are contiguous fp32 NCHW
 *     device pointers owned by the caller (PyTorch's caching allocator); the
 *     library allocates nothing, keeps no global state and is re-entrant.
 *   - every call only enqueues work on `stream` (a cudaStream_t) of the
 *     current device: no host synchronisation, no allocation => capturable in
 *     a CUDA graph.
 *   - return 0 on success; negative = argument error found on the host before
 *     any launch (PPEA_E_*); positive = cudaError_t of the launch.
 *     ppea_strerror() maps both.  Nothing throws or exits.
 */
#ifndef PPEA_VSL_H_
#define PPEA_VSL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPEA_ABI_VERSION 1

/* error codes (negative) */
#define PPEA_OK 0
#define PPEA_E_NULL (-1)       /* required pointer is NULL */
#define PPEA_E_SHAPE (-2)      /* non-positive / inconsistent / unsupported shape */
#define PPEA_E_ALIGN (-3)      /* pointer not 16-byte aligned */
#define PPEA_E_FLAGS (-4)      /* contradictory flags or missing optional input */
#define PPEA_E_VERSION (-5)    /* struct_size / abi mismatch */
#define PPEA_E_WORKSPACE (-6)  /* workspace too small */

/* flags */
#define PPEA_F_MULTI (1u << 0)         /* is_multi=True: T detached, mask = cons*(1-aug), consistency term (trainer.py:900-902, 1101-1141) */
#define PPEA_F_AUTOMASK (1u << 1)      /* not opt.disable_automasking (trainer.py:1084-1091) */
#define PPEA_F_SELEC_REPROJ (1u << 2)  /* opt.selec_reproj (default True, options.py:428-430; trainer.py:1077-1083) */
#define PPEA_F_NO_SSIM (1u << 3)       /* opt.no_ssim (trainer.py:1001-1002) */
#define PPEA_F_DETERMINISTIC (1u << 4) /* two-pass disparity-gradient scatter instead of atomics */
#define PPEA_F_MOTION_MASK (1u << 5)   /* not opt.disable_motion_masking (trainer.py:1103-1105) */
#define PPEA_F_MATCH_AUG (1u << 6)     /* not opt.no_matching_augmentation (trainer.py:1106-1108) */
#define PPEA_F_GRAD_POSE (1u << 7)     /* backward: also produce dL/dT (mono path) */

/* sel map (uint8 per full-res pixel, written by forward, read by backward):
 *   bits 0-1: source frame whose loss is propagated: 0 -> frame_ids[1] (-1),
 *             1 -> frame_ids[2] (+1), 2 -> none (both warped pixels dark => loss 0)
 *   bit  2  : automask bit (reprojection loss <= identity loss + noise); always 1
 *             when automasking is disabled or on the multi path                  */
#define PPEA_SEL_SRC_MASK 3u
#define PPEA_SEL_AUTOMASK 4u

/* number of floats in one per-scale row of the `sums` / `losses` / `grad_losses` vectors */
#define PPEA_SUMS_PER_SCALE 8
/* sums[s*8 + k]:   0: sum(r*mask) 1: sum(mask) 2: sum(|depth-mono|*(1-mask)) 3: smooth_x sum 4: smooth_y sum
 *                  5..7: reserved
 * losses[0] = loss; losses[1 + s*4 + k]: 0: loss/s  1: reproj_loss/s  2: consistency_loss/s  3: smooth term (unweighted) */
#define PPEA_LOSSES_PER_SCALE 4

typedef struct PpeaVslParams {
  uint32_t struct_size;  /* sizeof(PpeaVslParams) */
  uint32_t flags;        /* PPEA_F_* */
  int32_t batch, height, width; /* B, H, W of the images / loss maps at this call's source scale */
  int32_t disp_h, disp_w;       /* resolution of disp (== H,W at scale 0 or with v1_multiscale) */
  int32_t scale;                /* pyramid index s (smoothness weight is disparity_smoothness / 2^s) */
  int32_t num_scales;           /* S = sclm + 1 */
  int32_t smooth_h, smooth_w;   /* resolution of color_s (== disp_h, disp_w) */
  float disp_lo;     /* 1/max_depth                     (layers.py:19) */
  float disp_range;  /* 1/min_depth - 1/max_depth       (layers.py:20-21) */
  float eps;         /* Project3D eps, 1e-7             (layers.py:189) */
  float disparity_smoothness; /* opt.disparity_smoothness, 1e-3 */
  /* inputs */
  const float* disp;       /* (B,1,disp_h,disp_w)  outputs[("disp", s)] */
  const float* tgt;        /* (B,3,H,W)  inputs[("color", 0, source_scale)] */
  const float* src[2];     /* (B,3,H,W)  inputs[("color", frame_ids[1|2], source_scale)] */
  const float* K;          /* (B,4,4) */
  const float* inv_K;      /* (B,4,4) */
  const float* T[2];       /* (B,4,4)  outputs[("cam_T_cam", 0, f)] */
  const float* noise;      /* (B,1,H,W) standard normal, required iff AUTOMASK && !MULTI */
  const float* cons_mask;  /* (B,H,W)   required iff MULTI && MOTION_MASK */
  const float* aug_mask;   /* (B,)      required iff MULTI && MATCH_AUG */
  const float* mono_depth; /* (B,1,H,W) required iff MULTI */
  const float* color_s;    /* (B,3,smooth_h,smooth_w) inputs[("color", 0, s)] */
  /* outputs of forward (inputs of backward) */
  float* depth;    /* (B,1,H,W)  outputs[("depth", 0, s)] -- always materialised (trainer.py:893) */
  float* loss_px;  /* (B,1,H,W)  per-pixel reprojection loss after min/selec_reproj; may be NULL */
  uint8_t* sel;    /* (B,H,W)    selection map, see PPEA_SEL_* */
  float* sums;     /* (S*8,) device vector; this call atomically owns row `scale`; see PPEA_SUMS_PER_SCALE.
                      Forward zero-fills its row itself, then block partials are reduced into `partials`. */
  void* workspace; /* >= ppea_vsl_workspace_bytes(), 16-byte aligned, private to this (call, scale) until finish */
  size_t workspace_bytes;
} PpeaVslParams;

typedef struct PpeaVslGrads {
  uint32_t struct_size;
  const float* grad_losses; /* (1 + S*4,) upstream gradient of the `losses` vector (device) */
  const float* sums;        /* (S*8,) reduced sums written by ppea_vsl_finish */
  float* grad_disp;         /* (B,1,disp_h,disp_w): dL/d disp_s, fully overwritten */
  float* grad_pose_partials;/* workspace row for this scale: see ppea_vsl_pose_partials_bytes(); NULL iff !GRAD_POSE */
} PpeaVslGrads;

typedef struct PpeaVslFinish {
  uint32_t struct_size;
  uint32_t flags;
  int32_t batch, height, width, num_scales;
  float disparity_smoothness;
  const int32_t* disp_h;   /* host array [S] */
  const int32_t* disp_w;   /* host array [S] */
  const int32_t* map_h;    /* host array [S]: H of the loss maps of scale s (differs per scale only with v1_multiscale) */
  const int32_t* map_w;    /* host array [S] */
  void* const* workspaces; /* host array [S] of the per-scale forward workspaces */
  float* sums;             /* (S*8,) out */
  float* losses;           /* (1 + S*4,) out */
} PpeaVslFinish;

typedef struct PpeaVslPoseFinish {
  uint32_t struct_size;
  int32_t batch, num_scales;
  const int32_t* map_h;    /* host array [S] */
  const int32_t* map_w;    /* host array [S] */
  const float* K;          /* (B,4,4) of source_scale 0; with v1_multiscale pass per-scale K via K_per_scale */
  const float* const* K_per_scale; /* host array [S] or NULL (=> K for every scale) */
  const float* const* pose_partials; /* host array [S] */
  float* grad_T[2];        /* (B,4,4) out, overwritten */
} PpeaVslPoseFinish;

int ppea_abi_version(void);
const char* ppea_strerror(int code);

/* bytes of forward workspace (block partial sums + smoothness scratch) for one scale */
size_t ppea_vsl_workspace_bytes(int batch, int height, int width, int disp_h, int disp_w);
/* bytes of the per-scale pose-gradient partial buffer used by backward */
size_t ppea_vsl_pose_partials_bytes(int batch, int height, int width);
/* extra scratch (full-res dL/d disp_up) needed by PPEA_F_DETERMINISTIC backward, else 0 */
size_t ppea_vsl_backward_scratch_bytes(int batch, int height, int width, uint32_t flags);

int ppea_vsl_forward(const PpeaVslParams* p, void* stream);
int ppea_vsl_finish(const PpeaVslFinish* f, void* stream);
int ppea_vsl_backward(const PpeaVslParams* p, const PpeaVslGrads* g, void* scratch, void* stream);
int ppea_vsl_backward_finish(const PpeaVslPoseFinish* f, void* stream);

/* ---- piecewise operators behind the reference's nn.Module API ---- */
int ppea_ssim_forward(const float* x, const float* y, float* out, int n_planes, int height, int width, void* stream);
int ppea_ssim_backward(const float* x, const float* y, const float* grad_out, float* grad_x, float* grad_y,
                       int n_planes, int height, int width, void* stream);
int ppea_backproject_forward(const float* depth, const float* inv_K, float* cam_points,
                             int batch, int height, int width, void* stream);
int ppea_backproject_backward(const float* grad_cam, const float* inv_K, float* grad_depth,
                              int batch, int height, int width, void* stream);
int ppea_project3d_forward(const float* points, const float* K, const float* T, float* pix, float* z_or_null,
                           int batch, int height, int width, float eps, void* stream);
int ppea_project3d_backward(const float* points, const float* K, const float* T, const float* grad_pix,
                            const float* grad_z_or_null, float* grad_points, float* grad_P_partials,
                            int batch, int height, int width, float eps, void* stream);
size_t ppea_project3d_partials_bytes(int batch, int height, int width);
int ppea_project3d_backward_finish(const float* K, const float* T, const float* grad_P_partials, float* grad_K,
                                   float* grad_T, int batch, int height, int width, void* stream);
size_t ppea_smooth_workspace_bytes(int batch, int height, int width);
int ppea_smooth_forward(const float* disp, const float* img, float* out_scalar, void* workspace,
                        int batch, int height, int width, void* stream);
int ppea_smooth_backward(const float* disp, const float* img, const float* grad_scalar, float* grad_disp,
                         int batch, int height, int width, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PPEA_VSL_H_ */
