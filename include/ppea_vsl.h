/*
 * ppea_vsl.h -- C ABI of libppea_vsl.so: the B200 (sm_100a) implementation of
 * PPEA-Depth's self-supervised view-synthesis loss path.
 *
 * The reference has no native layer and no plugin registry; the boundary it
 * offers is Python name binding.  Every entry point below
 * therefore names the reference Python call it replaces (paths relative to
 * /root/reference/ppeadepth):
 *
 *   ppea_vsl_forward / ppea_vsl_backward
 *       Trainer.generate_images_pred   trainer.py:871-918
 *     + Trainer.compute_losses         trainer.py:1032-1160
 *       for ALL pyramid scales of one call (range(opt.sclm+1), trainer.py:881/1041),
 *       which inline  F.interpolate :886, disp_to_depth layers.py:14-23,
 *       BackprojectDepth layers.py:163-168, Project3D layers.py:184-199,
 *       F.grid_sample trainer.py:911-914, SSIM layers.py:243-257,
 *       compute_reprojection_loss trainer.py:995-1007, the min over sources and
 *       the selec_reproj dark-pixel rule :1076-1083, the identity loss + noise
 *       :1060-1069/:1084-1087, compute_loss_masks :1009-1027, the multi-frame
 *       mask and consistency term :1101-1141, get_smooth_loss layers.py:210-223
 *       with the mean-normalisation of trainer.py:1147-1149, the masked mean
 *       :1113-1114 and the multi-scale accumulation :1145-1158; backward also
 *       applies autograd of layers.py:185 (dL/dT from dL/d(K@T)).
 *   ppea_ssim_forward/backward                 SSIM.forward             layers.py:243-257
 *   ppea_backproject_forward/backward          BackprojectDepth.forward layers.py:163-168
 *   ppea_project3d_forward/backward            Project3D.forward        layers.py:184-199
 *   ppea_smooth_forward/backward               get_smooth_loss          layers.py:210-223
 *   ppea_warp_forward/backward                 F.grid_sample(border, align_corners=True) trainer.py:911-914
 *   ppea_reprojection_forward/backward         Trainer.compute_reprojection_loss trainer.py:995-1007
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

#define PPEA_ABI_VERSION 11

/* error codes (negative) */
#define PPEA_OK 0
#define PPEA_E_NULL (-1)       /* required pointer is NULL */
#define PPEA_E_SHAPE (-2)      /* non-positive / inconsistent / unsupported shape */
#define PPEA_E_ALIGN (-3)      /* pointer not 4-byte (float) / 16-byte (workspace) aligned */
#define PPEA_E_FLAGS (-4)      /* contradictory flags or missing optional input */
#define PPEA_E_VERSION (-5)    /* struct_size / abi mismatch */
#define PPEA_E_WORKSPACE (-6)  /* workspace too small */

/* flags */
#define PPEA_F_MULTI (1u << 0)         /* is_multi=True: T detached, mask = cons*(1-aug), consistency term (trainer.py:900-902, 1101-1141) */
#define PPEA_F_AUTOMASK (1u << 1)      /* mask = [reproj <= identity (+ noise)] on the mono path.  The reference ALWAYS applies it:
                                          opt.disable_automasking only drops the noise (trainer.py:1084-1091) => pass noise = NULL */
#define PPEA_F_SELEC_REPROJ (1u << 2)  /* opt.selec_reproj (default True, options.py:428-430; trainer.py:1077-1083) */
#define PPEA_F_NO_SSIM (1u << 3)       /* opt.no_ssim (trainer.py:1001-1002) */
#define PPEA_F_DETERMINISTIC (1u << 4) /* bit-reproducible disparity gradient: two-pass gather (ppea_vsl_backward) or 64-bit fixed-point
                                          accumulation (fused step) instead of float atomics */
#define PPEA_F_MOTION_MASK (1u << 5)   /* not opt.disable_motion_masking (trainer.py:1103-1105) */
#define PPEA_F_MATCH_AUG (1u << 6)     /* not opt.no_matching_augmentation (trainer.py:1106-1108) */
#define PPEA_F_GRAD_POSE (1u << 7)     /* backward: also produce dL/dT (mono path) */
#define PPEA_F_GRAD_PREZEROED (1u << 8) /* backward: every grad_disp was zero-filled by the forward (which does so when it is
                                           handed non-NULL grad_disp pointers); skips the backward's own zero-fill */

#define PPEA_F_RAW_PREZEROED (1u << 9)  /* fused step: the coarse-scale raw gradient fields inside PpeaVslFused.workspace are zero on
                                           entry of ppea_vsl_fused_forward (zero-filled once by the caller); ppea_vsl_fused_backward
                                           then clears them again on its way out, so a replayed forward/backward pair needs no
                                           memset.  Every fused_forward must be followed by exactly one fused_backward. */

#define PPEA_F_FUSED_TILES (1u << 10)   /* fused step: run the round-1 shared-memory tile kernel (vsl_fused.cu, TMA-staged tiles) instead of
                                           the warp-streaming kernel (vsl_stream.cu); same results to the documented tolerances.  Pass it to
                                           ppea_vsl_fused_workspace_bytes / _forward / _backward alike. */

/* sel map (uint8 per full-res pixel, written by forward, read by backward):
 *   bits 0-1: source frame whose loss is propagated: 0 -> frame_ids[1] (-1),
 *             1 -> frame_ids[2] (+1), 2 -> none (both warped pixels dark => loss 0)
 *   bit  2  : automask bit (reprojection loss <= identity loss + noise); always 1
 *             when automasking is disabled or on the multi path                  */
#define PPEA_SEL_SRC_MASK 3u
#define PPEA_SEL_AUTOMASK 4u

#define PPEA_MAX_SCALES 4

/* sums: device vector of num_scales rows of (PPEA_SUMS_PER_SCALE + 4*batch) floats, written by forward,
 * read by backward.  row[0] sum(r*mask)  row[1] sum(mask)  row[2] sum(|depth-mono|*(1-mask))
 *   row[3] smooth_x sum  row[4] smooth_y sum  row[5..7] reserved
 *   row[8 + 4*b + k]: per image b: 0 sum(disp_s)  1 raw smooth_x sum  2 raw smooth_y sum  3 image b's share of the smoothness term
 * losses: device vector (1 + 4*num_scales):
 *   losses[0] = sum_s loss_s / total_scales
 *   losses[1 + 4*s + k]: 0 loss/s  1 reproj_loss/s  2 consistency_loss/s  3 smoothness term of scale s (unweighted) */
#define PPEA_SUMS_PER_SCALE 8
#define PPEA_LOSSES_PER_SCALE 4

typedef struct PpeaVslScale {
  int32_t disp_h, disp_w;  /* resolution of disp / color for this scale (H>>s, W>>s; == H,W with v1_multiscale) */
  const float* disp;       /* (B,1,disp_h,disp_w)  outputs[("disp", s)] */
  const float* color;      /* (B,3,disp_h,disp_w)  inputs[("color", 0, s)] -- the smoothness image */
  const float* noise;      /* (B,1,H,W) standard normal tie-break draws (AUTOMASK && !MULTI); NULL = no noise */
  const float* mono_depth; /* (B,1,H,W) required iff MULTI */
  float* depth;            /* out (B,1,H,W)  outputs[("depth", 0, s)] -- always materialised (trainer.py:893) */
  float* loss_px;          /* out (B,1,H,W)  per-pixel reprojection loss after min/selec_reproj; may be NULL */
  uint8_t* sel;            /* out (B,H,W)    selection map, see PPEA_SEL_* (input of backward) */
  float* grad_disp;        /* backward out (B,1,disp_h,disp_w), fully overwritten.  Forward: NULL, or the same buffer to have
                              it zero-filled on the fly (then pass PPEA_F_GRAD_PREZEROED to backward) */
} PpeaVslScale;

typedef struct PpeaVslParams {
  uint32_t struct_size;  /* sizeof(PpeaVslParams) */
  uint32_t flags;        /* PPEA_F_* */
  int32_t batch, height, width; /* B, H, W of the images / loss maps (source_scale 0) */
  int32_t num_scales;    /* scales handled by this call, 1..PPEA_MAX_SCALES */
  int32_t first_scale;   /* pyramid index of scales[0] (smoothness weight is disparity_smoothness / 2^(first_scale+i)) */
  int32_t total_scales;  /* divisor of the final loss, opt.sclm + 1 (trainer.py:1157) */
  float disp_lo;         /* 1/max_depth                     (layers.py:19) */
  float disp_range;      /* 1/min_depth - 1/max_depth       (layers.py:20-21) */
  float eps;             /* Project3D eps, 1e-7             (layers.py:189) */
  float disparity_smoothness; /* opt.disparity_smoothness, 1e-3 */
  const float* tgt;        /* (B,3,H,W)  inputs[("color", 0, 0)] */
  const float* src[2];     /* (B,3,H,W)  inputs[("color", frame_ids[1|2], 0)] */
  const float* K;          /* (B,4,4) */
  const float* inv_K;      /* (B,4,4) */
  const float* T[2];       /* (B,4,4)  outputs[("cam_T_cam", 0, f)] */
  const float* cons_mask;  /* (B,H,W)   required iff MULTI && MOTION_MASK */
  const float* aug_mask;   /* (B,)      required iff MULTI && MATCH_AUG */
  PpeaVslScale scales[PPEA_MAX_SCALES];
  float* sums;             /* out: see PPEA_SUMS_PER_SCALE */
  float* losses;           /* out: see PPEA_LOSSES_PER_SCALE */
  void* workspace;         /* >= ppea_vsl_workspace_bytes(), 16-byte aligned; private to this call until it completes */
  size_t workspace_bytes;
  void* const* trace_events; /* NULL, or PPEA_TRACE_EVENTS cudaEvent_t handles (ppea_event_create) recorded on `stream`:
                                [0] before the first kernel, then after each stage --
                                forward:  [1] (unused)  [2] fused forward kernel  [3] (unused)  [4] finish
                                backward: [1] grad zero-fill (or smoothness backward when deterministic)  [2] fused backward kernel  [3] upsample gather  [4] pose finish;
                                fused step: forward [1] field zero-fill (unless RAW_PREZEROED)  [2] fused kernel  [4] finish; backward [2] gradient finish */
} PpeaVslParams;

typedef struct PpeaVslGrads {
  uint32_t struct_size;
  const float* grad_losses; /* (1 + 4*num_scales,) device: upstream gradient of every entry of `losses` */
  float* grad_T[2];         /* (B,4,4) out, overwritten; required iff GRAD_POSE */
  void* workspace;          /* >= ppea_vsl_backward_workspace_bytes(), 16-byte aligned */
  size_t workspace_bytes;
} PpeaVslGrads;

#define PPEA_TRACE_EVENTS 5

int ppea_abi_version(void);

/* profiling helpers for bench.py: timing events created/destroyed by the caller through the library's
 * own CUDA runtime instance (the library is linked against the static cudart). */
void* ppea_event_create(void);
void ppea_event_destroy(void* event);
int ppea_event_record(void* event, void* stream);
int ppea_event_elapsed_ms(void* start, void* stop, float* ms); /* synchronises on `stop` */
const char* ppea_strerror(int code);

/* forward workspace (block partial sums) */
size_t ppea_vsl_workspace_bytes(int batch, int height, int width, int num_scales);
/* backward workspace: pose-gradient block partials (+ full-res dL/d disp_up per scale with PPEA_F_DETERMINISTIC) */
size_t ppea_vsl_backward_workspace_bytes(int batch, int height, int width, int num_scales, uint32_t flags);
/* floats in the `sums` vector */
size_t ppea_vsl_sums_floats(int batch, int num_scales);

int ppea_vsl_forward(const PpeaVslParams* p, void* stream);
/* `p` must describe the same call as the forward (same inputs, sel and sums as written by it) */
int ppea_vsl_backward(const PpeaVslParams* p, const PpeaVslGrads* g, void* stream);

/* ---- fused training step (mono and multi path, atomic or deterministic backward) -----------------
 * Same reference calls as ppea_vsl_forward + ppea_vsl_backward (Trainer.generate_images_pred
 * trainer.py:871-918 + Trainer.compute_losses :1032-1160 + autograd of both), organised as ONE main
 * launch: the masked-mean normaliser 1/(sum(mask_s) + 1e-7) (trainer.py:1113-1114) is the only global
 * quantity of the path, so ppea_vsl_fused_forward evaluates the loss AND the un-normalised gradient
 * fields (into `workspace`: photometric term, smoothness stencil, and on the multi path the consistency
 * term trainer.py:1128-1132), and ppea_vsl_fused_backward only combines them with the upstream gradient
 * of `losses` and reduces the pose gradient.  Outputs (depth, sel, loss_px, sums, losses, grad_disp,
 * grad_T) are the same tensors, to the same tolerances, as the two-call path.
 * PPEA_F_DETERMINISTIC: the coarse-scale fields are accumulated as 64-bit fixed-point integers (2^-40
 * resolution, contributions clamped to +-8.3e6), so the gradients are bit-reproducible without a
 * full-resolution scratch field or a second pass (and, on the streaming step, faster than float atomics:
 * the host mirror makes it the default).  A contribution that is not a finite number within +-8.3e6
 * cannot be represented: on the streaming step it raises a sticky word in the workspace instead and
 * ppea_vsl_fused_backward returns NaN for the coarse-scale gradients of that step (what float atomics
 * would carry in the cells it reaches); the tile kernel (PPEA_F_FUSED_TILES) saturates it.  Pass PPEA_F_GRAD_POSE to BOTH calls if dL/dT is
 * wanted (ignored with PPEA_F_MULTI: T is detached there, trainer.py:900-902). */
typedef struct PpeaVslFused {
  uint32_t struct_size;
  void* workspace;        /* >= ppea_vsl_fused_workspace_bytes(p), 16-byte aligned; written by fused_forward,
                             read by fused_backward (keep it untouched in between) */
  size_t workspace_bytes;
} PpeaVslFused;
size_t ppea_vsl_fused_workspace_bytes(const PpeaVslParams* p);
int ppea_vsl_fused_forward(const PpeaVslParams* p, const PpeaVslFused* f, void* stream);
int ppea_vsl_fused_backward(const PpeaVslParams* p, const PpeaVslGrads* g, const PpeaVslFused* f, void* stream);

/* ---- piecewise operators behind the reference's nn.Module / function API ---- */
int ppea_ssim_forward(const float* x, const float* y, float* out, int n_planes, int height, int width, void* stream);
int ppea_ssim_backward(const float* x, const float* y, const float* grad_out, float* grad_x, float* grad_y,
                       int n_planes, int height, int width, void* stream);
/* pred, target (B,3,H,W) -> out (B,1,H,W); backward gives d/d pred only (the target is data) */
int ppea_reprojection_forward(const float* pred, const float* target, float* out, int batch, int height, int width,
                              int no_ssim, void* stream);
int ppea_reprojection_backward(const float* pred, const float* target, const float* grad_out, float* grad_pred,
                               int batch, int height, int width, int no_ssim, void* stream);
int ppea_backproject_forward(const float* depth, const float* inv_K, float* cam_points,
                             int batch, int height, int width, void* stream);
int ppea_backproject_backward(const float* grad_cam, const float* inv_K, float* grad_depth,
                              int batch, int height, int width, void* stream);
int ppea_project3d_forward(const float* points, const float* K, const float* T, float* pix, float* z_or_null,
                           int batch, int height, int width, float eps, void* stream);
size_t ppea_project3d_partials_bytes(int batch, int height, int width);
int ppea_project3d_backward(const float* points, const float* K, const float* T, const float* grad_pix,
                            const float* grad_z_or_null, float* grad_points, float* grad_T, void* partials,
                            int batch, int height, int width, float eps, void* stream);
int ppea_warp_forward(const float* src, const float* grid, float* out, int batch, int channels, int height, int width,
                      int out_h, int out_w, void* stream);
int ppea_warp_backward(const float* src, const float* grid, const float* grad_out, float* grad_grid,
                       int batch, int channels, int height, int width, int out_h, int out_w, void* stream);
size_t ppea_smooth_workspace_bytes(int batch, int height, int width);
int ppea_smooth_forward(const float* disp, const float* img, float* out_scalar, void* workspace,
                        int batch, int height, int width, void* stream);
int ppea_smooth_backward(const float* disp, const float* img, const float* grad_scalar, float* grad_disp,
                         int batch, int height, int width, void* stream);

/* ---- input format: uint8 images expanded on the device --------------------
 * The loss reads the un-augmented ("color", f, s) frames, which the reference produces on the CPU as
 * torchvision ToTensor of a uint8 PIL image (datasets/mono_dataset.py:62, :106): float32(k) / 255,
 * one correctly rounded division.  Shipping the uint8 planes and expanding them here gives the loss
 * bit-identical inputs for a quarter of the host->device bytes.  Any layout: `count` bytes in,
 * `count` floats out, same order. */
int ppea_images_u8_to_f32(const uint8_t* src, float* dst, size_t count, void* stream);

/* ---- plane-sweep cost volume of the multi-frame encoder -------------------
 * `match_features`, networks/replk_matching_adapter.py:261-340 (= replk_matching.py:127-206, resnet_encoder.py:164-246):
 * the lookup features (B,F,C,h,w) are warped into the current frame at every hypothesised depth `depth_bins[d]`
 * (BackprojectDepth + Project3D + F.grid_sample(padding_mode="zeros", align_corners=True)); cost_volume (B,D,h,w) is the
 * channel-mean L1 difference to current_feats (B,C,h,w), masked at the borders, averaged over the lookup frames whose pose
 * is not all-zero; missing_mask (B,D,h,w) flags the bins that never landed inside the image, which get the per-pixel
 * maximum when set_missing_to_max.  K / inv_K are the (B,4,4) intrinsics of the matching scale, relative_poses (B,F,4,4).
 * One launch for the volume plus a small fix-up launch, no workspace, no gradients (the reference runs it under no_grad). */
int ppea_match_features(const float* current_feats, const float* lookup_feats, const float* relative_poses, const float* K,
                        const float* inv_K, const float* depth_bins, float* cost_volume, float* missing_mask, int batch,
                        int num_lookup, int channels, int height, int width, int num_bins, int set_missing_to_max, float eps,
                        void* stream);

/* The same call with a scratch buffer of ppea_match_workspace_bytes(...) bytes (16-byte aligned; 0 bytes: the shape has no
 * fast path, e.g. channels % 4 != 0): both feature tensors are first re-laid as (N, C/4, h, w, 4) -- four channels of a cell in
 * one 16-byte word -- so that a bilinear corner is one 128-bit load for four channels; results are bit-identical to
 * ppea_match_features (same arithmetic, same channel order).  workspace == NULL takes the planar kernel. */
size_t ppea_match_workspace_bytes(int batch, int num_lookup, int channels, int height, int width);
int ppea_match_features_ws(const float* current_feats, const float* lookup_feats, const float* relative_poses, const float* K,
                           const float* inv_K, const float* depth_bins, float* cost_volume, float* missing_mask, int batch,
                           int num_lookup, int channels, int height, int width, int num_bins, int set_missing_to_max, float eps,
                           void* workspace, size_t workspace_bytes, void* stream);

/* `match_features_dyn`, networks/replk_matching_adapter.py:163-258 -- the variant the encoder takes when it is given a teacher
 * depth (:400, :439-442; no caller inside the reference does): the same plane sweep with (a) an occlusion map of the lookup image
 * (occlusion (N,h,w), 1 where sum_c RGB < 0.15 nearest-resized to the matching resolution, :166; indexed by batch item) projected
 * into every layer; where its sample exceeds pool_threshold and aug_mask[b] == 0 (:196) the warped features become 1 (set_1, :202-203)
 * or the 3-D max over the (2 pool_radius + 1)^3 neighbourhood of the un-occluded warped features (pool, :204-209; radius <= 2);
 * (b) cv_min: the lookup frames are combined by minimum (zeros count as 1 before, ones become 0 after, :238-246) instead of averaged.
 * occlusion / aug_mask may be NULL when neither set_1 nor pool is given. */
int ppea_match_features_dyn(const float* current_feats, const float* lookup_feats, const float* relative_poses, const float* K,
                            const float* inv_K, const float* depth_bins, const float* occlusion, const float* aug_mask,
                            float* cost_volume, float* missing_mask, int batch, int num_lookup, int channels, int height, int width,
                            int num_bins, int set_missing_to_max, int cv_min, int set_1, int pool, int pool_radius,
                            float pool_threshold, float eps, void* stream);

/* Tail of the matching block (replk_matching_adapter.py:380-387 compute_confidence_mask, :439-453 in forward), one sweep:
 * confidence (B,h,w) = [#(cost * (1 - missing) > 0 over the bins) == threshold] (missing may be NULL: cost is taken as is);
 * (mins, argmin) (B,h,w) = torch.min over the bins of the volume with exact zeros replaced by 100 (argmin int64, first
 * minimum); with mask_volume the volume is multiplied by the confidence in place.  Any output pointer may be NULL. */
int ppea_match_tail(float* cost_volume, const float* missing_mask_or_null, float* confidence_or_null, float* mins_or_null,
                    long long* argmin_or_null, int batch, int num_bins, int height, int width, int threshold,
                    int mask_volume, void* stream);

/* ---- pose-network output -> camera transform --------------------------------
 * `transformation_from_parameters` (layers.py:26-42 = rot_from_axisangle :62-100 + get_translation_matrix :45-59 + one
 * (B,4,4) matmul), one launch each way: axisangle (B,3), translation (B,3) -> T (B,4,4) row-major; the backward contracts
 * the Jacobian (dual numbers over the same program) with grad_T.  norm() at the origin has derivative 0 as in PyTorch. */
int ppea_pose_to_matrix_forward(const float* axisangle, const float* translation, int invert, float* T, int batch, void* stream);
int ppea_pose_to_matrix_backward(const float* axisangle, const float* translation, int invert, const float* grad_T,
                                 float* grad_axisangle, float* grad_translation, int batch, void* stream);

/* Trainer.compute_matching_mask (trainer.py:859-869): mask[i] = ((1/lowest_cost - mono) / mono < 1) && ((mono - 1/lowest_cost) /
 * (1/lowest_cost) < 1), one byte per element (torch.bool layout); mono_depth (B,1,H,W) and lowest_cost (B,H,W) flattened. */
int ppea_matching_mask(const float* mono_depth, const float* lowest_cost, uint8_t* mask, size_t count, void* stream);

/* ---- glue between the multi-frame encoder and the loss (remainder) -----------------------------
 * ppea_matching_glue: one launch for  outputs["lowest_cost"] = F.interpolate(lowest_cost[:, None], [H, W], "nearest")[:, 0]
 * (networks/repdepth.py:615-617),  outputs["consistency_mask"] = F.interpolate(confidence_mask[:, None], [H, W], "nearest")[:, 0]
 * (:618-620) * Trainer.compute_matching_mask(outputs) (trainer.py:450-451, :859-869), and the per-image minimum / maximum of
 * mono_depth (B,1,H,W) that DepthBins.update reduces (trainer.py:54-55) into minmax_scratch (2*batch 32-bit words; NULL = skip).
 * lowest_cost / confidence are (B,low_h,low_w) float32; outputs (B,H,W) float32. */
int ppea_matching_glue(const float* lowest_cost, const float* confidence, const float* mono_depth, float* lowest_cost_up,
                       float* consistency_mask, void* minmax_scratch, int batch, int low_h, int low_w, int height, int width,
                       void* stream);
/* DepthBins.update (trainer.py:52-64) from the extrema ppea_matching_glue left in minmax_scratch, on the device:
 * min = max(opt_min_depth, mean_b(min_b) * 0.9), max = mean_b(max_b) * 1.1, state = 0.99 state + 0.01 new (both 1-element tensors). */
int ppea_depth_bins_update(const void* minmax_scratch, int batch, float opt_min_depth, float* min_depth_state, float* max_depth_state,
                           void* stream);
/* "set missing images to 0 pose" (networks/repdepth.py:502-505: `if feat.sum() == 0: pose[batch_idx] *= 0`, one host sync per
 * batch item in the reference): pose (B, pose_floats) is zeroed for the items whose pose features (floats_per_item each) are all zero. */
int ppea_zero_missing_poses(const float* pose_feats, size_t floats_per_item, float* pose, int pose_floats, int batch, void* stream);

/* ---- decoder tail: the disparity head -----------------------------------------------------------
 * `self.outputs[("disp", 0)] = self.sigmoid(self.disp_convs[0](x))` (networks/depth_decoder_v2.py:123-129, :239) with
 * Conv3x3 = nn.ReflectionPad2d(1) + nn.Conv2d(C, 1, 3) (layers.py:119-135): x (B,C,H,W), weight (1,C,3,3), bias (1) ->
 * disp (B,1,H,W); with depth_or_null also depth = 1 / (1/max_depth + (1/min_depth - 1/max_depth) disp) (disp_to_depth,
 * layers.py:14-23, what trainer.py:888 evaluates next).  One launch; the padded copy of x never exists.  C <= 256.
 * Backward (any of the three outputs may be NULL; x and the workspace are only needed for grad_weight / grad_bias):
 * grad_disp (B,1,H,W) -> grad_x (B,C,H,W) (adjoint of the reflection padding folded in), grad_weight (1,C,3,3), grad_bias (1);
 * fixed-order reductions (bit-reproducible). */
int ppea_disp_head_forward(const float* x, const float* weight, const float* bias, float* disp, float* depth_or_null, int batch,
                           int channels, int height, int width, float min_depth, float max_depth, void* stream);
size_t ppea_disp_head_workspace_bytes(int batch, int channels, int height, int width);
int ppea_disp_head_backward(const float* x, const float* weight, const float* disp, const float* grad_disp, float* grad_x_or_null,
                            float* grad_weight_or_null, float* grad_bias_or_null, void* workspace, int batch, int channels,
                            int height, int width, void* stream);

/* ---- input format, remainder: the dataset's LANCZOS pyramid and packed RGBx frames on the device ----
 * MonoDataset.preprocess (datasets/mono_dataset.py:96-112) resizes every colour frame with transforms.Resize((h >> i, w >> i),
 * interpolation=Image.LANCZOS) (:79-85), scale i from scale i - 1, on the CPU workers.  PIL's 8-bit resampler is an exact integer
 * algorithm (libImaging/Resample.c: windowed taps, double-precision Lanczos-3 weights normalised and rounded to 22-bit fixed point,
 * int32 accumulation from 2^21, shift, clip; horizontal pass into an 8-bit intermediate, then vertical); these entry points
 * reproduce it bit for bit.
 *   ppea_lanczos_ksize / ppea_lanczos_table   HOST functions: the tap table of one direction (in_size -> out_size):
 *       bounds[2*out_size] = (first tap, tap count) per output coordinate, coeffs[out_size * ksize] fixed-point weights; returns ksize.
 *   ppea_resize_lanczos_u8    src (n_planes,in_h,in_w) -> dst (n_planes,out_h,out_w), uint8 planes on the device; the tables are
 *       device copies of the two host tables; tmp = n_planes*in_h*out_w bytes (needed when both directions change).
 *   ppea_pack_rgbx_u8         planar (N,3,H,W) or interleaved (N,H,W,3) uint8 frames -> (N,H,W) words r | g << 8 | b << 16, the
 *       gather format of the streaming loss kernel. */
int ppea_lanczos_ksize(int in_size, int out_size);
int ppea_lanczos_table(int in_size, int out_size, int* bounds, int* coeffs);
int ppea_resize_lanczos_u8(const uint8_t* src, uint8_t* dst, uint8_t* tmp, size_t n_planes, int in_h, int in_w, int out_h, int out_w,
                           const int* bounds_x, const int* coeffs_x, int ksize_x, const int* bounds_y, const int* coeffs_y, int ksize_y,
                           void* stream);
int ppea_pack_rgbx_u8(const uint8_t* src, uint32_t* dst, size_t n_images, int height, int width, int interleaved, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PPEA_VSL_H_ */
