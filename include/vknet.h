/*
 * vknet.h -- C ABI of libvknet.so: the B200 (sm_100a) implementation of Video K-Net's
 * KernelUpdateHead hot path.  Plain C: raw device pointers, POD structs, an explicit
 * cudaStream_t (passed as void*).  No torch types cross this boundary.
 *
 * Each entry point names the reference code it replaces (paths under lxtGH/Video-K-Net):
 *   knet/det/kernel_update_head.py   KernelUpdateHead.forward        :170-277
 *   knet/kernel_updator.py           KernelUpdator.forward           :56-94
 *   knet/video/kernel_update_head.py VideoKernelUpdateHead.forward   :281-541
 *
 * Conventions
 *   - every function returns 0 on success or a negative VKN_E_* code; vkn_last_error() returns a
 *     thread-local message.  No exceptions cross the boundary.
 *   - the CALLER owns every buffer (inputs, outputs, workspace).  The library never allocates or
 *     frees device memory.  All work is enqueued on `stream`; nothing synchronises the device, so
 *     every call is CUDA-graph capturable.
 *   - layouts are the reference's: x [B,C,H,W] (NCHW, contiguous), mask_preds / new_mask_preds
 *     [B,N,H,W], proposal_feat / obj_feat [B,N,C] (the reference's [B,N,C,1,1]; conv_kernel_size
 *     is 1 in every shipped config and is the only size supported), cls_score [B,N,num_classes].
 *   - x_dtype selects the storage type of x and of the mask tensors (VKN_F32 or VKN_BF16).
 *     Kernel-side tensors (proposal_feat, obj_feat, cls_score, x_feat) are always fp32.
 *   - w_dtype selects the storage type of weight MATRICES (VKN_F32 or VKN_BF16); bias and
 *     LayerNorm vectors are always fp32.  Arithmetic is fp32 throughout (bf16 products are
 *     exact in the fp32 accumulators), so VKN_BF16 changes bytes moved, not the math.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     VKN_E_CUDA.
 */
#ifndef VKNET_H_
#define VKNET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VKN_VERSION 100

enum { VKN_F32 = 0, VKN_BF16 = 1 };

enum {
  VKN_OK = 0,
  VKN_E_INVALID = -1,     /* bad argument / inconsistent shape */
  VKN_E_UNSUPPORTED = -2, /* a configuration outside the shipped path (e.g. conv_kernel_size != 1) */
  VKN_E_CUDA = -3,        /* a CUDA runtime/driver call failed (message has the detail) */
  VKN_E_WORKSPACE = -4    /* workspace too small or misaligned */
};

/* GEMM engine selection for the two big contractions (mask pooling and dynamic mask conv). */
enum {
  VKN_ENGINE_AUTO = 0,  /* tcgen05/TMA when x_dtype == VKN_BF16 and the shape qualifies, else SIMT fp32 */
  VKN_ENGINE_SIMT = 1,  /* CUDA-core fp32 kernels (also the only engine for VKN_F32 storage) */
  VKN_ENGINE_TC = 2     /* force tcgen05 (fails with VKN_E_UNSUPPORTED if the shape does not qualify) */
};

typedef struct VknShape {
  int32_t B;            /* kernel sets in this call (= frames unless frames_per_set > 1) */
  int32_t N;            /* kernels (proposals) per frame: 100 / 117 / 166 in the shipped configs */
  int32_t C;            /* channels (in_channels == feat_channels == out_channels); multiple of 64, <= 256 */
  int32_t H, W;         /* feature-map size; HW arbitrary */
  int32_t ffn_dim;      /* feedforward_channels (2048) */
  int32_t num_classes;  /* fc_cls rows */
  int32_t num_heads;    /* 8 */
  int32_t x_dtype;      /* VKN_F32 | VKN_BF16: storage of x, mask_preds, new_mask_preds */
  int32_t w_dtype;      /* VKN_F32 | VKN_BF16: storage of weight matrices */
  int32_t with_ffn;     /* KernelUpdateHead(with_ffn=...) */
  int32_t engine;       /* VKN_ENGINE_* */
  int32_t frames_per_set; /* F (0 or 1 = one frame per kernel set).  F > 1 is the clip head
                           * KernelUpdateHeadVideo (knet_vis/tracker/kernel_update_head.py:209-374, gathered
                           * mode): x / mask tensors hold B*F frames, ONE kernel set per clip pools the MEAN
                           * over its F frames (:245-246) and convolves all of them (:329-339). */
  float mask_thr_logit; /* logit(hard_mask_thr): sigmoid(m) > thr  <=>  m > mask_thr_logit (0 for thr=0.5) */
} VknShape;

/* KernelUpdator parameters (knet/kernel_updator.py:36-54).  *_w are [out,in] row-major like nn.Linear. */
typedef struct VknUpdatorW {
  const void *dyn_w;  const float *dyn_b;   /* dynamic_layer  [2C,C] */
  const void *inp_w;  const float *inp_b;   /* input_layer    [2C,C] */
  const void *ig_w;   const float *ig_b;    /* input_gate     [C,C]  */
  const void *ug_w;   const float *ug_b;    /* update_gate    [C,C]  */
  const float *norm_in_g, *norm_in_b;       /* norm_in        (update gate LN) */
  const float *norm_out_g, *norm_out_b;     /* norm_out       (param_out LN)   */
  const float *inorm_in_g, *inorm_in_b;     /* input_norm_in  (input gate LN)  */
  const float *inorm_out_g, *inorm_out_b;   /* input_norm_out (input_out LN)   */
  const void *fc_w;   const float *fc_b;    /* fc_layer       [C,C]  */
  const float *fc_norm_g, *fc_norm_b;       /* fc_norm */
} VknUpdatorW;

/* mmcv MultiheadAttention (nn.MultiheadAttention packed in_proj) + the LayerNorm that follows it. */
typedef struct VknAttnW {
  const void *in_w;  const float *in_b;     /* in_proj  [3C,C] */
  const void *out_w; const float *out_b;    /* out_proj [C,C]  */
  const float *norm_g, *norm_b;
} VknAttnW;

/* mmcv FFN(num_fcs=2) + the LayerNorm that follows it. */
typedef struct VknFfnW {
  const void *w1; const float *b1;          /* layers.0.0 [F,C] */
  const void *w2; const float *b2;          /* layers.1   [C,F] */
  const float *norm_g, *norm_b;
} VknFfnW;

#define VKN_MAX_FCS 4

/* One KernelUpdateHead stage (state_dict contract: SURVEY.md Appendix C). */
typedef struct VknHeadW {
  const void *ft_w;        /* feat_transform.conv.weight [C,C] (identity when feat_transform_cfg is None) */
  const float *ft_b;       /* feat_transform.conv.bias   [C]   */
  const void *ft_wt_ext;   /* [C+1,C]: rows 0..C-1 = ft_w^T, row C = ft_b  (host-prepared fold operand) */
  VknUpdatorW upd;         /* kernel_update_conv.* */
  VknAttnW attn;           /* attention.attn.*, attention_norm.* */
  VknFfnW ffn;             /* ffn.*, ffn_norm.* (ignored when with_ffn == 0) */
  int32_t num_cls_fcs, num_mask_fcs;
  const void *cls_fc_w[VKN_MAX_FCS];  const float *cls_ln_g[VKN_MAX_FCS], *cls_ln_b[VKN_MAX_FCS];
  const void *fc_cls_w;  const float *fc_cls_b;     /* [num_classes,C]; NULL = head built with with_cls=False */
  const void *mask_fc_w[VKN_MAX_FCS]; const float *mask_ln_g[VKN_MAX_FCS], *mask_ln_b[VKN_MAX_FCS];
  const void *fc_mask_w; const float *fc_mask_b;    /* [C,C] */
  const void *fc_pack;     /* optional host-prepared operand of the single-frame row engine (vkn_frame_chain_pack): the weight
                            * matrices re-laid as the 32-row chunks each CTA of a cluster streams, one bulk copy per chunk.
                            * NULL: the engine copies the chunks row by row from the matrices above (slower, same results). */
} VknHeadW;

/* One cross-frame link block of VideoKernelUpdateHead (knet/video/kernel_update_head.py:167-260):
 * optional KernelUpdator on the previous kernels, cross attention + LN, FFN + LN. */
typedef struct VknLinkW {
  int32_t has_updator;     /* 1: prev' = KernelUpdator(x_feat, prev) first ('update', 'update_dynamic_cov') */
  VknUpdatorW upd;
  VknAttnW attn;
  VknFfnW ffn;
} VknLinkW;

int vkn_version(void);
const char *vkn_last_error(void);

/* Names of the device kernels this library launches, '\n' separated (for the bench's launch audit). */
const char *vkn_kernel_names(void);

/* Kernel launches issued by this library since it was loaded (graph replays are not counted: multiply
 * the launches recorded during capture by the number of replays). */
unsigned long long vkn_launch_count(void);

/* Live per-kernel timing.  Between vkn_profile_begin() and vkn_profile_end() every launch is
 * bracketed by CUDA events on the launching stream; _end synchronises on the last event and returns,
 * in launch order, each kernel's name and device time in ms (time to the next launch on the stream).
 * Not capturable in a CUDA graph; meant for bench.py's roofline section, not for production calls. */
int vkn_profile_begin(void);
int vkn_profile_end(const char **names, float *ms, int max_entries, int *count);

/* Debug only: when `buf` (device memory, n_u64 zero-initialised 64-bit words) is set, every following row-operator
 * launch writes per-CTA phase timestamps (%globaltimer, ns) into its own block of the buffer: block i = launch i,
 * 8 words per CTA (entry, prefetch issued, dependency resolved, panel built, tile visible, main loop done,
 * stores issued).  Returns the block stride in words.  Pass NULL to switch it off.  tools/linear_timeline.py */
int vkn_debug_timestamps(unsigned long long *buf, size_t n_u64);

/* Single-frame row engine (csrc/framechain.cu; taken by the stage / iter entry points below ~400 kernel rows when C = 256,
 * 8 heads, ffn_dim = 2048, bf16 weights, one cls / mask FC): size of, and the one-time fill of, VknHeadW.fc_pack for the weights
 * `w` points at (only the weight-related fields of `shape` matter).  *bytes = 0 when the engine does not apply to this head.
 * Re-run after the weights change.  `out` must be 16-byte aligned device memory. */
int vkn_frame_chain_pack_bytes(const VknShape *shape, const VknHeadW *w, size_t *bytes);
int vkn_frame_chain_pack(const VknShape *shape, const VknHeadW *w, void *out, size_t bytes, void *stream);

/* Row f4 (training side): the cost matrix of MaskHungarianAssigner.assign for one image
 * (knet/det/mask_hungarian_assigner.py:228-247 with DiceCost :43-75, MaskCost :93-110, mmdet FocalLossCost; pred_act=True,
 * act_mode='sigmoid'):  cost [N,M] fp32 = w_cls * focal + w_mask * mask + w_dice * dice  (a zero weight skips the term).
 *   mask_logits [N,HW] fp32, cls_logits [N,ncls] fp32 or NULL, gt_masks [M,HW] fp32, gt_labels [M] int64.
 *   params [7] (host): w_cls, w_mask, w_dice, dice_eps (1e-3), focal_alpha (0.25), focal_gamma (2), focal_eps (1e-12).
 * One pass over the masks (the reference materialises two activations and runs three einsums); deterministic.  The Hungarian
 * solve on the cost matrix stays the caller's (scipy.optimize.linear_sum_assignment in the reference, :246-251). */
int vkn_match_cost_workspace_bytes(int N, int M, int HW, size_t *bytes);
int vkn_match_cost(const float *mask_logits, const float *cls_logits, const float *gt_masks, const int64_t *gt_labels, int N,
                   int M, int HW, int ncls, const float *params, float *cost, void *workspace, size_t workspace_bytes,
                   void *stream);

/* Bytes of caller-provided scratch needed by the stage / link / iter entry points for `shape`. */
int vkn_workspace_bytes(const VknShape *shape, size_t *bytes);

/* ---- individual operators (each is also a step of vkn_stage_forward) ------------------------- */

/* a3+a4 with the feat_transform folded out: knet/det/kernel_update_head.py:179-195.
 *   x_feat[b,n,:] = sum_p 1[mask[b,n,p] > thr] * (ft_w x[b,:,p] + ft_b)      -> fp32 [B,N,C]  */
int vkn_mask_pool(const VknShape *s, const VknHeadW *w, const void *x, const void *mask_preds,
                  float *x_feat, void *workspace, size_t workspace_bytes, void *stream);

/* a5: KernelUpdator.forward, knet/kernel_updator.py:56-94.  update_feature = x_feat [P,C],
 * input_feature = proposal_feat [P,C] (K=1), out [P,C];  P = B*N rows. */
int vkn_kernel_update(const VknShape *s, const VknUpdatorW *w, const float *x_feat,
                      const float *proposal_feat, float *out, void *workspace, size_t workspace_bytes,
                      void *stream);

/* a6: LN(q_in + MHA(q_in, kv_in, kv_in)), attention across the N kernels of each frame
 * (knet/det/kernel_update_head.py:204-208; cross form: knet/video/kernel_update_head.py:337-345). */
int vkn_mhsa_ln(const VknShape *s, const VknAttnW *w, const float *q_in, const float *kv_in, float *out,
                void *workspace, size_t workspace_bytes, void *stream);

/* a7: LN(in + W2 relu(W1 in + b1) + b2), knet/det/kernel_update_head.py:214-215. */
int vkn_ffn_ln(const VknShape *s, const VknFfnW *w, const float *in, float *out, void *workspace,
               size_t workspace_bytes, void *stream);

/* a8 + the fold operand for a9: cls/mask FC stacks (knet/det/kernel_update_head.py:217-227).
 *   cls_score [P,num_classes]; mask_kernel [P,C] = fc_mask(...) (the reference's mask_feat). */
int vkn_heads(const VknShape *s, const VknHeadW *w, const float *obj_feat, float *cls_score,
              float *mask_kernel, void *workspace, size_t workspace_bytes, void *stream);

/* a9: dynamic 1x1 mask convolution with feat_transform folded in
 * (knet/det/kernel_update_head.py:179-180 + :247-260):
 *   new_mask[b,n,p] = sum_c mask_kernel[b,n,c] * (ft_w x[b,:,p] + ft_b)[c]                         */
int vkn_mask_gemm(const VknShape *s, const VknHeadW *w, const void *x, const float *mask_kernel,
                  void *new_mask_preds, void *workspace, size_t workspace_bytes, void *stream);

/* ---- whole stage / loop ----------------------------------------------------------------------- */

/* KernelUpdateHead.forward (knet/det/kernel_update_head.py:170-277) for conv_kernel_size == 1.
 * Outputs: cls_score [B,N,num_classes] fp32, new_mask_preds [B,N,H,W] x_dtype, obj_feat [B,N,C] fp32,
 * x_feat_out [B,N,C] fp32 or NULL (the pooled feature VideoKernelUpdateHead also returns).
 * x_feat_in, when not NULL, is a pooled feature already computed by vkn_mask_pool for these
 * mask_preds (the video head pools once, runs its link block, then the stage): pooling is skipped
 * and mask_preds may be NULL.  new_mask_preds may be NULL (skip a9: callers that only need kernels). */
int vkn_stage_forward(const VknShape *s, const VknHeadW *w, const void *x, const float *proposal_feat,
                      const void *mask_preds, const float *x_feat_in, float *cls_score, void *new_mask_preds,
                      float *obj_feat, float *x_feat_out, void *workspace, size_t workspace_bytes,
                      void *stream);

/* The S-stage loop of KernelIterHead.simple_test (knet/det/kernel_iter_head.py:246-253):
 * stage s consumes stage s-1's obj_feat and mask logits.  Only the last stage's outputs are
 * returned (that is all simple_test reads); intermediate masks live in the workspace. */
int vkn_iter_forward(const VknShape *s, const VknHeadW *stages, int num_stages, const void *x,
                     const float *proposal_feat, const void *mask_preds, float *cls_score,
                     void *new_mask_preds, float *obj_feat, void *workspace, size_t workspace_bytes,
                     void *stream);

/* Cross-frame link block of VideoKernelUpdateHead (knet/video/kernel_update_head.py:324-348,
 * :394-415, :417-444):  prev' = has_updator ? KernelUpdator(x_feat, prev) : prev;
 *   t = LN(cur + MHA(q=cur, k=v=prev'));  out = LN(FFN(t)).   cur, prev, out: [B,N,C] fp32. */
int vkn_link_attend(const VknShape *s, const VknLinkW *w, const float *cur, const float *prev,
                    const float *x_feat, float *out, void *workspace, size_t workspace_bytes, void *stream);

/* ---- next row: the tail of ConvKernelHead._decode_init_proposals ----------------------
 * knet/det/kernel_head.py:212, 234-254 with the shipped settings proposal_feats_with_obj=True, use_binary=True:
 *   mask_preds[b,n,p]     = init_w[n,:] . loc_feats[b,:,p] + init_b[n]              (init_kernels, a 1x1 conv)
 *   obj[b,n,:]            = sum_p 1[sigmoid(mask_preds[b,n,p]) > 0.5] * x_feats[b,:,p]
 *   proposal_feats[b,n,:] = init_w[n,:] + obj[b,n,:]
 * The same two tcgen05 kernels as the stages, with ONE static kernel set shared by all B frames.
 * shape: B frames, N = num_proposals, C, H, W, x_dtype (loc_feats / x_feats / mask_preds storage); the row-operator
 * fields (ffn_dim, num_classes, num_heads) are ignored but must be valid.  init_w [N,C] fp32, init_b [N] fp32 or NULL. */
int vkn_init_proposals(const VknShape *s, const float *init_w, const float *init_b, const void *loc_feats,
                       const void *x_feats, void *mask_preds, float *proposal_feats, void *workspace,
                       size_t workspace_bytes, void *stream);

/* ---- next row: the post-loop mask path -------------------------------------------------
 * One launch for: the last-stage bilinear x`mask_upsample_stride` of _mask_forward (knet/det/kernel_iter_head.py:122-128),
 * KernelUpdateHead.rescale_masks (knet/det/kernel_update_head.py:443-458: sigmoid -> bilinear to batch_input_shape ->
 * crop [:img_h,:img_w] -> bilinear to ori_shape) and the `> mask_thr` of get_seg_masks (:460-462).
 *   masks   [K, H, W] logits of the K selected kernels (dtype = VKN_F32 / VKN_BF16), device memory
 *   up      mask_upsample_stride (1 = the masks are already the scaled_mask_preds)
 *   probs   [K, ori_h, ori_w] fp32 or NULL;  bits [K, ori_h, ori_w] uint8 (1 where prob > mask_thr) or NULL
 * All interpolations are torch's bilinear, align_corners=False.  VKN_E_UNSUPPORTED for resize ratios whose per-tile
 * dependency cone does not fit shared memory (strong down-scaling). */
int vkn_rescale_masks(const void *masks, int dtype, int K, int H, int W, int up, int batch_h, int batch_w, int img_h,
                      int img_w, int ori_h, int ori_w, float mask_thr, float *probs, unsigned char *bits, void *stream);

/* Joint panoptic merge (knet/video/kernel_iter_head.py:832-895, merge_stuff_thing_stuff_joint; the shipped video configs
 * set merge_joint=True): the pixel goes to the kernel with the highest score x probability; kernels are then visited in
 * descending score order and kept when (thing: score >= instance_score_thr) and won_area > 0, area(prob >= 0.5) > 0 and
 * won_area / area(prob >= 0.5) >= overlap_thr.  No host synchronisation: three launches on `stream`.
 *   masks   [T, H, W] fp32 probabilities (things first, then stuff -- the reference's torch.cat order), device memory
 *   scores  [T] fp32, labels [T] int32 (label < num_thing_classes = thing)
 *   panoptic_seg  [H, W] int32: segment id per pixel (0 = none)
 *   segments      [T, 5] int32 rows (id, isthing, category_id, instance_id | -1, area | -1) for the first counts[0] rows
 *                 (category_id of stuff = label - num_thing_classes + 1, as the reference writes it)
 *   segment_scores [T] fp32 (score of each segment row), kept_things [T] int32 (indices of kept thing kernels, in order),
 *   counts [2] int32 = {number of segments, number of kept things}
 *   workspace: (H*W + 3*T) * 4 bytes. */
int vkn_panoptic_merge(const float *masks, const float *scores, const int32_t *labels, int num_kernels, int H, int W,
                       int num_thing_classes, double instance_score_thr, double overlap_thr, int32_t *panoptic_seg,
                       int32_t *segments, float *segment_scores, int32_t *kept_things, int32_t *counts, void *workspace,
                       size_t workspace_bytes, void *stream);

/* ---- next row: tracking embeddings + association ----------------------------------------------
 * A stack of nn.Linear layers over [rows, features] fp32: y = W x + b, then (optional) LayerNorm, then (optional) ReLU.
 * Covers the tracking-embedding path of the video detectors on the last-stage kernels:
 *   embed_fcs (Linear(no bias) -> LN -> ReLU) + fc_embed     knet/video/knet_quansi_dense_embed_fc_joint_train.py:113-126, 572-580
 *   QuasiDenseMaskEmbedHeadGTMask.forward (fcs: Linear -> ReLU, fc_embed)   knet/video/track_heads.py:632-642
 * One fused launch per Linear (the LayerNorm / ReLU of layer i is the prologue of layer i+1).  w [out_dim, in_dim] in
 * w_dtype, b / ln_g / ln_b fp32 or NULL.  workspace: 256-byte aligned, 2 * round_up(rows * max(out_dim) * 4, 256) bytes. */
typedef struct VknMlpLayer {
  const void *w; const float *b;
  const float *ln_g, *ln_b;      /* LayerNorm(out_dim) after the Linear, or both NULL */
  int32_t in_dim, out_dim;
  int32_t relu;                  /* ReLU after the (normalised) output */
} VknMlpLayer;
int vkn_mlp(const VknMlpLayer *layers, int num_layers, int w_dtype, const float *in, float *out, int rows, void *workspace,
            size_t workspace_bytes, void *stream);

/* QuasiDenseEmbedTracker.match up to the memory update (knet/video/qdtrack/trackers/quasi_dense_embed_tracker.py:137-204,
 * match_metric='bisoftmax'): sort by score, duplicate removal by box IoU, bi-directional softmax of embeds . memo_embeds^T,
 * same-category mask, greedy assignment with column knock-out, ids of new tracks.  ONE single-CTA launch.
 *   bboxes [n,5] fp32 (x1,y1,x2,y2,score), labels [n] int64, embeds [n,D] fp32; memory: memo_labels [m] int64,
 *   memo_embeds [m,D] fp32, memo_ids [m] int64 (-1 = backdrop); thresholds[6] = {obj_score_thr, match_score_thr,
 *   init_score_thr, nms_conf_thr, nms_backdrop_iou_thr, nms_class_iou_thr}; num_tracklets = the tracker's id counter.
 *   selected [n] int32: indices of the kept detections in score order (the rows of the reference's returned bboxes);
 *   ids [n] int64: track id per kept detection (-1 none, -2 suppressed duplicate); counts[2] = {kept, new tracks}.
 *   workspace: 2 * n * m * 4 bytes (at least 8). */
int vkn_track_match(const float *bboxes, const int64_t *labels, const float *embeds, int n, int embed_dim,
                    const int64_t *memo_labels, const float *memo_embeds, const int64_t *memo_ids, int m,
                    const float *thresholds, int with_cats, int64_t num_tracklets, int32_t *selected, int64_t *ids,
                    int32_t *counts, void *workspace, size_t workspace_bytes, void *stream);

/* Mask -> box reduction of VideoKernelUpdateHead.segm2result (knet/video/kernel_update_head.py:734-744; unitrack
 * tensor_mask2box): boxes[k] = (x_min, y_min, x_max, y_max) of the non-zero pixels of mask k, (-1, -1, 10, 10) when the
 * mask is empty (the caller clips at 0 like the reference).  masks [K, H, W], elem_bytes 1 (bool / uint8) or 4 (float32). */
int vkn_mask_boxes(const void *masks, int elem_bytes, int K, int H, int W, float *boxes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* VKNET_H_ */
