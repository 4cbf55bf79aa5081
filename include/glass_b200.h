/*
 * glass_b200.h -- C ABI of libglass_b200.so: the B200 (sm_100a) kernels behind GLASS's
 * per-image dense forward path.
 *
 * The reference (amazon-science/glass-text-spotting) contains no native code and no FFI:
 * every native operator on its hot path lives in detectron2 v0.6's `_C` extension or in
 * PyTorch/cuDNN/cuBLAS and is reached through Python.  Each entry point below therefore
 * cites the reference *call site* (file:line under /root/reference) whose arithmetic it
 * replaces; INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers + sizes only; all pointers are DEVICE pointers unless named host_*.
 *   - every op enqueues on `stream` (a cudaStream_t passed as void*), never synchronises,
 *     never allocates device memory; the caller owns every buffer incl. workspaces.
 *   - return 0 on success, negative on error; glass_last_error() gives the message
 *     (thread-local).  Shapes/alignments are validated on the host before launch.
 *
 * Activation storage ("split-fp16 padded NHWC"): an fp32 tensor [N,C,H,W] is held as two
 * fp16 planes hi, lo (16*x ~= hi + lo, 22 mantissa bits; the power-of-two
 * pre-scale keeps lo out of the fp16 subnormals) each laid out [N, H+2b, W+2b, Cp]
 * (b = border of zero pixels, Cp = C rounded up to 64, pad channels zero).  The zero
 * border is the conv padding; kernels never write it.
 */
#ifndef GLASS_B200_H_
#define GLASS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Shared-border planes: OR this flag into any `border` argument.  A plane [h, w] is then laid out [h + b, w + b] with
 * only LEADING zero rows / columns: in the flattened pixel order the right neighbour of a row's last pixel is the next
 * row's leading zero column, and the row below a plane's last row is the next plane's leading zero row (the caller keeps
 * one more zero plane after the last).  A 16 x 33 map then costs 17 x 34 = 578 GEMM rows instead of 18 x 35 = 630. */
#define GLASS_BORDER_SHARED 0x100
#define GLASS_MAX_TAPS 16
#define GLASS_MAX_LEVELS 5

const char* glass_last_error(void);
int glass_abi_version(void);
/* number of kernels launched by this library in this process so far (bench "gpu_launches") */
int64_t glass_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * glass_conv_gemm -- the tcgen05 implicit-GEMM used by every conv / Linear on the path.
 *   D[m, n] = sum_t sum_k A[m + tap_shift[t], k] * W[n, t*k_per_tap + k]
 *   y = D*scale[n] + bias[n];  (relu_pre) y = max(y,0);  y += residual;  (relu_post) y = max(y,0)
 * replaces: torch.nn.Conv2d / detectron2 Conv2d(+norm+act) / nn.Linear as used by
 *   d2 ResNet+FPN (configs/glass_pretrain.yaml:41-54), StandardRPNHead (rotated_rpn.py:17),
 *   FastRCNNConvFCHead (recognizers_hybrid_head.py:321), P2P3Fusion (fusion_modules.py:281-286),
 *   ResNetFeatureExtractor (local_feature_extraction.py:154-188), MultiAspectGCAttention.out
 *   (fusion_modules.py:157), CNN_V1_1 (recognizer_backbone.py:77-81), BiLSTM input projections
 *   (recognizer_encoder.py:141-143), AttentionUnit.xEmbed (prediction_aster.py:250).
 * A rows are pixels of a padded NHWC tensor flattened over (img, y, x); a conv tap is a constant
 * row shift.  mode: 0 = fp16x3 split (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM; fp32-grade),
 * 1 = single fp16 pass (fast, 11-bit operands).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  /* A operand */
  const void* a_hi;  /* fp16 [rows_a, k_per_tap] */
  const void* a_lo;  /* fp16, same shape (unused when mode == 1) */
  int64_t rows_a;
  int32_t a_ld;      /* element stride between consecutive A rows; 0 = k_per_tap.  a_ld < k_per_tap (8, 16 or 32
                        with k_per_tap = 64) is the compact-channel mode: one 64-wide k-block then spans 64/a_ld
                        consecutive pixels of a narrow NHWC tensor (the s-taps of a 3x3 conv are contiguous) */
  int32_t k_per_tap; /* multiple of 64 */
  int32_t ntaps;     /* 1..GLASS_MAX_TAPS */
  int32_t tap_shift[GLASS_MAX_TAPS];
  /* B operand: packed weights [n, ntaps*k_per_tap], fp16 hi/lo */
  const void* b_hi;
  const void* b_lo;
  int32_t n;         /* multiple of 16; if > 256 a multiple of 128 */
  int32_t mode;
  /* M space: rows m in [0, m_imgs*m_h*m_w) decode to (img, yy, xx); rows whose yy/xx fall in the
   * m_border ring are computed but not stored */
  int32_t m_imgs, m_h, m_w, m_border;
  /* epilogue */
  const float* scale; /* [n] or NULL (=1) */
  const float* bias;  /* [n] or NULL (=0) */
  int32_t relu_pre, relu_post;
  /* residual (optional): split fp16 planes [m_imgs, res_hp, res_wp, n]; pixel (y>>res_shift, x>>res_shift) */
  const void* res_hi;
  const void* res_lo;
  int32_t res_hp, res_wp, res_border, res_shift;
  /* outputs (any subset): split fp16 planes and/or fp32, rows laid out [m_imgs, out_hp, out_wp, ld] */
  void* out_hi;
  void* out_lo;
  float* out_f32;
  int32_t out_hp, out_wp, out_border;
  int32_t ld_out;     /* channel stride of out_hi/out_lo/res rows (>= n) */
  int32_t ld_f32;     /* channel stride of out_f32 rows (>= n) */
  int32_t n_store;    /* columns actually stored (<= n); 0 = n */
  int32_t kb_per_chunk; /* 64-wide k-blocks summed inside the tensor core between drains; 0 = default (2 in split
                           mode: relL2 4.6e-7 vs fp64, 1.5x an fp32 CPU GEMM; 1 = 2.6e-7; larger = faster, error grows
                           with the chain) */
  int32_t pair_mode;    /* 0 = auto, 1 = one CTA per tile, 2 = CTA pairs (tcgen05 cta_group::2, shared weight tile) */
  int32_t tap_mode;     /* 0 = auto (3x3 taps ordered (r,s) with consecutive s-rows share one staged activation block),
                           1 = every tap loads its own 128-row tile */
  /* Pixel-grouped mode (a_inner > 0) for narrow activations: ONE GEMM row = P consecutive pixels (row stride
   * a_ld = P * channels), its K window = the (P + kw - 1) pixels the group's outputs read (a_inner >= a_col0 +
   * k_per_tap elements of the overlapping-row tensor), the weights are the block-Toeplitz matrix [P * Cout,
   * taps * k_per_tap], and the output row is P * Cout contiguous elements = the NHWC rows of the P pixels.
   * P-fold fewer (128-byte) TMA rows per pixel than the compact mode.  The caller passes an all-valid M space
   * (m_border 0) and re-zeroes the activation border afterwards (glass_zero_border). */
  int32_t a_col0;       /* first column of every k-block box inside the tensor row (multiple of 8) */
  int32_t a_inner;      /* inner extent of the tensor map in elements; 0 = k_per_tap (not grouped) */
  /* Live M extent on the device (optional): only rows m < *m_count_dev * m_rows_per_count are computed.  The host sizes
   * the launch for the worst case (m_imgs = capacity); the count -- e.g. the number of detected words, written by an
   * earlier kernel of the same stream -- is read by the persistent grid itself, so no host round trip sizes the GEMM. */
  const int32_t* m_count_dev;
  int32_t m_rows_per_count;
  /* Epilogue: 0 = auto (TMA stores + TMA-loaded residual whenever the layer is "flat": split-fp16 output whose padded
   * plane is the M space itself, n a multiple of 64 per tile, residual -- if any -- of the same geometry; direct stores
   * otherwise), 1 = direct stores, 2 = TMA (error if not eligible).  Both epilogues produce the same bits. */
  int32_t epi_mode;
  /* Debug (optional, NULL in production): device counter incremented by the number of stored outputs whose storage value
   * 16*y lies beyond +-60000, i.e. that SATURATE in the split-fp16 format (|y| > 3750).  BatchNorm-normalised networks stay
   * far inside; an uncalibrated checkpoint may not -- this makes the otherwise silent clamp visible. */
  int32_t* sat_count;
} GlassConvGemmParams;
int glass_conv_gemm(const GlassConvGemmParams* p, void* stream);
/* Launch plans (SURVEY.md 8b "cached in an opaque handle the caller owns"): glass_plan_create does everything the host
 * derives from the parameters -- validation, the four TMA descriptors, tile / pipeline geometry, grid -- once;
 * glass_plan_launch then costs one cudaLaunchKernelEx.  A plan stays valid as long as the buffers it names do (the
 * library never owns device memory).  glass_conv_gemm(p, s) == create + launch + destroy. */
int glass_plan_create(const GlassConvGemmParams* p, void** plan_out);
int glass_plan_launch(const void* plan, void* stream);
int glass_plan_destroy(void* plan);
/* glass_prepack_weights -- weights into the layout glass_conv_gemm consumes, once, into caller-owned buffers:
 * w fp32 [n, k] (K = tap-major, channel-minor, already zero padded to the k the GEMM will be given) -> hi / lo fp16 [n, k]
 * with a per-row power-of-two pre-scale (the row's largest magnitude lands in [256, 512), so both planes stay in fp16's
 * normal range); scale fp32 [n] holds the folded epilogue scale (e.g. the BatchNorm scale) on entry and is multiplied by
 * 2^-s / 16 (the inverse of the weight and activation pre-scales) on exit. */
int glass_prepack_weights(const float* w, int n, int k, void* hi, void* lo, float* scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * Layout / glue kernels.  `n_dev` (where present, optional): device pointer to the LIVE image / word count (<= n, which
 * then is the capacity the launch is sized for) -- the count of detected words stays on the device, no host sync.
 * ------------------------------------------------------------------------------------------ */
/* fp32 NCHW [n,c,h,w] -> split-fp16 padded NHWC [n,h+2b,w+2b,cp] (interior only; border/pad channels must be 0) */
int glass_pack_nchw(const float* src, int n, int c, int h, int w, void* dst_hi, void* dst_lo, int cp, int border,
                    void* stream);
/* split-fp16 padded NHWC -> fp32 NCHW (hi + lo) */
int glass_unpack_nchw(const void* src_hi, const void* src_lo, int n, int c, int h, int w, int cp, int border,
                      float* dst, void* stream);
/* fp32 padded NHWC [n,h+2b,w+2b,ld] -> fp32 NCHW [n,c,h,w] */
int glass_nhwc_f32_to_nchw(const float* src, int n, int c, int h, int w, int ld, int border, float* dst,
                           void* stream);

/* Stem pre-pass with fused (x - mean)/std (d2 GeneralizedRCNN.preprocess_image, called at glass_rcnn.py:82), no im2col
 * matrix: raw fp32 NCHW [n,3,h,w] -> normalised space-to-depth map, split-fp16 NHWC
 * [n, h/2+2b, w/2+2b, 16] with a b-pixel zero border, b = 1 or 2 (pixel (Y,X), channel (dy*2+dx)*3+c = (img[c,2Y+dy,2X+dx]-mean)*inv_std;
 * channels 12..15 zero).  BasicStem's 7x7/s2/p3 conv (d2 resnet.py via configs/glass_pretrain.yaml:41-50) then runs as a
 * 4x4 stride-1 conv in glass_conv_gemm's compact-channel mode (a_ld = 16, 4 taps of one 64-wide k-block,
 * packing.pack_stem_s2d).  The destination's border must be zero (allocate it zeroed once; only the interior is written).
 * b = 1 suffices for the 2-pixel reach of the taps (flattened order: the cell left of a row's left border is the previous
 * row's right border, the row above the top border is the previous image's bottom border) and gives the map the geometry
 * of the conv's output plane, so that the GEMM is flat (TMA-store epilogue). */
int glass_stem_s2d(const float* img, int n, int h, int w, const float* mean, const float* inv_std, void* dst_hi,
                   void* dst_lo, int border, void* stream);

/* Generic tap gather (im2col) for strided / odd-shaped convs:
 * src split padded NHWC [n,h+2b,w+2b,cp] -> rows of kh*kw*cp elements, tap-major K.  dst_border = 0: dense rows
 * [n*ho*wo]; otherwise a border code (GLASS_BORDER_SHARED allowed) of the OUTPUT plane: the row of output pixel (y, x)
 * sits at that plane's flattened position, border rows are left untouched -- the GEMM's M space is then the output
 * plane itself and the layer finishes through the TMA-store epilogue. */
int glass_gather_taps(const void* src_hi, const void* src_lo, int n, int h, int w, int cp, int border, int kh,
                      int kw, int sh, int sw, int ph, int pw, int ho, int wo, void* dst_hi, void* dst_lo,
                      int dst_border, const int32_t* n_dev, void* stream);

/* max_pool2d on split-fp16 padded NHWC (F.max_pool2d: BasicStem, local_feature_extraction.py:163-178). */
int glass_maxpool(const void* src_hi, const void* src_lo, int n, int h, int w, int cp, int border, int kh, int kw,
                  int sh, int sw, int ph, int pw, int ho, int wo, void* dst_hi, void* dst_lo, int dst_border,
                  const int32_t* n_dev, void* stream);

/* Re-zero the 1-pixel border of a split-fp16 padded NHWC activation [n, h+2, w+2, cp] (after a pixel-grouped
 * glass_conv_gemm, which writes every pixel of the padded plane). */
int glass_zero_border(void* hi, void* lo, int n, int h, int w, int cp, const int32_t* n_dev, void* stream);

/* Pre-processing of GlassRunner._image_to_tensor (glass/inference/glass_runner.py:123-148): uint8 HWC image ->
 * fp32 CHW tensor, bilinear resize with align_corners=False semantics of torch.nn.functional.interpolate(size=...),
 * optional channel reversal (RGB input_format, glass_runner.py:83-85).  ho == h && wo == w is a plain convert. */
int glass_resize_bilinear_u8(const uint8_t* src_hwc, int h, int w, int flip_channels, float* dst_chw, int ho, int wo,
                             void* stream);

/* ------------------------------------------------------------------------------------------
 * glass_roi_align_rotated -- multi-level rotated RoIAlign (detectron2 ROIPooler + ROIAlignRotated).
 * replaces: torch.ops.detectron2.roi_align_rotated_forward as called through ROIPooler at
 *   recognizers_hybrid_head.py:320 (box pooler 7x7, 5 levels, sampling 2), :550 (recognizer pooler
 *   8x32, 1 level, adaptive sampling) and :556 (image pooler 128x128, scale 1, sampling 2).
 * Feature maps are padded NHWC [n, h+2b, w+2b, ld]: either fp32 (the conv kernel's out_f32) or the
 * split-fp16 hi/lo planes of an activation (feat_is_split; value = hi + lo, same bytes per element).
 * rois: fp32 [n_rois, 6] = (batch_idx, cx, cy, w, h, angle_deg).  With num_levels > 1 the level of
 * each RoI follows d2's assign_boxes_to_levels (canonical size 224 at level 4).
 * Outputs (any subset): out_f32 [n_rois, ph, pw, c] (NHWC order) and split fp16 rows written at
 *   row ((roi*out_hp + y + out_border)*out_wp + x + out_border), channel offset out_coff, stride ld_out.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t num_levels;
  const void* feat[GLASS_MAX_LEVELS];    /* fp32 map, or the fp16 hi plane when feat_is_split */
  const void* feat_lo[GLASS_MAX_LEVELS]; /* fp16 lo plane (feat_is_split only) */
  int32_t feat_is_split;
  int32_t feat_h[GLASS_MAX_LEVELS], feat_w[GLASS_MAX_LEVELS];
  float spatial_scale[GLASS_MAX_LEVELS];
  int32_t feat_border, feat_ld, channels;
  int32_t min_level; /* level index of feat[0] (2 for p2) */
  const float* rois;
  const int32_t* n_rois_dev; /* optional device count (<= n_rois); NULL = n_rois */
  int32_t n_rois;
  int32_t pooled_h, pooled_w, sampling_ratio;
  float* out_f32;
  void* out_hi;
  void* out_lo;
  int32_t out_hp, out_wp, out_border, out_coff, ld_out;
} GlassRoiAlignParams;
int glass_roi_align_rotated(const GlassRoiAlignParams* p, void* stream);

/* Image pooler variant (recognizers_hybrid_head.py:556): raw fp32 NCHW image [n,3,h,w], normalised on
 * the fly with (x-mean)*inv_std, padded to (hp_img, wp_img) with zeros (ImageList.from_tensors). */
typedef struct {
  const float* img;
  int32_t n, h, w, h_pad, w_pad;
  float mean[3], inv_std[3];
  const float* rois;
  const int32_t* n_rois_dev;
  int32_t n_rois, pooled_h, pooled_w, sampling_ratio;
  float* out_f32; /* optional [n_rois, 3, ph, pw] NCHW */
  void* out_hi;   /* optional split fp16 [n_rois, ph+2b, pw+2b, ld_out], channels 0..2 */
  void* out_lo;
  int32_t out_border, ld_out;
  /* optional scratch, n*h*w*16 bytes, 16-byte aligned: the image is normalised ONCE per pixel into pixel-interleaved
   * float4 there and the gather reads one 16-byte vector per bilinear tap (same values, a third of the loads); NULL =
   * gather from the planar image, normalising every tap */
  void* workspace;
  int64_t workspace_bytes;
} GlassImageRoiAlignParams;
int glass_image_roi_align_rotated(const GlassImageRoiAlignParams* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * Detector decision kernels (no host round trips).
 * ------------------------------------------------------------------------------------------ */
/* glass_rpn_topk_decode -- one FPN level of d2 RRPN inference: exact top-k of the objectness logits per
 * image (descending, ties by anchor index), RotatedAnchorGenerator anchors for the selected cells and
 * Box2BoxTransformRotated.apply_deltas.  replaces: RPN.predict_proposals + find_top_rrpn_proposals'
 * per-level sort/top-k (RotatedRPN inherits them: glass/modeling/proposal_generator/rotated_rpn.py:17;
 * configs/glass_pretrain.yaml:55-71).  pred = the RPN head's fused 1x1 output, fp32 [n_img,h,w,ld]:
 * columns [0,A) objectness, [A,6A) deltas ordered (anchor, 5). */
typedef struct {
  const float* pred;
  int32_t n_img, h, w, ld, num_anchors, stride;
  float anchor_w[16], anchor_h[16], anchor_angle[16];
  float weights[5];
  int32_t topk;               /* <= 1024 */
  int32_t level, num_levels;  /* output slot [level*topk, (level+1)*topk) */
  float* out_boxes;           /* [n_img, num_levels*topk, 5] */
  float* out_scores;          /* [n_img, num_levels*topk]; -inf marks an empty slot */
  void* workspace;            /* dense sortable keys + top-byte histogram */
  int64_t workspace_bytes;    /* >= glass_rpn_topk_workspace_bytes(n_img, h, w, num_anchors) */
} GlassRpnTopkParams;
int64_t glass_rpn_topk_workspace_bytes(int n_img, int h, int w, int num_anchors);
int glass_rpn_topk_decode(const GlassRpnTopkParams* p, void* stream);

/* glass_nms_rotated -- per image: optional RotatedBoxes.clip + nonempty filter, score threshold,
 * batched_nms_rotated (group offsets), stable descending sort, greedy rotated NMS, first max_keep kept.
 * replaces: torch.ops.detectron2.nms_rotated via batched_nms_rotated at rotated_fast_rcnn.py:131 and in
 * d2 find_top_rrpn_proposals.  Suppression test: iou > iou_thresh. */
typedef struct {
  const float* boxes;    /* [n_img, m, 5] */
  const float* scores;   /* [n_img, m] (non-finite = invalid slot) */
  const int32_t* group;  /* optional [n_img, m] category ids; NULL: group = index / group_size */
  int32_t group_size;    /* 0 = one group */
  const int32_t* m_dev;  /* optional per-image candidate count */
  int32_t n_img, m;      /* m <= 8192 */
  const float* img_hw;   /* [n_img, 2] (height, width), needed when clip != 0 */
  int32_t clip, filter_empty;
  float score_thresh;    /* keep score > score_thresh; pass -INFINITY to keep all */
  float iou_thresh;
  int32_t max_keep;      /* <= 128 */
  float* out_boxes;      /* [n_img, max_keep, 5] (cleaned boxes, zero padded) */
  float* out_scores;     /* [n_img, max_keep] */
  int32_t* out_index;    /* [n_img, max_keep] index of the kept candidate, -1 padded */
  int32_t* out_count;    /* [n_img] */
  void* workspace;
  int64_t workspace_bytes; /* >= glass_nms_workspace_bytes(n_img, m) */
} GlassNmsParams;
int64_t glass_nms_workspace_bytes(int n_img, int m);
int glass_nms_rotated(const GlassNmsParams* p, void* stream);

/* glass_box_decode -- RotatedFastRCNNOutputs.inference up to the score filter (rotated_fast_rcnn.py:
 * 335-373, 480-491): apply_deltas(host_weights) on the proposals, softmax over (fg, bg), orientation
 * (argmax, max prob), finite filter.  pred fp32 [n_img*per_img, ld]: cols 0-1 class scores, 2-6 box
 * deltas, 7-10 orientation logits.  out_scores = P(fg) or -inf for invalid / padded rows. */
int glass_box_decode(const float* pred, int ld, const float* proposals, const int32_t* counts, int n_img,
                     int per_img, const float* host_weights, float* out_boxes, float* out_scores,
                     float* out_orient, void* stream);

/* glass_pack_rois -- the hand-over from the box branch to the recognizer (recognizers_hybrid_head.py:176-181 runs
 * forward_with_given_boxes on the detected boxes), kept on the device: boxes fp32 [n_img, max_det, 5] + counts int32
 * [n_img] -> rois fp32 [n_img*max_det, 6] = (image, cx, cy, w, h, angle) of the live detections grouped by image (zero
 * rows after them), word_start int32 [n_img+1] prefix offsets, total int32 [1] = live words.  The recognizer's kernels
 * take `total` as their device-side count, so the step has no host round trip. */
int glass_pack_rois(const float* boxes, const int32_t* counts, int n_img, int max_det, float* rois, int32_t* word_start,
                    int32_t* total, void* stream);
/* glass_pack_detections -- the fixed-size per-image record of the single all-gather at the end of the loop
 * (replaces the pickled comm.gather of glass/evaluation/text_evaluator.py:246-249):
 * rec fp32 [n_img, max_det, 10 + steps*classes] = (valid, box 5, score, class, orientation 2, text probabilities), zero
 * rows past each image's count.  probs fp32 [words, steps, classes] in word_start order; orient optional [n_img, max_det, 2]. */
int glass_pack_detections(const float* boxes, const float* scores, const float* orient, const int32_t* counts,
                          const float* probs, const int32_t* word_start, int n_img, int max_det, int steps, int classes,
                          float* rec, void* stream);

/* ------------------------------------------------------------------------------------------
 * Recognizer-head kernels.
 * ------------------------------------------------------------------------------------------ */
/* glass_gc_attention -- MultiAspectGCAttention up to (not including) its 3x3 output conv
 * (glass/modeling/fusion/fusion_modules.py:91-127 spatial_pool, :129-156 channel_add): per word, 8-head
 * softmax-pooled global context over the h*w positions, 512->256->LayerNorm->ReLU->512 MLP, broadcast add.
 * f / y: split-fp16 padded NHWC [n_words, h+2b, w+2b, 512] in CONCAT channel order (0..255 local,
 * 256..511 global); the reference's channel interleave (`order`, :50-53) is folded into the weights:
 * w_mask[f], w1t[k=f][256], w2t[k][f], b2[f] are given in concat order. */
typedef struct {
  const void* f_hi;
  const void* f_lo;
  void* y_hi;
  void* y_lo;
  int32_t n_words, h, w, border, channels;
  const float* w_mask; /* [512] */
  float b_mask;
  const float* w1t;    /* [512][256] */
  const float* b1;     /* [256] */
  const float* ln_g;   /* [256] */
  const float* ln_b;   /* [256] */
  const float* w2t;    /* [256][512] */
  const float* b2;     /* [512] */
  const int32_t* n_words_dev; /* optional live word count on the device (<= n_words, which then is the capacity) */
} GlassGcAttentionParams;
int glass_gc_attention(const GlassGcAttentionParams* p, void* stream);

/* mean over H of a split activation [n,h+2b,w+2b,cp] -> rows [n*w, cp] (split fp16 + optional fp32):
 * BiLSTMBlockV2.forward's feats.mean(dim=2) (recognizer_encoder.py:118-120). */
int glass_hmean_rows(const void* src_hi, const void* src_lo, int n, int h, int w, int cp, int border, void* dst_hi,
                     void* dst_lo, float* dst_f32, const int32_t* n_dev, void* stream);

/* nn.LSTM(bidirectional) recurrence (recognizer_encoder.py:141-142).  gates_in fp32 [n_seq*T, 8*hidden] =
 * x W_ih^T + b_ih + b_hh (forward gates | backward gates, order i,f,g,o); whh_t fp32 [2][hidden][4*hidden].
 * Output rows [n_seq*T, 2*hidden] (forward | backward) as split fp16 (+ optional fp32). hidden must be 256. */
int glass_lstm_bidir(const float* gates_in, const float* whh_t, int n_seq, int T, int hidden, void* out_hi,
                     void* out_lo, float* out_f32, const int32_t* n_dev, void* stream);

/* glass_aster_decode -- AttentionRecognitionHead.sample (prediction_aster.py:63-99; DecoderUnit :291-302,
 * AttentionUnit :247-266): greedy additive-attention GRU decoding, all steps in one persistent kernel.
 * Weights are transposed to [in][out].  probs = softmax outputs (pred_text_prob); logits = pre-softmax tap.
 * first_eos[w] = first step whose argmax is class 0, or `steps`.
 * The GRU's input product W_ih . [Emb[y_prev] ; context] + b_ih is taken from two precomputed tensors instead of
 * streaming W_ih on every step: emb_gi fp32 [num_classes][3*dim] = W_ih[:, :dim] . Emb[y] + b_ih (built at weight-packing
 * time: y takes num_classes values) and pctx fp32 [n_words, T, 3*dim] = x . W_ih[:, dim:]^T (one glass_conv_gemm per
 * launch, like xproj), since W . (sum_t alpha_t x_t) = sum_t alpha_t (W . x_t). */
typedef struct {
  const float* xproj;  /* [n_words, T, dim] xEmbed(x) */
  const float* pctx;   /* [n_words, T, 3*dim] */
  int32_t n_words;     /* words (capacity of the buffers when n_words_dev is given) */
  const int32_t* n_words_dev; /* optional live word count on the device (<= n_words) */
  int32_t T, steps, num_classes, dim;
  /* [sEmbed ; W_hh] (4*dim x dim: every product with the hidden state) as mma.sync m16n8k16 A FRAGMENTS of its split-fp16
   * planes (weights x 64): uint4 [4*dim/16 m-tiles][dim/16 k-steps][2 planes hi, lo][32 lanes], lane (g = lane / 4,
   * q = lane % 4) holding {(row g, k 2q..2q+1), (row g+8, same k), (row g, k 2q+8..2q+9), (row g+8, same k)} of its tile
   * (packing.pack_decoder_h_weights).  The step streams them from L2 in this order: 512 contiguous bytes per warp load. */
  const void* wh_frag;
  const float* bh;     /* [4*dim] sEmbed bias | b_hh */
  const float* we;     /* [dim] */
  float be;
  const float* emb_gi; /* [num_classes][3*dim] */
  const float* wo_t;   /* [dim][num_classes] */
  const float* bo;
  float temperature;
  float* probs;        /* [n_words, steps, num_classes] */
  float* logits;       /* optional */
  float* alphas;       /* optional [n_words, steps, T] */
  int32_t* first_eos;  /* [n_words] */
} GlassAsterParams;
int glass_aster_decode(const GlassAsterParams* p, void* stream);
/* The reference's batch-level early break (prediction_aster.py:91-93): zero the rows after the step at which
 * every word of an image has emitted class 0.  word_start: int32 [n_img+1] prefix offsets of each image's words. */
int glass_aster_finalize(float* probs, const int32_t* first_eos, const int32_t* word_start, int n_img, int steps,
                         int num_classes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Post-processing: the step that follows the hot path in every caller.
 * ------------------------------------------------------------------------------------------ */
/* glass_postprocess_merge -- PostProcessorRotatedBoxes.__call__ (glass/postprocess/post_processor_rotated_boxes.py:
 * 66-184) + the text-score filter of PostProcessorAcademic.__call__ (post_processor_academic.py:26-34), one CTA per
 * image, no host round trips: min(w,h) >= min_box_dim and score >= valid_score filters; then until no pair merges:
 * rotated intersection-over-min-area of every pair (glass/structures/boxes.py:24-49), pair masks (IoA, angle,
 * height ratio), min-area-rectangle merge of each valid pair (replaces cv2.minAreaRect on the host, :251-286),
 * reference write-back order, nms_rotated(nms_iou); finally score >= detect_threshold, text score >= text_threshold
 * (when text_scores is given) and boxes_to_polygons (:221-249).  Survivors are written in the reference's order
 * (descending score once a merge round has run), zero / -1 padded to m. */
typedef struct {
  const float* boxes;       /* [n_img, m, 5] (cx, cy, w, h, angle_deg) */
  const float* scores;      /* [n_img, m] */
  const float* text_scores; /* optional [n_img, m] word scores (glass_text_scores) */
  const int32_t* counts;    /* optional [n_img] detections per image (default m) */
  int32_t n_img, m;         /* m <= 128 */
  float min_box_dim, valid_score, detect_threshold, text_threshold;           /* cfg.POST_PROCESSING.* */
  float merge_ioa_thresh, pairs_height_ratio_thresh, max_angle_diff;
  float minimal_ioa_thresh; /* 0.01, post_processor_rotated_boxes.py:40 */
  float nms_iou;            /* 0.99, :181 */
  int32_t max_iters;        /* safety bound on merge rounds (the reference loops until nothing merges) */
  float* out_boxes;         /* [n_img, m, 5] */
  float* out_scores;        /* [n_img, m] */
  float* out_polygons;      /* optional [n_img, m, 4, 2] */
  int32_t* out_index;       /* [n_img, m] index of the surviving detection in the input, -1 padded */
  int32_t* out_count;       /* [n_img] */
  int32_t* out_iters;       /* optional [n_img] merge rounds run */
} GlassPostprocessParams;
int glass_postprocess_merge(const GlassPostprocessParams* p, void* stream);

/* glass_text_scores -- numeric part of get_instances_text (glass/evaluation/text_evaluator.py:323-331) +
 * TextEncoder.decode_attention's word score (glass/modeling/recognition/text_encoder.py:80-151): per decoding
 * step max / argmax over the classes; score = product of the max probabilities up to and including the first
 * stop symbol.  probs fp32 [n_words, steps, classes]; out_idx / out_maxp optional [n_words, steps]. */
int glass_text_scores(const float* probs, int n_words, int steps, int classes, int stop_index, float* score,
                      int32_t* out_idx, float* out_maxp, const int32_t* n_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * Mask branch (MODEL.ROI_MASK_HEAD.MASK_INFERENCE): the head's convolutions run on glass_conv_gemm
 * (the 2x2/s2 deconv as a GEMM with one 256-column block per output sub-pixel, the predictor block-diagonal over them).
 * ------------------------------------------------------------------------------------------ */
/* mask_rcnn_inference (detectron2) + sub-pixel scatter: logits fp32 [K*(h+2)*(w+2), ld] over the padded pooled plane,
 * columns dy*2+dx -> masks fp32 [K, 2h, 2w] = sigmoid(logit). */
int glass_mask_finalize(const float* logits, int ld, int k_words, int h, int w, float* masks, void* stream);
/* paste_masks_in_image / _do_paste_mask for rotated boxes (glass/postprocess/post_processor_academic.py:187-335):
 * masks fp32 [k, m, m] probabilities, boxes [k, 5] in image coordinates -> out uint8 [k, img_h, img_w] = (sampled mask
 * >= threshold); out_soft optional fp32 [k, img_h, img_w] (the sampled values, for tests). */
int glass_paste_masks_rotated(const float* masks, const float* boxes, int k, int m, int img_h, int img_w, float threshold,
                              uint8_t* out, float* out_soft, void* stream);

/* ------------------------------------------------------------------------------------------
 * detectron2 operator surface outside the fused path: what the reference's own Python reaches through
 * torch.ops.detectron2.* when it is NOT inside the hot path's fused kernels (post-processor, evaluation, user code).
 * ------------------------------------------------------------------------------------------ */
/* glass_box_iou_rotated -- pairwise IoU of rotated boxes (cx, cy, w, h, angle_deg): boxes1 fp32 [n1,5], boxes2 [n2,5]
 * -> out fp32 [n1, n2].  mode 0 replaces torch.ops.detectron2.box_iou_rotated (pairwise_iou_rotated, called at
 * glass/structures/boxes.py:34); mode 1 = glass.structures.boxes.pairwise_ioa_rotated (:24-49: intersection over the
 * smaller area, recovered from the IoU in the reference's operation order; call site
 * glass/postprocess/post_processor_rotated_boxes.py:128). */
int glass_box_iou_rotated(const float* boxes1, int n1, const float* boxes2, int n2, int mode, float* out, void* stream);
/* glass_nms_rotated_all -- torch.ops.detectron2.nms_rotated(boxes, scores, iou_threshold) with EVERY survivor returned
 * (call sites: glass/postprocess/post_processor_rotated_boxes.py:181; batched through coordinate offsets at
 * glass/modeling/roi_heads/rotated_fast_rcnn.py:131).  boxes fp32 [n,5]; order int64 [n] = indices by descending score
 * (the caller's sort; detectron2 sorts the same way before its mask kernel); a box is suppressed when its IoU with an
 * earlier kept box is > iou_thresh.  keep int64 [n]: kept indices in score order, -1 padded; keep_count int32 [1].
 * No host round trip (detectron2 scans its bitmask on the host). */
int64_t glass_nms_rotated_all_workspace_bytes(int n);
int glass_nms_rotated_all(const float* boxes, const int64_t* order, int n, float iou_thresh, int64_t* keep,
                          int32_t* keep_count, void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GLASS_B200_H_ */
