"""ctypes binding of libglass_b200.so (C ABI: include/glass_b200.h)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GLASS_B200_LIB", os.path.join(HERE, "_lib", "libglass_b200.so"))  # env: A/B builds

MAX_TAPS = 16
MAX_LEVELS = 5

# every symbol include/glass_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "glass_last_error", "glass_abi_version", "glass_launch_count", "glass_conv_gemm", "glass_pack_nchw",
    "glass_unpack_nchw", "glass_nhwc_f32_to_nchw", "glass_gather_taps", "glass_maxpool",
    "glass_roi_align_rotated", "glass_image_roi_align_rotated", "glass_rpn_topk_decode", "glass_rpn_topk_workspace_bytes",
    "glass_nms_workspace_bytes",
    "glass_nms_rotated", "glass_box_decode", "glass_gc_attention", "glass_hmean_rows", "glass_lstm_bidir",
    "glass_aster_decode", "glass_aster_finalize", "glass_resize_bilinear_u8", "glass_postprocess_merge", "glass_text_scores", "glass_zero_border", "glass_stem_s2d", "glass_mask_finalize", "glass_paste_masks_rotated",
    "glass_box_iou_rotated", "glass_nms_rotated_all", "glass_nms_rotated_all_workspace_bytes",
    "glass_plan_create", "glass_plan_launch", "glass_plan_destroy", "glass_prepack_weights", "glass_pack_rois",
    "glass_pack_detections",
]


class ConvGemmParams(C.Structure):
    _fields_ = [
        ("a_hi", C.c_void_p), ("a_lo", C.c_void_p), ("rows_a", C.c_int64), ("a_ld", C.c_int32),
        ("k_per_tap", C.c_int32), ("ntaps", C.c_int32), ("tap_shift", C.c_int32 * MAX_TAPS),
        ("b_hi", C.c_void_p), ("b_lo", C.c_void_p), ("n", C.c_int32), ("mode", C.c_int32),
        ("m_imgs", C.c_int32), ("m_h", C.c_int32), ("m_w", C.c_int32), ("m_border", C.c_int32),
        ("scale", C.c_void_p), ("bias", C.c_void_p), ("relu_pre", C.c_int32), ("relu_post", C.c_int32),
        ("res_hi", C.c_void_p), ("res_lo", C.c_void_p),
        ("res_hp", C.c_int32), ("res_wp", C.c_int32), ("res_border", C.c_int32), ("res_shift", C.c_int32),
        ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("out_f32", C.c_void_p),
        ("out_hp", C.c_int32), ("out_wp", C.c_int32), ("out_border", C.c_int32),
        ("ld_out", C.c_int32), ("ld_f32", C.c_int32), ("n_store", C.c_int32), ("kb_per_chunk", C.c_int32), ("pair_mode", C.c_int32), ("tap_mode", C.c_int32),
        ("a_col0", C.c_int32), ("a_inner", C.c_int32), ("m_count_dev", C.c_void_p), ("m_rows_per_count", C.c_int32), ("epi_mode", C.c_int32), ("sat_count", C.c_void_p),
    ]


class RoiAlignParams(C.Structure):
    _fields_ = [
        ("num_levels", C.c_int32), ("feat", C.c_void_p * MAX_LEVELS), ("feat_lo", C.c_void_p * MAX_LEVELS),
        ("feat_is_split", C.c_int32),
        ("feat_h", C.c_int32 * MAX_LEVELS), ("feat_w", C.c_int32 * MAX_LEVELS),
        ("spatial_scale", C.c_float * MAX_LEVELS),
        ("feat_border", C.c_int32), ("feat_ld", C.c_int32), ("channels", C.c_int32), ("min_level", C.c_int32),
        ("rois", C.c_void_p), ("n_rois_dev", C.c_void_p), ("n_rois", C.c_int32),
        ("pooled_h", C.c_int32), ("pooled_w", C.c_int32), ("sampling_ratio", C.c_int32),
        ("out_f32", C.c_void_p), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p),
        ("out_hp", C.c_int32), ("out_wp", C.c_int32), ("out_border", C.c_int32), ("out_coff", C.c_int32),
        ("ld_out", C.c_int32),
    ]


class ImageRoiAlignParams(C.Structure):
    _fields_ = [
        ("img", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("h_pad", C.c_int32), ("w_pad", C.c_int32), ("mean", C.c_float * 3), ("inv_std", C.c_float * 3),
        ("rois", C.c_void_p), ("n_rois_dev", C.c_void_p), ("n_rois", C.c_int32),
        ("pooled_h", C.c_int32), ("pooled_w", C.c_int32), ("sampling_ratio", C.c_int32),
        ("out_f32", C.c_void_p), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p),
        ("out_border", C.c_int32), ("ld_out", C.c_int32), ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
    ]


class RpnTopkParams(C.Structure):
    _fields_ = [
        ("pred", C.c_void_p), ("n_img", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("ld", C.c_int32),
        ("num_anchors", C.c_int32), ("stride", C.c_int32),
        ("anchor_w", C.c_float * 16), ("anchor_h", C.c_float * 16), ("anchor_angle", C.c_float * 16),
        ("weights", C.c_float * 5), ("topk", C.c_int32), ("level", C.c_int32), ("num_levels", C.c_int32),
        ("out_boxes", C.c_void_p), ("out_scores", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
    ]


class NmsParams(C.Structure):
    _fields_ = [
        ("boxes", C.c_void_p), ("scores", C.c_void_p), ("group", C.c_void_p), ("group_size", C.c_int32),
        ("m_dev", C.c_void_p), ("n_img", C.c_int32), ("m", C.c_int32), ("img_hw", C.c_void_p),
        ("clip", C.c_int32), ("filter_empty", C.c_int32), ("score_thresh", C.c_float), ("iou_thresh", C.c_float),
        ("max_keep", C.c_int32), ("out_boxes", C.c_void_p), ("out_scores", C.c_void_p), ("out_index", C.c_void_p),
        ("out_count", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
    ]


class GcAttentionParams(C.Structure):
    _fields_ = [
        ("f_hi", C.c_void_p), ("f_lo", C.c_void_p), ("y_hi", C.c_void_p), ("y_lo", C.c_void_p),
        ("n_words", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("border", C.c_int32), ("channels", C.c_int32),
        ("w_mask", C.c_void_p), ("b_mask", C.c_float), ("w1t", C.c_void_p), ("b1", C.c_void_p), ("ln_g", C.c_void_p),
        ("ln_b", C.c_void_p), ("w2t", C.c_void_p), ("b2", C.c_void_p), ("n_words_dev", C.c_void_p),
    ]


class AsterParams(C.Structure):
    _fields_ = [
        ("xproj", C.c_void_p), ("pctx", C.c_void_p), ("n_words", C.c_int32), ("n_words_dev", C.c_void_p),
        ("T", C.c_int32), ("steps", C.c_int32), ("num_classes", C.c_int32), ("dim", C.c_int32),
        ("wh_frag", C.c_void_p), ("bh", C.c_void_p), ("we", C.c_void_p), ("be", C.c_float), ("emb_gi", C.c_void_p),
        ("wo_t", C.c_void_p), ("bo", C.c_void_p), ("temperature", C.c_float),
        ("probs", C.c_void_p), ("logits", C.c_void_p), ("alphas", C.c_void_p), ("first_eos", C.c_void_p),
    ]


class PostprocessParams(C.Structure):
    _fields_ = [
        ("boxes", C.c_void_p), ("scores", C.c_void_p), ("text_scores", C.c_void_p), ("counts", C.c_void_p),
        ("n_img", C.c_int32), ("m", C.c_int32),
        ("min_box_dim", C.c_float), ("valid_score", C.c_float), ("detect_threshold", C.c_float),
        ("text_threshold", C.c_float), ("merge_ioa_thresh", C.c_float), ("pairs_height_ratio_thresh", C.c_float),
        ("max_angle_diff", C.c_float), ("minimal_ioa_thresh", C.c_float), ("nms_iou", C.c_float),
        ("max_iters", C.c_int32),
        ("out_boxes", C.c_void_p), ("out_scores", C.c_void_p), ("out_polygons", C.c_void_p),
        ("out_index", C.c_void_p), ("out_count", C.c_void_p), ("out_iters", C.c_void_p),
    ]


_lib = None


def load() -> C.CDLL:
    """Load the extension; fails loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m glass_text_spotting_b200.build` "
            "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.glass_last_error.restype = C.c_char_p
    lib.glass_abi_version.restype = C.c_int
    lib.glass_launch_count.restype = C.c_int64
    i, p, f = C.c_int, C.c_void_p, C.POINTER(C.c_float)
    lib.glass_conv_gemm.argtypes = [C.POINTER(ConvGemmParams), p]
    lib.glass_plan_create.argtypes = [C.POINTER(ConvGemmParams), C.POINTER(C.c_void_p)]
    lib.glass_plan_launch.argtypes = [p, p]
    lib.glass_plan_destroy.argtypes = [p]
    lib.glass_prepack_weights.argtypes = [p, i, i, p, p, p, p]
    lib.glass_pack_rois.argtypes = [p, p, i, i, p, p, p, p]
    lib.glass_pack_detections.argtypes = [p, p, p, p, p, p, i, i, i, i, p, p]
    lib.glass_pack_nchw.argtypes = [p, i, i, i, i, p, p, i, i, p]
    lib.glass_unpack_nchw.argtypes = [p, p, i, i, i, i, i, i, p, p]
    lib.glass_nhwc_f32_to_nchw.argtypes = [p, i, i, i, i, i, i, p, p]
    lib.glass_gather_taps.argtypes = [p, p] + [i] * 13 + [p, p, i, p, p]
    lib.glass_maxpool.argtypes = [p, p] + [i] * 13 + [p, p, i, p, p]
    lib.glass_roi_align_rotated.argtypes = [C.POINTER(RoiAlignParams), p]
    lib.glass_image_roi_align_rotated.argtypes = [C.POINTER(ImageRoiAlignParams), p]
    lib.glass_rpn_topk_decode.argtypes = [C.POINTER(RpnTopkParams), p]
    lib.glass_rpn_topk_workspace_bytes.argtypes = [i, i, i, i]
    lib.glass_rpn_topk_workspace_bytes.restype = C.c_int64
    lib.glass_nms_workspace_bytes.argtypes = [i, i]
    lib.glass_nms_workspace_bytes.restype = C.c_int64
    lib.glass_nms_rotated.argtypes = [C.POINTER(NmsParams), p]
    lib.glass_box_decode.argtypes = [p, i, p, p, i, i, f, p, p, p, p]
    lib.glass_gc_attention.argtypes = [C.POINTER(GcAttentionParams), p]
    lib.glass_hmean_rows.argtypes = [p, p, i, i, i, i, i, p, p, p, p, p]
    lib.glass_lstm_bidir.argtypes = [p, p, i, i, i, p, p, p, p, p]
    lib.glass_aster_decode.argtypes = [C.POINTER(AsterParams), p]
    lib.glass_aster_finalize.argtypes = [p, p, p, i, i, i, p]
    lib.glass_resize_bilinear_u8.argtypes = [p, i, i, i, p, i, i, p]
    lib.glass_postprocess_merge.argtypes = [C.POINTER(PostprocessParams), p]
    lib.glass_text_scores.argtypes = [p, i, i, i, i, p, p, p, p, p]
    lib.glass_zero_border.argtypes = [p, p, i, i, i, i, p, p]
    lib.glass_stem_s2d.argtypes = [p, i, i, i, f, f, p, p, i, p]
    lib.glass_mask_finalize.argtypes = [p, i, i, i, i, p, p]
    lib.glass_paste_masks_rotated.argtypes = [p, p, i, i, i, i, C.c_float, p, p, p]
    lib.glass_box_iou_rotated.argtypes = [p, i, p, i, i, p, p]
    lib.glass_nms_rotated_all_workspace_bytes.argtypes = [i]
    lib.glass_nms_rotated_all_workspace_bytes.restype = C.c_int64
    lib.glass_nms_rotated_all.argtypes = [p, p, i, C.c_float, p, p, p, C.c_int64, p]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name.startswith("glass_") and name not in ("glass_last_error", "glass_abi_version", "glass_launch_count",
                                                        "glass_nms_workspace_bytes", "glass_rpn_topk_workspace_bytes",
                                                        "glass_nms_rotated_all_workspace_bytes"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError("libglass_b200: " + load().glass_last_error().decode())
