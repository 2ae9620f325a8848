"""ctypes signatures for the non-GEMM entry points of include/b200vsgg.h (same order as the header;
tests/test_capi_symbols.py checks that every function the header declares is exported and listed)."""
import ctypes as C

i32, i64, vp, f32, u64 = C.c_int32, C.c_int64, C.c_void_p, C.c_float, C.c_uint64


class GmmHead(C.Structure):
    """Mirror of struct b200vsgg_gmm_head."""
    _fields_ = [("col_base", i32), ("num_classes", i32), ("softmax", i32), ("eps", vp), ("out", vp),
                ("out2", vp), ("dout", vp)]


class ObjTokens(C.Structure):
    """Mirror of struct b200vsgg_obj_tokens."""
    _fields_ = [("features", vp), ("feat_dim", i32), ("dist", vp), ("n_cls", i32), ("embed", vp), ("e", i32),
                ("boxes", vp), ("bn_mean", vp), ("bn_rstd", vp), ("bn_gamma", vp), ("bn_beta", vp), ("video_of_box", vp),
                ("wp", vp), ("bp", vp), ("h", i32), ("pe", vp), ("src", vp), ("pos", vp), ("rows", i32),
                ("p_pos", f32), ("seed_pos", u64), ("p_pe", f32), ("seed_pe", u64)]


# name -> argtypes (restype is always int32)
SIGNATURES = {
    "b200vsgg_frame_offsets": [vp, i32, i32, vp, vp],
    "b200vsgg_gather_rows": [vp, i32, vp, vp, vp, i32, i32, vp, i32, vp, i32, vp, i32, vp],
    "b200vsgg_gather2_sum_rows": [vp, i32, vp, vp, i32, i32, i32, vp, i32, vp, i32, vp],
    "b200vsgg_gather_rows_bf16": [vp, i32, vp, i32, i32, vp, i32, vp],
    "b200vsgg_gather2_sum_rows_bf16": [vp, i32, vp, i32, i32, vp, i32, vp],
    "b200vsgg_pair_concat_fwd": [vp, vp, vp, vp, vp, i32, vp, vp, vp],
    "b200vsgg_pair_concat_bwd": [vp, vp, vp, i32, vp, vp, vp, vp],
    "b200vsgg_layernorm_fwd": [vp, i32, vp, vp, i32, i32, f32, vp, i32, vp, i32, vp, vp, vp, i32, vp, vp, vp],
    "b200vsgg_layernorm_bwd": [vp, i32, vp, i32, vp, vp, vp, i32, i32, vp, i32, vp, i32, f32, u64, vp, vp, vp],
    "b200vsgg_layernorm_bwd_add": [vp, i32, vp, i32, vp, vp, vp, i32, i32, vp, i32, vp, i32, f32, u64, vp, vp, vp, vp, i32],
    "b200vsgg_cast_dropout_bf16": [vp, i32, i32, i32, vp, i32, f32, u64, vp],
    "b200vsgg_split3_bf16": [vp, i32, i32, i32, vp, i32, vp],
    "b200vsgg_colsum": [vp, i32, i32, i32, i32, vp, i32, vp, vp],
    "b200vsgg_attn_small_fwd": [vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, i32, f32, vp, i32, f32, u64, vp],
    "b200vsgg_attn_small_bwd": [vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, i32, f32, vp, i32, vp, i32,
                                vp, i32, f32, u64, vp],
    "b200vsgg_attn_rows_fwd": [vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, f32, vp, i32, vp, f32, u64, vp],
    "b200vsgg_attn_rows_bwd": [vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, vp, vp, vp, i32, i32, i32, f32, vp, i32, vp, i32,
                               vp, i32, f32, u64, vp],
    "b200vsgg_gmm_head_fwd": [vp, i32, i32, i32, C.POINTER(GmmHead), i32, i32, u64, vp],
    "b200vsgg_gmm_head_bwd": [vp, i32, i32, i32, C.POINTER(GmmHead), i32, i32, u64, vp, i32, i32, vp],
    "b200vsgg_nchw_to_nhwc_bf16": [vp, i32, i32, i32, vp, vp],
    "b200vsgg_nchw_to_nhwc_f32": [vp, i32, i32, i32, vp, vp],
    "b200vsgg_nhwc_to_nchw_f32": [vp, i32, i32, i32, vp, vp],
    "b200vsgg_mask_im2col": [vp, i32, vp, i32, vp],
    "b200vsgg_mask_im2col_bf16": [vp, i32, vp, i32, vp],
    "b200vsgg_seg_colstats": [vp, i32, i32, vp, i32, i32, vp, i32, vp, vp, vp],
    "b200vsgg_seg_affine": [vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, vp, vp],
    "b200vsgg_bn_pool_fwd": [vp, i32, vp, vp, vp, i32, i32, i32, vp, vp, vp],
    "b200vsgg_pool_bwd": [vp, vp, i32, i32, i32, vp, vp],
    "b200vsgg_im2col3x3": [vp, i32, i32, i32, vp, vp],
    "b200vsgg_col2im3x3": [vp, i32, i32, i32, vp, vp],
    "b200vsgg_attn_tc_fwd": [vp, i32, vp, i32, vp, i32, i32, vp, vp, vp, i32, i32, i32, f32, vp, i32, vp, f32, u64, vp],
    "b200vsgg_attn_tc_bwd": [vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, vp, vp, i32, vp, vp, vp, i32, i32, i32, f32, vp, i32,
                             vp, i32, vp, i32, f32, u64, vp],
    "b200vsgg_attn_flash_fwd": [vp, i32, vp, i32, vp, i32, vp, vp, vp, i32, i32, i32, f32, vp, i32, vp, f32, u64, vp, i32,
                                i32],
    "b200vsgg_attn_flash_bwd": [vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32,
                                vp, i32, vp, i32, vp, i32, f32, u64, vp, i32, i32],
    "b200vsgg_node_tokens_fwd": [vp, i32, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp],
    "b200vsgg_node_tokens_bwd": [vp, vp, vp, vp, i32, i32, i32, vp, i32, vp, vp],
    "b200vsgg_teat_pair_flags": [vp, i32, vp, vp, vp, vp, i32, f32, f32, i32, vp, vp, vp],
    "b200vsgg_teat_assemble_fwd": [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "b200vsgg_teat_assemble_bwd": [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "b200vsgg_graph_attn_core": [vp, i32, vp, vp, i32, vp, vp, i32, vp, i32, vp],
    "b200vsgg_graph_attn_core_f32": [vp, i32, vp, vp, i32, vp, vp, i32, vp, i32, vp],
    "b200vsgg_gated_residual": [vp, vp, vp, i32, i32, vp],
    "b200vsgg_grad_sqnorm": [vp, vp, vp, i32, i32, vp, vp],
    "b200vsgg_adamw_clip_step": [vp, vp, vp, i32, i32, vp, f32, f32, f32, f32, f32, f32, vp],
    "b200vsgg_obj_tokens_fwd": [C.POINTER(ObjTokens), vp, vp, vp],
    "b200vsgg_obj_tokens_bwd": [C.POINTER(ObjTokens), vp, vp, vp, vp, vp, vp, vp],
    "b200vsgg_rel_loss": [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "b200vsgg_contrastive_loss": [vp, vp, vp, i32, i32, i32, f32, f32, vp, vp, vp, vp],
    "b200vsgg_graph_small_params_per_layer": [i32, i32],
    "b200vsgg_graph_small_fwd": [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp],
    "b200vsgg_upload": [vp, vp, i64, vp],
    "b200vsgg_consistency_kl_bwd": [vp, i32, vp, vp, vp, i32, vp, vp],
    "b200vsgg_attn_pool_bwd": [vp, i32, vp, i32, i32, vp, vp, vp, vp, vp, vp],
    "b200vsgg_weighted_colsum": [vp, i32, i32, i32, i32, vp, vp, vp],
    "b200vsgg_gated_residual_bwd": [vp, vp, vp, vp, i32, i32, vp, vp, vp, vp],
    "b200vsgg_graph_attn_core_bwd": [vp, i32, vp, vp, i32, vp, vp, vp, i32, i32, vp, i32, vp, vp, vp],
    "b200vsgg_simt_linear": [vp, i32, vp, i32, i32, vp, i64, i32, i32, i32, vp, i32, vp, vp],
    "b200vsgg_simt_wgrad": [vp, i32, vp, i32, i32, i32, i32, vp, vp],
    "b200vsgg_gelu_bwd": [vp, vp, i64, vp, vp],
    "b200vsgg_ln_small_fwd": [vp, vp, vp, i32, i32, vp, vp, vp, vp],
    "b200vsgg_ln_small_bwd": [vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp],
    "b200vsgg_attn_pool": [vp, i32, vp, i32, i32, vp, vp, vp, vp],
    "b200vsgg_class_memory_accumulate": [vp, i32, i32, vp, vp, vp, i32, i32, vp, vp],
    "b200vsgg_interval_kl": [vp, i32, vp, vp, i32, vp, vp],
    "b200vsgg_eval_recall": [vp, vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, C.c_double,
                             C.c_double, vp, vp, vp],
    "b200vsgg_consistency_kl": [vp, i32, vp, vp, i32, vp, vp],
    "b200vsgg_act_dropout_bf16": [vp, i32, i64, i32, i32, f32, u64, vp, i32, vp],
}


def declare(lib):
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = i32
        fn.argtypes = argtypes
