"""ctypes binding of libdit4sr_b200.so (include/dit4sr_b200.h). There is no fallback: if the library is missing or
fails to load, using the ops raises."""
import ctypes as C
import os

from .build import LIB_PATH

c_i64, c_int, c_float, c_vp = C.c_int64, C.c_int, C.c_float, C.c_void_p


class LnModulateParams(C.Structure):
    _fields_ = [("x", c_vp), ("ldx", c_i64), ("x_batch_stride", c_i64),
                ("shift", c_vp), ("scale", c_vp), ("shift2", c_vp), ("scale2", c_vp),
                ("mod_batch_stride", c_i64), ("out", c_vp), ("out2", c_vp),
                ("ldo", c_i64), ("o_batch_stride", c_i64),
                ("batches", c_int), ("rows", c_int), ("dim", c_int), ("eps", c_float), ("out_f32", c_int),
                ("x_b", c_vp), ("ldx_b", c_i64), ("x_b_batch_stride", c_i64), ("shift_b", c_vp), ("scale_b", c_vp),
                ("out_b", c_vp), ("ldo_b", c_i64), ("o_b_batch_stride", c_i64), ("rows_b", c_int)]


class DwconvParams(C.Structure):
    _fields_ = [("u_hq", c_vp), ("u_lr", c_vp), ("ld", c_i64), ("hq_batch_stride", c_i64),
                ("lr_batch_stride", c_i64), ("weight9", c_vp), ("bias", c_vp),
                ("batches", c_int), ("grid", c_int), ("channels", c_int),
                ("grid_w", c_int), ("lr_grid_h", c_int), ("lr_row_offset", c_int)]


class GemmSp(C.Structure):
    _fields_ = [("world", c_int), ("rank", c_int), ("own_heads_only", c_int), ("inner_dim", c_int), ("col0", c_int),
                ("rows_dst", c_i64), ("row0", c_i64), ("peer", c_vp * 8)]


class GemmParams(C.Structure):
    _fields_ = [("A", c_vp), ("lda", c_i64), ("a_batch_stride", c_i64),
                ("W", c_vp), ("ldw", c_i64), ("bias", c_vp),
                ("C", c_vp), ("ldc", c_i64), ("c_batch_stride", c_i64), ("c_dtype", c_int),
                ("batches", c_int), ("rows", c_int), ("N", c_int), ("K", c_int), ("epilogue", c_int),
                ("gate", c_vp), ("gate_batch_stride", c_i64),
                ("resid", c_vp), ("ldr", c_i64), ("r_batch_stride", c_i64),
                ("norm_w0", c_vp), ("norm_w1", c_vp), ("sec_cols", c_int), ("eps", c_float),
                ("dbl_row_offset", c_i64), ("sp", C.POINTER(GemmSp)),
                ("conv_w", c_int), ("conv_h", c_int), ("conv_kw", c_int), ("conv_taps", c_int), ("conv_org", c_int), ("conv_w_valid", c_int),
                ("resid_bf16", c_vp), ("ld_rb", c_i64), ("rb_batch_stride", c_i64)]


class GroupNormParams(C.Structure):
    _fields_ = [("x", c_vp), ("ldx", c_i64), ("x_batch_stride", c_i64), ("y", c_vp), ("ldy", c_i64), ("y_batch_stride", c_i64),
                ("gamma", c_vp), ("beta", c_vp), ("stats", c_vp), ("stats_bytes", c_i64),
                ("batches", c_int), ("pixels", c_int), ("channels", c_int), ("groups", c_int), ("eps", c_float), ("silu", c_int),
                ("x_dtype", c_int), ("width", c_int), ("pitch", c_int)]


SP_MAX_RANKS, SP_CHANNELS = 8, 4
SP_CH_QKV, SP_CH_OUT, SP_CH_HALO, SP_CH_USER = 0, 1, 2, 3


class AttnSpOut(C.Structure):
    _fields_ = [("world", c_int), ("rows_per_rank", c_int), ("col_offset", c_int), ("o_peer", c_vp * SP_MAX_RANKS)]


class SpComm(C.Structure):
    _fields_ = [("world", c_int), ("rank", c_int), ("flags", c_vp * SP_MAX_RANKS), ("epoch", c_vp)]


class A2aParams(C.Structure):
    _fields_ = [("src", c_vp), ("src_row_stride", c_i64), ("src_batch_stride", c_i64), ("peer", c_vp * SP_MAX_RANKS),
                ("dst_row_stride", c_i64), ("dst_batch_stride", c_i64),
                ("batches", c_int), ("image_rows", c_int), ("text_rows", c_int), ("heads", c_int), ("no_barrier", c_int)]


class HaloParams(C.Structure):
    _fields_ = [("peer", c_vp * SP_MAX_RANKS), ("row_stride", c_i64), ("batch_stride", c_i64),
                ("batches", c_int), ("channels", c_int), ("grid_w", c_int),
                ("first_lr_row", c_i64), ("lr_rows", c_i64), ("top_halo_row", c_i64), ("bottom_halo_row", c_i64),
                ("no_barrier", c_int)]


class AttnParams(C.Structure):
    _fields_ = [("q", c_vp), ("k", c_vp), ("v", c_vp), ("qkv_row_stride", c_i64), ("qkv_batch_stride", c_i64),
                ("o", c_vp), ("o_row_stride", c_i64), ("o_batch_stride", c_i64),
                ("lr_res", c_vp), ("lr_row_stride", c_i64), ("lr_batch_stride", c_i64),
                ("lr_row_begin", c_int), ("lr_row_end", c_int),
                ("batches", c_int), ("heads", c_int), ("seq_q", c_int), ("seq_kv", c_int), ("scale", c_float),
                ("mode", c_int), ("score_bound", c_float), ("exp_offload", c_int), ("variant", c_int),
                ("sp_out", C.POINTER(AttnSpOut)), ("workspace", c_vp), ("workspace_bytes", c_i64), ("split_kv", c_int)]


# every symbol include/dit4sr_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "dit4sr_abi_version": (c_int, []),
    "dit4sr_last_error": (C.c_char_p, []),
    "dit4sr_check_device": (c_int, []),
    "dit4sr_ln_modulate": (c_int, [C.POINTER(LnModulateParams), c_vp]),
    "dit4sr_dwconv3x3_add": (c_int, [C.POINTER(DwconvParams), c_vp]),
    "dit4sr_gemm_bf16": (c_int, [C.POINTER(GemmParams), c_vp]),
    "dit4sr_gemm_bf16_grouped": (c_int, [C.POINTER(GemmParams), c_int, c_vp]),
    "dit4sr_gemm_schedule": (c_int, [C.POINTER(GemmParams), c_int, c_int, C.POINTER(c_int), c_int]),
    "dit4sr_attn_fwd": (c_int, [C.POINTER(AttnParams), c_vp]),
    "dit4sr_attn_workspace_bytes": (c_i64, []),
    "dit4sr_attn_last_plan": (c_int, [C.POINTER(c_int)]),
    "dit4sr_attn_plan": (c_int, [C.POINTER(AttnParams), c_int, C.POINTER(c_int)]),
    "dit4sr_patchify": (c_int, [c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_vp]),
    "dit4sr_unpatchify": (c_int, [c_vp, c_vp, c_int, c_vp, c_float, c_int, c_int, c_int, c_int, c_vp]),
    "dit4sr_unpatchify_cfg": (c_int, [c_vp, c_vp, c_int, c_vp, c_float, c_float, c_int, c_int, c_int, c_int, c_vp]),
    "dit4sr_feature_extract": (c_int, [c_vp, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp]),
    "dit4sr_timestep_features": (c_int, [c_vp, c_vp, c_int, c_vp]),
    "dit4sr_silu_cast": (c_int, [c_vp, c_vp, c_i64, c_vp]),
    "dit4sr_add_bf16": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_int, c_int, c_int, c_vp]),
    "dit4sr_split3": (c_int, [c_vp, c_i64, c_i64, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "dit4sr_split3_transposed": (c_int, [c_vp, c_i64, c_vp, c_int, c_int, c_int, c_vp]),
    "dit4sr_softmax_rows_f32": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_float, c_vp]),
    "dit4sr_add_f32": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_int, c_int, c_int, c_vp]),
    "dit4sr_dwconv3x3_add_f32": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int,
                                         c_int, c_vp]),
    "dit4sr_patchify_f32": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp]),
    "dit4sr_timestep_features_f32": (c_int, [c_vp, c_vp, c_int, c_vp]),
    "dit4sr_silu_f32": (c_int, [c_vp, c_vp, c_i64, c_vp]),
    "dit4sr_euler_update": (c_int, [c_vp, c_vp, c_vp, c_int, c_float, c_i64, c_vp]),
    "dit4sr_wavelet_blur_f32": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp]),
    "dit4sr_plane_stats_f32": (c_int, [c_vp, c_int, c_i64, c_vp, c_vp, c_vp]),
    "dit4sr_adain_apply_f32": (c_int, [c_vp, c_vp, c_int, c_i64, c_vp, c_vp, c_vp, c_vp, c_float, c_vp]),
    "dit4sr_cast": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp]),
    "dit4sr_groupnorm_nhwc": (c_int, [C.POINTER(GroupNormParams), c_vp]),
    "dit4sr_groupnorm_workspace_bytes": (c_i64, [c_int, c_int]),
    "dit4sr_upsample2x_nhwc": (c_int, [c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "dit4sr_space_to_depth2_nhwc": (c_int, [c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "dit4sr_softmax_rows_to_bf16": (c_int, [c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_float, c_vp]),
    "dit4sr_nchw_to_nhwc_bf16": (c_int, [c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_vp]),
    "dit4sr_nhwc_to_nchw_f32": (c_int, [c_vp, c_i64, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "dit4sr_peer_alloc": (c_int, [C.c_size_t, C.POINTER(c_vp)]),
    "dit4sr_peer_free": (c_int, [c_vp]),
    "dit4sr_peer_export": (c_int, [c_vp, C.c_char_p]),
    "dit4sr_peer_open": (c_int, [C.c_char_p, C.POINTER(c_vp)]),
    "dit4sr_peer_close": (c_int, [c_vp]),
    "dit4sr_sp_barrier": (c_int, [C.POINTER(SpComm), c_int, c_vp]),
    "dit4sr_a2a_qkv": (c_int, [C.POINTER(SpComm), C.POINTER(A2aParams), c_vp]),
    "dit4sr_a2a_out": (c_int, [C.POINTER(SpComm), C.POINTER(A2aParams), c_vp]),
    "dit4sr_halo_rows": (c_int, [C.POINTER(SpComm), C.POINTER(HaloParams), c_vp]),
}

_lib = None


def load():
    """Load the library (once). Raises RuntimeError with the build hint if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not built: run `python -m dit4sr_b200.build` (or __graft_entry__.build()); "
                           "the dit4sr_b200 ops have no fallback path")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class Dit4srError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        msg = load().dit4sr_last_error().decode("utf-8", "replace")
        raise Dit4srError(f"{what} failed (rc={rc}): {msg}")
