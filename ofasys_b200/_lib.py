"""ctypes binding of libofab.so (include/ofab.h).  There is NO fallback: if the library is missing
or a call fails, the product path raises."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int64, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OFAB_LIB") or os.path.join(_HERE, "libofab.so")  # OFAB_LIB: an alternative build (kernel A/B runs)

F32, BF16 = 0, 1


class OfabError(RuntimeError):
    pass


class Dropout(Structure):
    """ofab_dropout (include/ofab.h): counter-based keep mask, no mask tensor."""
    _fields_ = [("state", c_void_p), ("site", c_uint32), ("p", c_float), ("drop_path", c_float), ("rows_per_sample", c_int)]


class AdamTensor(Structure):
    """ofab_adam_tensor (56 bytes; the host builds an array of these and copies it to the device)."""
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("master", c_void_p), ("m", c_void_p), ("v", c_void_p), ("n", c_int64), ("first_block", c_int64)]


class AdamHyper(Structure):
    _fields_ = [("lr", c_double), ("weight_decay", c_double), ("beta1_d", c_double), ("beta2_d", c_double),
                ("beta1", c_float), ("beta2", c_float), ("eps", c_float), ("grad_scale", c_float), ("max_norm", c_float),
                ("step", c_int64), ("norm", c_void_p)]


class CtcArgs(Structure):
    _fields_ = [("logits", c_void_p), ("dt", c_int), ("t_stride", c_int64), ("b_stride", c_int64),
                ("B", c_int), ("T", c_int), ("C", c_int), ("input_lengths", c_void_p), ("targets", c_void_p), ("Lmax", c_int),
                ("target_lengths", c_void_p), ("blank", c_int), ("zero_infinity", c_int),
                ("lse", c_void_p), ("alpha", c_void_p), ("nll", c_void_p), ("gscale", c_void_p), ("dlogits", c_void_p)]


class CeRowsArgs(Structure):
    _fields_ = [("logits", c_void_p), ("rows", c_int64), ("V", c_int64), ("ld", c_int64), ("target", c_void_p), ("ignore_index", c_int64),
                ("cmask", c_void_p), ("c_lo", c_int64), ("c_hi", c_int64), ("label_smoothing", c_float),
                ("lse", c_void_p), ("n_allowed", c_void_p), ("row_loss", c_void_p), ("row_nll", c_void_p), ("row_scale", c_void_p), ("dlogits", c_void_p)]


class AttnFwdArgs(Structure):
    _fields_ = [
        ("B", c_int), ("H", c_int), ("Tq", c_int), ("Tk", c_int),
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p),
        ("q_bs", c_int64), ("q_rs", c_int64), ("k_bs", c_int64), ("k_rs", c_int64), ("v_bs", c_int64), ("v_rs", c_int64),
        ("pq", c_void_p), ("pk", c_void_p),
        ("pq_bs", c_int64), ("pq_rs", c_int64), ("pk_bs", c_int64), ("pk_rs", c_int64),
        ("rp_idx", c_void_p), ("table", c_void_p), ("n_buckets", c_int),
        ("kpm", c_void_p), ("causal", c_int), ("scale", c_float),
        ("o", c_void_p), ("o_bs", c_int64), ("o_rs", c_int64), ("lse", c_void_p),
        ("drop", POINTER(Dropout)),
        ("rp_ld", c_int), ("rp_idx_t", c_void_p), ("rp_ld_t", c_int),
        ("bias", c_void_p), ("bias_hs", c_int64), ("bias_ld", c_int), ("bias_t", c_void_p), ("bias_t_hs", c_int64), ("bias_t_ld", c_int),
    ]


class AttnBwdArgs(Structure):
    _fields_ = [
        ("f", AttnFwdArgs),
        ("d_o", c_void_p), ("do_bs", c_int64), ("do_rs", c_int64),
        ("dq", c_void_p), ("dk", c_void_p), ("dv", c_void_p),
        ("dq_bs", c_int64), ("dq_rs", c_int64), ("dk_bs", c_int64), ("dk_rs", c_int64), ("dv_bs", c_int64), ("dv_rs", c_int64),
        ("dpq", c_void_p), ("dpk", c_void_p), ("dtable", c_void_p), ("delta", c_void_p),
        ("dq_colsum", c_void_p), ("dk_colsum", c_void_p), ("dv_colsum", c_void_p),
        ("ds", c_void_p),
    ]


class EmbedLnArgs(Structure):
    _fields_ = [
        ("B", c_int), ("T", c_int), ("d", c_int),
        ("tokens", c_void_p), ("E", c_void_p), ("dense", c_void_p), ("cls", c_void_p), ("has_cls", c_int),
        ("pos", c_void_p), ("type", c_void_p), ("gamma", c_void_p), ("beta", c_void_p),
        ("zero_mask", c_void_p), ("eps", c_float),
        ("out", c_void_p), ("out_bs", c_int64), ("mean", c_void_p), ("rstd", c_void_p),
        ("drop", POINTER(Dropout)),
    ]


class EmbedLnBwdArgs(Structure):
    _fields_ = [
        ("f", EmbedLnArgs),
        ("dout", c_void_p), ("dout_bs", c_int64), ("dE", c_void_p), ("padding_idx", c_int64),
        ("ddense", c_void_p), ("dpos", c_void_p), ("dgb_partial", c_void_p),
    ]


_SIGS = {
    "ofab_version": (c_int, []),
    "ofab_last_error": (c_char_p, []),
    "ofab_device_check": (c_int, [c_int]),
    "ofab_num_sms": (c_int, []),
    "ofab_set_pdl": (c_int, [c_int]),
    "ofab_dropout_apply": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, POINTER(Dropout), c_void_p]),
    "ofab_ln_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int, c_float, c_int, POINTER(Dropout), c_void_p]),
    "ofab_ln_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int64, c_int, c_int, POINTER(Dropout), c_void_p]),
    "ofab_ln_partial_rows": (c_int, []),
    "ofab_ln_res_ln_fwd": (c_int, [c_void_p] * 9 + [c_int64, c_int, c_float, POINTER(Dropout), c_void_p]),
    "ofab_ln_res_ln_bwd": (c_int, [c_void_p] * 10 + [c_int64, c_int, POINTER(Dropout), c_void_p]),
    "ofab_colsum": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ofab_colsum_scratch_elems": (c_int64, [c_int64]),
    "ofab_reduce_rows": (c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_int, c_void_p]),
    "ofab_reduce_partials": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "ofab_gemm_bf16": (c_int, [c_int64, c_int64, c_int64, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p]),
    "ofab_gemm_splitk_workspace_elems": (c_int64, [c_int64, c_int64, c_int64]),
    "ofab_gemm_bf16_splitk": (c_int, [c_int64, c_int64, c_int64, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p]),
    "ofab_attn_fwd": (c_int, [POINTER(AttnFwdArgs), c_void_p]),
    "ofab_attn_bwd": (c_int, [POINTER(AttnBwdArgs), c_void_p]),
    "ofab_attn_bias_build": (c_int, [c_void_p, c_int, c_float, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "ofab_attn_bias_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_float, c_int, c_void_p]),
    "ofab_embed_ln_fwd": (c_int, [POINTER(EmbedLnArgs), c_void_p]),
    "ofab_embed_ln_bwd": (c_int, [POINTER(EmbedLnBwdArgs), c_void_p]),
    "ofab_ce_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_float, c_void_p, c_void_p]),
    "ofab_ce_bwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_float, c_void_p]),
    "ofab_ce_rows_fwd": (c_int, [POINTER(CeRowsArgs), c_void_p]),
    "ofab_ce_rows_bwd": (c_int, [POINTER(CeRowsArgs), c_void_p]),
    "ofab_cast_f32_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "ofab_cast_bf16_f32": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "ofab_video_frames": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "ofab_multi_copy": (c_int, [c_void_p, c_int64, c_void_p]),
    "ofab_add_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "ofab_scale_cols": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p]),
    "ofab_scale_cols_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p]),
    "ofab_patch_im2col": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_void_p]),
    "ofab_im2col_3x3s2": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ofab_col2im_3x3s2": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ofab_conv1_relu_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "ofab_conv1_relu_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "ofab_transpose_last2": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "ofab_relu_bwd_inplace": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "ofab_relu_inplace": (c_int, [c_void_p, c_int64, c_void_p]),
    "ofab_im2col_nchw": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_void_p]),
    "ofab_im2col_nhwc": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ofab_col2im_nhwc": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ofab_subsample2": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "ofab_maxpool3x3s2_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ofab_maxpool3x3s2_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ofab_bn_scratch_elems": (c_int64, [c_int]),
    "ofab_bn_stats": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p]),
    "ofab_bn_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_int, c_void_p]),
    "ofab_bn_bwd": (c_int, [c_void_p] * 9 + [c_int64, c_int, c_float, c_int, c_void_p, c_void_p]),
    "ofab_bn_bwd_eval": (c_int, [c_void_p] * 9 + [c_int64, c_int, c_float, c_int, c_void_p, c_void_p]),
    "ofab_fbank": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_int, c_void_p]),
    "ofab_utterance_cmvn": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ofab_spec_augment": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_float, c_int, c_void_p, c_void_p]),
    "ofab_image_normalize": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float), c_void_p, c_int, c_void_p]),
    "ofab_box_bins": (c_int, [c_void_p, c_int64, c_float, c_int, c_int64, c_void_p, c_void_p]),
    "ofab_ctc_fwd": (c_int, [POINTER(CtcArgs), c_void_p]),
    "ofab_ctc_bwd": (c_int, [POINTER(CtcArgs), c_void_p]),
    "ofab_adam_chunk_elems": (c_int, []),
    "ofab_grad_norm": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "ofab_adam_step": (c_int, [c_void_p, c_int, c_int64, POINTER(AdamHyper), c_void_p]),
}

EXPORTS = tuple(_SIGS)
_lib = None


def lib():
    """Load libofab.so (built in-tree by ofasys_b200/build.py).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OfabError(
                f"{LIB_PATH} not found: build it with `python -m ofasys_b200.build` (or __graft_entry__.build()). "
                "ofasys_b200 has no CPU / eager fallback."
            )
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)  # AttributeError if the ABI and this table drift
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


launch_count = 0  # kernels-launching C-ABI calls issued by this process (bench reports it)


def check(rc, what=""):
    if rc != 0:
        msg = lib().ofab_last_error().decode(errors="replace")
        raise OfabError(f"libofab call failed ({rc}) {what}: {msg}")


_KERNELS_PER_CALL = {"ofab_colsum": 2, "ofab_attn_bwd": 2, "ofab_bn_stats": 2, "ofab_bn_bwd": 3, "ofab_bn_bwd_eval": 3, "ofab_grad_norm": 2}


def call(name, *args):
    global launch_count
    launch_count += _KERNELS_PER_CALL.get(name, 1)
    check(getattr(lib(), name)(*args), name)
