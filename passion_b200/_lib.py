"""ctypes binding of libpassion_b200.so (the C ABI declared in include/passion_b200.h).

There is deliberately NO fallback: if the shared library is missing or a symbol is absent the
import of any op raises, and every op refuses non-CUDA tensors.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpassion_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "passion_b200.h")

PB_F32, PB_BF16 = 0, 1
PB_PAD_ZERO, PB_PAD_REFLECT = 0, 1

_lib = None


class ConvDesc(ctypes.Structure):
    """Mirror of pb_conv_desc."""
    _fields_ = [(n, ctypes.c_int32) for n in (
        "dtype", "n", "di", "hi", "wi", "dout", "ho", "wo", "c0", "c1", "cout", "ksize", "stride",
        "pad_mode", "groups")]


class WeightPrepDesc(ctypes.Structure):
    """Mirror of pb_weight_prep_desc."""
    _fields_ = [("w", ctypes.c_void_p * 4), ("b", ctypes.c_void_p * 4),
                ("groups", ctypes.c_int32), ("cin", ctypes.c_int32), ("cout", ctypes.c_int32), ("ksize", ctypes.c_int32),
                ("wk", ctypes.c_void_p), ("wt", ctypes.c_void_p), ("img", ctypes.c_void_p), ("nt", ctypes.c_int32),
                ("imgT", ctypes.c_void_p), ("ntT", ctypes.c_int32), ("bias", ctypes.c_void_p), ("w_cin_stride", ctypes.c_int32)]


class WeightUnpackDesc(ctypes.Structure):
    """Mirror of pb_weight_unpack_desc."""
    _fields_ = [("dw", ctypes.c_void_p), ("db", ctypes.c_void_p), ("dy_stats", ctypes.c_void_p), ("npg", ctypes.c_int32),
                ("gw", ctypes.c_void_p * 4), ("gb", ctypes.c_void_p * 4),
                ("groups", ctypes.c_int32), ("cin", ctypes.c_int32), ("cout", ctypes.c_int32), ("ksize", ctypes.c_int32),
                ("accumulate", ctypes.c_int32), ("w_cin_stride", ctypes.c_int32)]


class AugmentSample(ctypes.Structure):
    """Mirror of pb_augment_sample (112 bytes; checked against pb_augment_sample_size() at load)."""
    _fields_ = [("vol", ctypes.c_void_p), ("seg", ctypes.c_void_p), ("shape", ctypes.c_int32 * 3), ("start", ctypes.c_int32 * 3),
                ("flip", ctypes.c_int32 * 3), ("rot_axes", ctypes.c_int32 * 2), ("_pad", ctypes.c_int32),
                ("rot_m", ctypes.c_double * 4), ("rot_off", ctypes.c_double * 2)]


def declared_symbols():
    """Every function name declared in include/passion_b200.h."""
    with open(HEADER_PATH) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", src)))


def load():
    """dlopen the library and check that it exports every symbol of the header."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"passion_b200: CUDA library not built ({LIB_PATH}); run `python __graft_entry__.py`. "
            "There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    if missing:
        raise RuntimeError(f"passion_b200: library is missing symbols {missing}")
    lib.pb_last_error.restype = ctypes.c_char_p
    lib.pb_launch_count.restype = ctypes.c_longlong
    lib.pb_weight_batch_table_bytes.restype = ctypes.c_size_t
    lib.pb_gemm_tc_workspace_floats.restype = ctypes.c_longlong
    lib.pb_gemm_tc_workspace_floats.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.pb_weight_batch_table_bytes.argtypes = [ctypes.c_int]
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float
    cd = ctypes.POINTER(ConvDesc)
    sig = {
        "pb_conv3d_fwd": [cd, vp, vp, vp, vp, vp, vp, vp],
        "pb_conv3d_dgrad": [cd, vp, vp, vp, vp, vp],
        "pb_conv3d_wgrad": [cd, vp, vp, vp, vp, vp],
        "pb_conv3d_tc_ntile": [i32, i32],
        "pb_conv3d_tc_kws": [i32, i32],
        "pb_conv1_tc_ntile": [i32, i32],
        "pb_conv1_tc": [cd, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp],
        "pb_conv3d_tc": [cd, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp],
        "pb_conv3d_tc_full": [cd, vp, vp, vp, vp, i32, i32, vp, vp, vp],
        "pb_reflect_fold": [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp],
        "pb_conv3d_wgrad_tc": [cd, vp, vp, vp, vp, vp, vp],
        "pb_gemm_tc": [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp],
        "pb_gemm_tc_batched": [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i64, i64, i64, i64, i64, i64, vp, vp],
        "pb_attn_softmax_fwd": [vp, vp, vp, i64, i32, i32, f32, f32, vp, vp],
        "pb_attn_softmax_bwd": [vp, vp, vp, vp, i64, i32, i32, f32, vp],
        "pb_attn_scores_softmax": [vp, vp, vp, vp, i32, i32, i32, i32, i32, i64, i64, i32, f32, f32, vp, vp, vp],
        "pb_attn_delta": [vp, vp, vp, i32, i32, i32, i32, vp],
        "pb_attn_dsoftmax": [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i64, i64, i32, i64, i64, i32, f32, vp, vp],
        "pb_layernorm_fwd": [i32, vp, vp, vp, vp, vp, vp, i64, i32, f32, vp],
        "pb_layernorm_bwd": [i32, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, vp],
        "pb_conv1_wgrad_tc": [cd, vp, vp, vp, vp, vp, vp],
        "pb_conv3d_dgrad_reflect_fix": [cd, vp, vp, vp, vp, vp],
        "pb_conv3d_small_supported": [i32, i32],
        "pb_conv3d_small_fwd": [cd, vp, vp, vp, vp, vp, vp],
        "pb_conv3d_small_dgrad": [cd, vp, vp, vp, vp, vp],
        "pb_conv3d_small_wgrad": [cd, vp, vp, vp, vp],
        "pb_weight_prep": [ctypes.POINTER(WeightPrepDesc), vp],
        "pb_weight_grad_unpack": [ctypes.POINTER(WeightUnpackDesc), vp],
        "pb_weight_prep_batch": [ctypes.POINTER(WeightPrepDesc), i32, vp, i32, vp],
        "pb_weight_grad_unpack_batch": [ctypes.POINTER(WeightUnpackDesc), i32, vp, i32, vp],
        "pb_channel_stats": [i32, vp, vp, i32, i64, i32, vp],
        "pb_inorm_finalize": [vp, vp, i32, i32, i64, f32, vp],
        "pb_inorm_lrelu_fwd": [i32, vp, vp, vp, vp, i32, i64, i32, f32, vp],
        "pb_inorm_lrelu_fwd_stats": [i32, vp, vp, vp, vp, vp, i32, i64, i32, f32, f32, vp],
        "pb_inorm_lrelu_bwd": [i32, vp, vp, vp, vp, vp, i32, i64, i32, f32, vp],
        "pb_upsample_fwd": [i32, vp, vp, i32, i32, i32, i32, i32, i32, vp],
        "pb_upsample_bwd": [i32, vp, vp, i32, i32, i32, i32, i32, i32, vp],
        "pb_upsample_bwd_axis": [i32, vp, vp, i64, i32, i32, i64, i32, vp],
        "pb_softmax4": [i32, vp, vp, i64, f32, vp],
        "pb_softmax4_bwd": [i32, vp, vp, vp, i64, f32, vp],
        "pb_logit_loss_fwd": [i32, vp, vp, vp, vp, vp, i32, i32, i64, i32, f32, vp],
        "pb_logit_loss_bwd": [i32, vp, vp, vp, vp, vp, vp, i32, i32, i64, i32, f32, vp],
        "pb_cedice_fwd": [vp, vp, vp, i32, i32, i64, vp],
        "pb_cedice_bwd": [vp, vp, vp, vp, i32, i32, i64, vp],
        "pb_kl_fwd": [vp, vp, vp, i32, i32, i64, vp],
        "pb_kl_bwd": [vp, vp, vp, vp, i32, i32, i64, vp],
        "pb_proto_sums": [i32, vp, vp, vp, i32, i32, i64, i32, vp],
        "pb_proto_fwd": [i32, vp, vp, vp, vp, vp, vp, i32, i32, i64, i32, f32, vp],
        "pb_proto_bwd1": [i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i64, i32, f32, vp],
        "pb_proto_bwd2": [i32, vp, vp, vp, i32, i32, i64, i32, vp],
        "pb_masked_stack_fwd": [i32, vp, vp, vp, i32, i32, i64, i32, vp],
        "pb_masked_stack_bwd": [i32, vp, vp, vp, i32, i32, i64, i32, vp],
        "pb_rfm_pool": [i32, vp, vp, vp, vp, i32, i64, i32, vp],
        "pb_rfm_gate_fwd": [vp, vp, vp, vp, vp, i32, i32, i64, i32, i32, vp],
        "pb_rfm_gate_bwd": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i64, i32, i32, vp],
        "pb_rfm_mix": [i32, vp, vp, vp, vp, i32, i64, i32, i32, vp],
        "pb_rfm_mix_bwd_gate": [i32, vp, vp, vp, vp, i32, i64, i32, i32, vp],
        "pb_rfm_bwd_y": [i32, vp, vp, vp, vp, vp, i32, i64, i32, i32, vp],
        "pb_augment_sample_size": [],
        "pb_augment_batch": [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp],
    }
    for name, args in sig.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = ctypes.c_int
    if lib.pb_augment_sample_size() != ctypes.sizeof(AugmentSample):
        raise RuntimeError("passion_b200: pb_augment_sample layout mismatch between the header and _lib.AugmentSample")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"passion_b200.{what} failed ({rc}): {load().pb_last_error().decode()}")


def launch_count():
    return int(load().pb_launch_count())
