"""ctypes binding of libsunb200.so (C ABI declared in include/sunb200.h).

The shared library is built in-tree by ``make -C few-shot-vit_b200/csrc`` (``__graft_entry__.build()``).
There is no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SUNB200_LIB: developer override to A/B two builds of the same ABI on one box
LIB_PATH = os.environ.get("SUNB200_LIB") or os.path.join(_HERE, "libsunb200.so")

vp = C.c_void_p
fp = C.c_void_p      # float* passed as raw address


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("taps", C.c_int32), ("groups", C.c_int32),
        ("a_goff", C.c_int32), ("c_goff", C.c_int32),
        ("a_mode", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("bw", C.c_int32), ("bh", C.c_int32),
        ("A", vp), ("lda", C.c_int32),
        ("Wt", vp), ("ldw", C.c_int32),
        ("bias", fp), ("bias_mod", C.c_int32), ("bias_ld", C.c_int32),
        ("act", C.c_int32),
        ("resid", vp), ("ldr", C.c_int32),
        ("row_scale", fp), ("rows_per_img", C.c_int32),
        ("out", vp), ("ldc", C.c_int32),
        ("out_f32", fp), ("ldc_f32", C.c_int32),
        ("out_map", C.c_int32), ("oH", C.c_int32), ("oW", C.c_int32),
        ("out2", vp), ("ldc2", C.c_int32),
        ("dact_aux", vp), ("ld_aux", C.c_int32), ("dact", C.c_int32),
    ]


class WgradDesc(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("Ma", C.c_int32), ("Nb", C.c_int32), ("Ca", C.c_int32), ("Cb", C.c_int32),
        ("groups", C.c_int32), ("a_goff", C.c_int32), ("b_goff", C.c_int32),
        ("taps", C.c_int32),
        ("mode", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("bw", C.c_int32), ("bh", C.c_int32),
        ("dY", vp), ("ldy", C.c_int32),
        ("X", vp), ("ldx", C.c_int32),
        ("out", fp), ("ldo", C.c_int32),
        ("ksplit", C.c_int32),
    ]


class OptTensor(C.Structure):
    _fields_ = [("p", vp), ("g", vp), ("m", vp), ("v", vp), ("n", C.c_int64)]


class PackDesc(C.Structure):
    _fields_ = [("src", vp), ("dst", vp), ("off", C.c_int64), ("strides", C.c_int32 * 4), ("dims", C.c_int32 * 4),
                ("ldd", C.c_int32), ("valid2", C.c_int32)]


class ConvMlpW(C.Structure):
    _fields_ = [("w1", vp), ("b1", fp), ("w2", vp), ("w3", vp), ("w23", vp)]


class AttnBlockW(C.Structure):
    _fields_ = [("wqkv", vp), ("bqkv", fp), ("wproj", vp), ("w1", vp), ("b1", fp), ("w3", vp)]


class EncoderWeights(C.Structure):
    _fields_ = [
        ("stem_w1", fp), ("stem_b1", fp), ("stem_wd", fp), ("stem_bd", fp),
        ("stem_w2", vp), ("stem_b2", fp), ("stem_w3", vp), ("stem_b3", fp),
        ("pos1", fp),
        ("s1", ConvMlpW * 4),
        ("pe2_w", vp), ("pe2_bias", fp),
        ("s2", AttnBlockW * 2),
        ("pe3_w", vp), ("pe3_bias", fp),
        ("s3", AttnBlockW * 3),
        ("final_scale", fp), ("final_shift", fp),
    ]


class EncoderTaps(C.Structure):
    _fields_ = [("stem", vp), ("stage1", vp * 4), ("patch_embed2", vp), ("stage2", vp * 2),
                ("patch_embed3", vp), ("stage3", vp * 3)]


# every symbol include/sunb200.h declares, with its ctypes signature
SIGNATURES = {
    "sunb_abi_version": (C.c_int, []),
    "sunb_last_error": (C.c_char_p, []),
    "sunb_gemm": (C.c_int, [C.POINTER(GemmDesc), C.c_int, vp]),
    "sunb_encoder_workspace_bytes": (C.c_int, [C.c_int, C.POINTER(C.c_size_t)]),
    "sunb_encoder_forward": (C.c_int, [C.POINTER(EncoderWeights), fp, C.c_int, vp, C.c_size_t, fp, fp, vp, vp,
                                       C.POINTER(EncoderTaps), vp]),
    "sunb_attention": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "sunb_episode_logits": (C.c_int, [fp, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp, C.c_float, vp]),
    "sunb_logits_ce_acc": (C.c_int, [fp, vp, C.c_int, C.c_int, fp, vp]),
    "sunb_hard_ce_backward": (C.c_int, [fp, vp, C.c_int, C.c_int, fp, C.c_float, fp, vp]),
    "sunb_softlabel": (C.c_int, [fp, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_double, fp, vp]),
    "sunb_soft_ce_forward": (C.c_int, [fp, C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp, vp]),
    "sunb_soft_ce_backward": (C.c_int, [fp, C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_int, fp, C.c_float, fp,
                                        C.c_int, vp]),
    # train-mode building blocks
    "sunb_wgrad": (C.c_int, [C.POINTER(WgradDesc), vp]),
    "sunb_stem_in": (C.c_int, [fp, fp, fp, fp, fp, vp, vp, C.c_int, C.c_int, vp]),
    "sunb_stem_wgrad": (C.c_int, [fp, vp, vp, fp, fp, C.c_int, vp, vp]),
    "sunb_colstats": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_long, C.c_int, fp, fp, vp]),
    "sunb_bn_finalize": (C.c_int, [fp, fp, C.c_float, fp, fp, fp, fp, vp, C.c_float, C.c_float, C.c_int, fp, fp, fp, fp, vp]),
    "sunb_bn_stats_forward": (C.c_int, [vp, C.c_int, C.c_long, C.c_int, fp, fp, vp, fp, fp, fp, fp, vp, C.c_float, C.c_float,
                                        fp, fp, fp, fp, vp]),
    "sunb_bn_stats_backward": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_long, C.c_int, fp, fp, vp, C.c_float, fp, fp, fp, C.c_int,
                                         fp, fp, fp, fp, fp, vp]),
    "sunb_bn_apply": (C.c_int, [vp, C.c_int, fp, fp, C.c_int, fp, C.c_int, vp, C.c_int, C.c_long, C.c_int, vp]),
    "sunb_bn_frozen": (C.c_int, [fp, fp, fp, fp, C.c_float, C.c_int, fp, fp, fp, fp, vp]),
    "sunb_bn_bwd_finalize": (C.c_int, [fp, fp, C.c_float, fp, fp, fp, C.c_int, C.c_int, fp, fp, fp, fp, fp, vp]),
    "sunb_bn_bwd_apply": (C.c_int, [vp, C.c_int, vp, C.c_int, fp, fp, fp, fp, vp, C.c_int, vp, C.c_int, C.c_long, C.c_int, vp]),
    "sunb_stem_tail_forward": (C.c_int, [vp, vp, fp, fp, fp, fp, fp, vp, C.c_int, vp]),
    "sunb_stem_tail_backward": (C.c_int, [vp, vp, fp, fp, fp, fp, vp, vp, C.c_int, vp]),
    "sunb_final_norm_pool": (C.c_int, [vp, fp, fp, fp, vp, fp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "sunb_pool_backward": (C.c_int, [fp, fp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "sunb_scale_rows": (C.c_int, [vp, fp, C.c_int, vp, C.c_long, C.c_int, vp]),
    "sunb_s2d_reorder": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "sunb_batch_sum": (C.c_int, [vp, C.c_int, C.c_long, fp, vp]),
    "sunb_pack_weights": (C.c_int, [C.POINTER(PackDesc), C.c_int, vp]),
    "sunb_permute_cast": (C.c_int, [fp, C.c_long, C.c_long, C.c_long, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "sunb_grouped_pairs": (C.c_int, [fp, vp, C.c_int, vp]),
    "sunb_grouped_wgrad_extract": (C.c_int, [fp, fp, vp]),
    "sunb_gconv3x3": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "sunb_gconv_pack": (C.c_int, [fp, vp, C.c_int, vp]),
    "sunb_convmlp_tail": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, vp]),
    "sunb_layernorm_rows": (C.c_int, [fp, fp, fp, fp, C.c_long, C.c_int, C.c_float, vp]),
    "sunb_mlp_fused": (C.c_int, [vp, vp, fp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "sunb_attention_backward": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "sunb_episode_logits_backward": (C.c_int, [fp, fp, fp, fp, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.c_int, fp, C.c_float, vp]),
    "sunb_emd_head": (C.c_int, [fp, fp, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, vp]),
    "sunb_preprocess_u8": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, fp, fp, vp]),
    # fused multi-tensor optimizers
    "sunb_fused_sgd": (C.c_int, [C.POINTER(OptTensor), C.c_int, fp, vp]),
    "sunb_fused_adamw": (C.c_int, [C.POINTER(OptTensor), C.c_int, fp, vp]),
}

_lib = None
ABI_VERSION = 7


def lib() -> C.CDLL:
    """Load libsunb200.so once; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
                "g.build()' or make -C few-shot-vit_b200/csrc).  sunb200 has no CPU or PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        if l.sunb_abi_version() != ABI_VERSION:
            raise RuntimeError("libsunb200.so ABI version mismatch; rebuild")
        _lib = l
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().sunb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed with status {status}: {msg}")


def ptr(t) -> int:
    """Raw device address of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("sunb200 kernels run on CUDA tensors only (sm_100a); there is no CPU path")
