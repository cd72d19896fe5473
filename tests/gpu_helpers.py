"""Shared helpers for the -m gpu tests: build SunbGemmDesc calls and torch fp32 references on the device."""
import ctypes as C

import torch
import torch.nn.functional as F

import os

from sunb200 import native as N

_CHECK = None


def check_lib():
    """tests/native/libsunb200_check.so: SIMT GEMM, CUDA-core stem, warp-MMA grouped conv and warp-MMA attention (forward /
    backward, any head stride incl. the reference's packed layout) cross-check kernels (built by
    `make -C few-shot-vit_b200/csrc`; test-only, never linked into the product library)."""
    global _CHECK
    if _CHECK is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native", "libsunb200_check.so")
        l = C.CDLL(path)
        l.sunb_check_last_error.restype = C.c_char_p
        l.sunb_check_gemm.restype, l.sunb_check_gemm.argtypes = C.c_int, [C.POINTER(N.GemmDesc), C.c_void_p]
        vp = C.c_void_p
        l.sunb_check_stem_in.restype = C.c_int
        l.sunb_check_stem_in.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp]
        l.sunb_check_attention.restype = C.c_int
        l.sunb_check_attention.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        l.sunb_check_attention_backward.restype = C.c_int
        l.sunb_check_attention_backward.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        l.sunb_check_gconv3x3.restype = C.c_int
        l.sunb_check_gconv3x3.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        _CHECK = l
    return _CHECK


def check_call(status, what):
    if status != 0:
        raise RuntimeError(f"{what} failed: {check_lib().sunb_check_last_error().decode()}")


def run_gemm(A, Wt, M, Nn, K, impl=0, taps=1, groups=1, a_goff=0, c_goff=0, conv=None, bias=None, bias_mod=1,
             act=0, resid=None, row_scale=None, rows_per_img=1, out_cols=None, out_f32=False, out_map=0, oHW=(0, 0),
             out=None):
    """A bf16 [rows, lda] (or NHWC [B,H,W,lda]); Wt bf16 [groups*taps*N, ldw].  Returns the output tensor."""
    dev = A.device
    lda = A.shape[-1]
    out_cols = out_cols or (groups * c_goff if groups > 1 else Nn)
    d = N.GemmDesc()
    d.M, d.N, d.K, d.taps, d.groups, d.a_goff, d.c_goff = M, Nn, K, taps, groups, a_goff, c_goff
    if conv:
        d.a_mode, d.H, d.W, d.bw, d.bh = 1, conv[0], conv[1], conv[2], conv[3]
    d.A, d.lda, d.Wt, d.ldw = A.data_ptr(), lda, Wt.data_ptr(), Wt.shape[-1]
    d.bias = N.ptr(bias)
    d.bias_mod, d.bias_ld = bias_mod, (bias.shape[-1] if bias is not None and bias_mod > 1 else 0)
    d.act = act
    d.resid, d.ldr = N.ptr(resid), (resid.shape[-1] if resid is not None else 0)
    d.row_scale, d.rows_per_img = N.ptr(row_scale), rows_per_img
    if out is None:
        out = torch.zeros(M, out_cols, dtype=torch.float32 if out_f32 else torch.bfloat16, device=dev)
    if out_f32:
        d.out_f32, d.ldc_f32 = out.data_ptr(), out.shape[-1]
    else:
        d.out, d.ldc = out.data_ptr(), out.shape[-1]
    d.out_map, d.oH, d.oW = out_map, oHW[0], oHW[1]
    if impl == 1:          # SIMT cross-check kernel of the test-only library
        check_call(check_lib().sunb_check_gemm(C.byref(d), N.current_stream()), "sunb_check_gemm")
    else:
        N.check(N.lib().sunb_gemm(C.byref(d), 0, N.current_stream()), "sunb_gemm")
    torch.cuda.synchronize()
    return out


def act_ref(v, act):
    if act == 1:
        return F.leaky_relu(v, 0.1)
    if act == 2:
        return 0.5 * v * (1 + torch.erf(v * 0.70710678118654752))
    return v


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def max_err(a, b):
    return (a.float() - b.float()).abs().max().item()
