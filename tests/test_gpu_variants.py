"""-m gpu: the product kernels against the independent cross-check kernels of tests/native/libsunb200_check.so
(CUDA-core stem entry convolutions, warp-MMA grouped 3x3 convolution; the SIMT GEMM is used by test_gpu_kernels.py).
The product library itself has a single implementation per operation and no run-time switches."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_helpers import check_lib, check_call, rel_err  # noqa: E402
from sunb200 import native as N, packing  # noqa: E402

DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


@pytest.mark.parametrize("B,lrelu", [(3, 1), (5, 0)])
def test_stem_in_tcgen05_vs_cuda_core(B, lrelu):
    x = rnd(B, 3, 80, 80, seed=1)
    w1, b1 = rnd(64, 27, seed=2, scale=0.2), rnd(64, seed=3, scale=0.1)
    wd, bd = rnd(128, 27, seed=4, scale=0.2), rnd(128, seed=5, scale=0.1)
    outs = []
    for which in ("tc", "check"):
        a1 = torch.zeros(B * 1600, 64, device=DEV, dtype=torch.bfloat16)
        idn = torch.zeros(B * 1600, 128, device=DEV, dtype=torch.bfloat16)
        args = (x.data_ptr(), w1.data_ptr(), b1.data_ptr(), wd.data_ptr(), bd.data_ptr(), a1.data_ptr(), idn.data_ptr(), B, lrelu,
                N.current_stream())
        if which == "tc":
            N.check(N.lib().sunb_stem_in(*args), "sunb_stem_in")
        else:
            check_call(check_lib().sunb_check_stem_in(*args), "sunb_check_stem_in")
        torch.cuda.synchronize()
        outs.append((a1, idn))
    # the tensor-core kernel rounds its im2col operand to bf16 (K = 27); the CUDA-core checker keeps fp32 inputs
    assert rel_err(outs[0][0], outs[1][0]) < 1e-2
    assert rel_err(outs[0][1], outs[1][1]) < 1e-2


def test_gconv_tcgen05_vs_warp_mma():
    B = 4
    x = rnd(B * 400, 256, seed=40).bfloat16()
    w = rnd(256, 32, 3, 3, seed=41, scale=(9 * 32) ** -0.5)
    wg = packing._grouped_taps(w).bfloat16().contiguous()
    aux = rnd(B * 400, 256, seed=43).bfloat16()
    for act, dact, use_aux in ((2, 0, False), (0, 2, True)):
        ys = []
        for which in ("tc", "check"):
            y = torch.zeros(B * 400, 256, device=DEV, dtype=torch.bfloat16)
            y2 = torch.zeros_like(y)
            args = (x.data_ptr(), 256, wg.data_ptr(), y.data_ptr(), 256, None if use_aux else y2.data_ptr(), 0 if use_aux else 256,
                    aux.data_ptr() if use_aux else None, 256 if use_aux else 0, B, act, dact, N.current_stream())
            if which == "tc":
                N.check(N.lib().sunb_gconv3x3(*args), "sunb_gconv3x3")
            else:
                check_call(check_lib().sunb_check_gconv3x3(*args), "sunb_check_gconv3x3")
            torch.cuda.synchronize()
            ys.append((y, y2))
        assert rel_err(ys[0][0], ys[1][0]) < 5e-3
        if not use_aux:
            assert rel_err(ys[0][1], ys[1][1]) < 5e-3
