"""-m gpu: the opt-in / fall-back kernel variants behind environment switches.

The switches are read once per process, so every variant runs the relevant kernel tests in a child pytest process.  They are
not on the default product path (DESIGN.md section 6 lists why each one lost its A/B), but they are kept as cross-checks and
must keep producing the reference's numbers."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

K = "test_gpu_kernels.py"
VARIANTS = [
    # (environment, test file, -k expression)
    ({"SUNB_CONV_SLAB": "0"}, K, "conv3x3 and tcgen05"),              # dense 3x3 through the tap-per-K-block GEMM
    ({"SUNB_CONV_SLAB_2CTA": "0"}, K, "conv3x3 and tcgen05"),         # single-CTA slab kernel
    ({"SUNB_GEMM_2CTA": "1", "SUNB_CONV_SLAB": "0"}, K, "(gemm_plain or conv3x3 or grouped_conv_pairs) and tcgen05"),
    ({"SUNB_GEMM_BSTAT": "1", "SUNB_CONV_SLAB": "0"}, K, "(gemm_plain or conv3x3) and tcgen05"),
    ({"SUNB_GCONV": "mma"}, K, "gconv3x3"),                            # warp-MMA grouped conv
    ({"SUNB_GCONV_TMA": "1"}, K, "gconv3x3"),                          # TMA-fed grouped conv
    # the SIMT GEMM / stem behind the encoder schedule (the tests above pick the GEMM implementation explicitly)
    ({"SUNB_GEMM": "simt", "SUNB_STEM": "simt"}, "test_gpu_encoder.py", "small_episode_logits or layer_boundaries_calibrated"),
]


@pytest.mark.parametrize("env,fname,expr", VARIANTS, ids=["+".join(f"{k}={v}" for k, v in e.items()) for e, _, _ in VARIANTS])
def test_variant_matches_reference(env, fname, expr):
    child_env = dict(os.environ)
    child_env.update(env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", fname), "-q", "-x", "-k", expr,
                        "-p", "no:cacheprovider"], cwd=ROOT, env=child_env, capture_output=True, text=True, timeout=600)
    tail = (r.stdout + r.stderr)[-1500:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "no tests ran" not in r.stdout, tail
