"""The optional GEMM feeding paths of the conv kernel -- CTA pairs (tcgen05 cta_group::2, DIRB200_CTA2=1) and the
im2col-mode TMA A operand (DIRB200_IM2COL=1) -- against torch on the same shapes as test_gpu_conv (plus full-size
layers).  The switches are read once per process, so each variant runs tests/cta2_check.py in a subprocess."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{"DIRB200_CTA2": "1"}, {"DIRB200_IM2COL": "1"}, {"DIRB200_ATMA": "0"}],
                         ids=["cta_pairs", "im2col_tma", "cp_async_gather_only"])
def test_conv_variant_parity(env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "cta2_check.py"), "parity"], env=e,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "11/11 ok" in r.stdout
