"""Every AGB_* environment switch the shipped library still reads selects a variant of a kernel or of the schedule.
Each one runs the same parity check (tests/switch_check.py: small path vs oracle under the BASELINE gate + the attention
op on the path's shapes) in its own process, so that no variant ships untested."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SWITCHES = [
    {},                                                  # defaults
    {"AGB_ZIGZAG": "0"},                                 # every kernel walks its rows in ascending order
    {"AGB_VIT_DENSE_LAST": "1"},                         # last ViT block dense instead of CLS-row only
    {"AGB_OVERLAP": "1"},                                # HuBERT chain on a side stream
    {"AGB_ATTN_STRAY_TILE": "1"},                        # odd query row as a third MMA tile instead of CUDA cores
    {"AGB_ATTN_TRACE": "1"},                             # clock64 trace of the attention pipeline (diagnostic printf)
    {"AGB_FUSED_LN": "1"},                               # LayerNorm in the residual GEMM's tail
    {"AGB_FUSED_LN": "1", "AGB_LNF_COOP": "1"},          # ... launched cooperatively
]


@pytest.mark.parametrize("switch", SWITCHES, ids=lambda s: ",".join(f"{k}={v}" for k, v in s.items()) or "defaults")
def test_switch_variant_passes_the_parity_check(switch):
    env = {k: v for k, v in os.environ.items() if not k.startswith("AGB_")}
    env.update(switch)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "switch_check.py")], env=env, cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "switch-check ok" in r.stdout
