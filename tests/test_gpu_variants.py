"""Every kernel variant behind the C ABI is parity-checked, not only the default dispatch: the generic dual-number sweep
(the fallback of every model), the thread-per-node sweep with vector duals, and its hand-structured node Jacobians."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [
    {"UNGAR_B200_FORCE_GENERIC": "1"},
    {"UNGAR_B200_FORCE_TPN": "1", "UNGAR_B200_TPN_STRUCT": "0"},
    {"UNGAR_B200_FORCE_TPN": "1", "UNGAR_B200_TPN_STRUCT": "1"},
], ids=["generic", "tpn-vector-duals", "tpn-structured"])
def test_kernel_variant_matches_oracle(oracle, env):
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "variant_parity_script.py")],
                          env=dict(os.environ, PYTHONPATH=ROOT, **env), capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert proc.stdout.count("OK ") == 6, proc.stdout
