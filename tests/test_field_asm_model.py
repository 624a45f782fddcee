"""CPU check of DEVICE code: tools/field_asm_sim.py parses ethrex_b200/csrc/field.cuh, executes the inline-PTX carry
chains of `Fe::sqr` (operand numbering, carry flags, lost-carry assertions) in Python along the exact call sequence of
the function body, and compares with big-integer arithmetic.  The GPU parity test of the same routine is
test_gpu_parity.py::test_field_mul_matches_oracle."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sqr_inline_asm_is_exact_in_simulation():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "field_asm_sim.py"), "300"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "inputs ok" in r.stdout
