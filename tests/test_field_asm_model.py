"""CPU check of DEVICE code: tools/field_asm_sim.py parses ethrex_b200/csrc/field.cuh, executes the inline-PTX carry
chains of `Fe::mul`, `Fe::sqr`, `Fe::mul2_add` and `Fe::mul4_add` (operand numbering, carry flags, lost-carry
assertions) in Python along the translated statement sequence of each function body -- loops and conditions included --
for both moduli, and compares with big-integer arithmetic.  The GPU parity test of the same routine is
test_gpu_parity.py::test_field_mul_matches_oracle."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_field_products_inline_asm_is_exact_in_simulation():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "field_asm_sim.py"), "150"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "inputs ok" in r.stdout
