"""Generates tests/golden/groth16_toy.json: proofs of the toy Groth16 instance (tests/groth16_toy.py, log_n = 3 and 5)
computed IN THE EXPONENT by the CPU oracle, with the verifier calldata whose ecpairing check must be 1.
Run from the repo root:  python tests/golden/make_groth16_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, ".."))
import pyref  # noqa: E402
import pyref_tower as tw  # noqa: E402
from groth16_toy import ToyGroth16  # noqa: E402


def pairs(cd):
    return [(pyref.g1_from_be(cd[i:i + 64]), pyref.g2_from_be(cd[i + 64:i + 192])) for i in range(0, len(cd), 192)]


out = {"source": "tests/groth16_toy.py ToyGroth16(log_n, seed=0xB2001616); proofs computed in the exponent (no blinding)", "instances": []}
for log_n in (3, 5):
    toy = ToyGroth16(log_n)
    cases = []
    for x in (7, pyref.R - 5, 0):
        proof = toy.expected_proof(toy.assign(x))
        cd = toy.verifier_calldata(proof, x)
        assert tw.pairing_check(pairs(cd))
        cases.append({"public_input": hex(x), "proof": proof.hex(), "verifier_calldata": cd.hex()})
    out["instances"].append({"log_n": log_n, "cases": cases})
json.dump(out, open(os.path.join(HERE, "groth16_toy.json"), "w"), indent=1)
print("wrote groth16_toy.json")
