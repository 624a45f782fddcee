"""Extracts the reference's own BN254 known-answer tests into tests/golden/pairing_kats.json.

    python tests/golden/make_pairing_kats.py        (needs /root/reference; run in the build container only)

Source: /root/reference/test/tests/levm/precompile_tests.rs:17-151 -- 14 `ecpairing` vectors (geth's
bn256Pairing.json) with their expected 32-byte boolean, plus the coordinate-out-of-range calldata of :143-151.
Only the test DATA is copied (hex strings); nothing under /root/reference is read at test time.
"""
import json
import os
import re

SRC = "/root/reference/test/tests/levm/precompile_tests.rs"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    src = open(SRC).read()
    out = {"source": "lambdaclass/ethrex test/tests/levm/precompile_tests.rs:17-151", "vectors": []}
    for m in re.finditer(r'fn (test_ec_pairing_\w+)\(\)\s*\{\s*test_ec_pairing\(\s*"([0-9a-f]*)",\s*"([0-9a-f]+)",\s*(\d+)', src):
        out["vectors"].append({"name": m.group(1), "calldata": m.group(2), "expected": int(m.group(3), 16), "gas": int(m.group(4))})
    m = re.search(r'fn test_ec_pairing_coordinate_out_of_bounds.*?hex::decode\("([0-9a-f]+)"\)', src, re.S)
    out["coordinate_out_of_bounds_calldata"] = m.group(1)
    assert len(out["vectors"]) == 14
    json.dump(out, open(os.path.join(HERE, "pairing_kats.json"), "w"), indent=1)
    print("wrote", len(out["vectors"]), "vectors")


if __name__ == "__main__":
    main()
