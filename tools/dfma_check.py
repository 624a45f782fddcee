"""Big-integer check of tools/dfma_mul_probe.cu's host path: every line "a b r s" must satisfy
r == a*b*2^-260 mod p and s == a*a*2^-260 mod p with r, s < p.   tools/build/dfma_mul_probe cpu 10000 | python tools/dfma_check.py"""
import sys
P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
RINV = pow(1 << 260, -1, P)
n = bad = 0
for line in sys.stdin:
    a, b, r, s = (int(x, 16) for x in line.split())
    n += 1
    if r != a * b * RINV % P or s != a * a * RINV % P:
        bad += 1
        if bad < 5:
            print("MISMATCH", line.strip())
print(f"{n} vectors, {bad} mismatches")
sys.exit(1 if bad or not n else 0)
