import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle')); 
import numpy as np, torch
import ethrex_b200 as eb, pyref, cpu_oracle as orc
ctx = eb.Context(0)
g = pyref.g1_to_be(pyref.G1_GEN)
R = pyref.R
for a, b in ((5, R - 5), (3, 4), (1, R - 1), (5, R - 4), (R - 5, 5), (7, R - 7), (2**200, R - 2**200)):
    sc = a.to_bytes(32, 'big') + b.to_bytes(32, 'big')
    try:
        out = ctx.g1_msm(g + g, sc, 2, eb.POINTS_BE | eb.SCALARS_BE)
        exp = pyref.g1_to_be(pyref.g1_mul((a + b) % R, pyref.G1_GEN))
        print(hex(a)[:12], hex(b)[:12], 'ok' if out == exp else 'MISMATCH', out.hex()[:16], exp.hex()[:16])
    except Exception as e:
        print(hex(a)[:12], hex(b)[:12], 'EXC', e)
