"""Debug probe: multi-MSM vs single calls vs closed form at large n (precomputed tables)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import ethrex_b200 as eb, pyref, cpu_oracle as orc
from helpers import expected_chain_msm_g1, expected_chain_msm_g2, to_dev
ctx = eb.Context(0)
for log_n in (16, 20, 22):
    n = 1 << log_n
    cols = []
    for tag, g2 in ((3, False), (5, False), (7, True), (11, False)):
        k, d = 12345 * tag + 1, 777 + tag
        pts = torch.empty((16 if g2 else 8) * n, dtype=torch.int64, device="cuda")
        (ctx.g2_chain_device if g2 else ctx.g1_chain_device)(pts, 0, n, k, d)
        h = (ctx.g2_bases_from_device if g2 else ctx.g1_bases_from_device)(pts, n)
        del pts
        ctx.bases_precompute(h, 0)
        cols.append((h, g2, k, d))
    s = orc.rand_fr(0xB2000077, 0, n)
    ds = to_dev(s)
    exp = [(expected_chain_msm_g2 if g2 else expected_chain_msm_g1)(s, k, d) for (h, g2, k, d) in cols]
    single = [(ctx.g2_msm_resident_device if g2 else ctx.g1_msm_resident_device)(h, ds, n) for (h, g2, k, d) in cols]
    multi = ctx.msm_multi_resident_device([c[0] for c in cols], [c[1] for c in cols], ds, n)
    multi2 = ctx.msm_multi_resident_device([c[0] for c in cols], [c[1] for c in cols], ds, n)
    print(log_n, "single==exp", [a == b for a, b in zip(single, exp)], "multi==exp", [a == b for a, b in zip(multi, exp)], "multi2==exp", [a == b for a, b in zip(multi2, exp)])
    for c in cols:
        ctx.bases_free(c[0])
