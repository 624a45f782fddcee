import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ethrex_b200 as eb, pyref
ctx = eb.Context(0)
k, dd = pyref.chain_scalar(pyref.SEED_POINTS)
n = 1 << 22
p = torch.empty(8 * n, dtype=torch.int64, device="cuda"); s = torch.empty(4 * n, dtype=torch.int64, device="cuda")
ctx.g1_chain_device(p, 0, n, k, dd); ctx.fr_random_device(s, n, pyref.SEED_SCALARS, 0)
h = ctx.g1_bases_from_device(p, n); ctx.bases_precompute(h, 18)
ctx.set_msm_pair_rounds(2)
for _ in range(2):
    ctx.g1_msm_resident_device(h, s, n)
