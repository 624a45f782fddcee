"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import ethrex_b200 as eb, pyref, cpu_oracle as orc

ctx = eb.Context(0)
k, d = pyref.chain_scalar(pyref.SEED_POINTS)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64).copy()).cuda()
n = 5000
pts, s = orc.g1_chain(n, k, d), orc.rand_fr(7, 0, n)
exp = orc.g1_msm(pts, s)
dp, ds = dev(pts), dev(s)
assert ctx.g1_msm_device(dp, ds, n) == exp                       # one-shot plain
ctx.set_msm_chunks(3); assert ctx.g1_msm_device(dp, ds, n) == exp; ctx.set_msm_chunks(0)   # pipelined
ctx.set_msm_pair_rounds(2); assert ctx.g1_msm_device(dp, ds, n) == exp; ctx.set_msm_pair_rounds(-1)
h = ctx.g1_bases_upload(pts, n); ctx.bases_precompute(h, 8)
assert ctx.g1_msm_resident(h, s, n) == exp                        # table + host scalars
ctx.bases_free(h)
p2 = orc.g2_chain(600, k, d)
assert ctx.g2_msm_device(dev(p2), dev(s[:600]), 600) == orc.g2_msm(p2, s[:600])
for log_n in (3, 9, 12, 14):
    a = orc.fr_to_mont(orc.rand_fr(9, 0, 1 << log_n)); da = dev(a)
    ctx.fr_ntt_device(da, log_n, 0)
    assert (da.cpu().numpy().view(np.uint64).reshape(-1, 4) == orc.fr_ntt(a, log_n)).all()
    ctx.fr_ntt_device(da, log_n, eb.NTT_INVERSE)
    ctx.fr_ntt_device(da, log_n, eb.NTT_COSET)
    ctx.fr_ntt_device(da, log_n, eb.NTT_COSET | eb.NTT_INVERSE)
    assert (da.cpu().numpy().view(np.uint64).reshape(-1, 4) == a).all()
g = torch.empty(8 * 300, dtype=torch.int64, device="cuda"); ctx.g1_chain_device(g, 0, 300, k, d); assert ctx.g1_check_device(g, 300) == 300
print("sanitize workload ok, launches:", ctx.launch_count)
ctx.close()
