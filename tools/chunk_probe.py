import os, sys, json, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ethrex_b200 as eb, pyref, cpu_oracle as orc
import numpy as np
ctx = eb.Context(0)
k, dd = pyref.chain_scalar(pyref.SEED_POINTS)
def timed(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return sorted(ts)[len(ts) // 2]
for log_n in (24,):
    n = 1 << log_n
    p = torch.empty(8 * n, dtype=torch.int64, device="cuda"); s = torch.empty(4 * n, dtype=torch.int64, device="cuda")
    ctx.g1_chain_device(p, 0, n, k, dd); ctx.fr_random_device(s, n, pyref.SEED_SCALARS, 0)
    hs = torch.empty(4 * n, dtype=torch.int64).pin_memory(); hs.copy_(s)
    dot = orc.chain_dot(hs.numpy().view(np.uint64).reshape(n, 4), k, dd); exp = orc.g1_mul_be(pyref.g1_to_be(pyref.G1_GEN), dot.to_bytes(32, "big"))[1]
    h = ctx.g1_bases_from_device(p, n); ctx.bases_precompute(h, 0)
    for K in (1, 2, 4, 6):
        ctx.set_msm_chunks(K)
        ok1 = ctx.g1_msm_resident_device(h, s, n) == exp
        ok2 = ctx.g1_msm_resident(h, hs, n) == exp
        ok3 = ctx.g1_msm_device(p, s, n) == exp
        print(json.dumps({"log_n": log_n, "chunks": K, "ok": [ok1, ok2, ok3],
                          "resident_ms": round(timed(lambda: ctx.g1_msm_resident_device(h, s, n)), 3),
                          "e2e_host_scalars_ms": round(timed(lambda: ctx.g1_msm_resident(h, hs, n)), 3),
                          "plain_ms": round(timed(lambda: ctx.g1_msm_device(p, s, n)), 3)}), flush=True)
    ctx.set_msm_chunks(0)
    ctx.bases_free(h); del p, s, hs
