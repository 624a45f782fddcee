"""Exploration probe: G2 MSM time and phases, plain and precomputed bases (env B200ZK_G2_MINB selects the register cap)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import ethrex_b200 as eb
import pyref

def timed(fn, iters=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]

ctx = eb.Context(0)
k, d = pyref.chain_scalar(pyref.SEED_POINTS)
for log_n in (20, 22):
    n = 1 << log_n
    p = torch.empty(16 * n, dtype=torch.int64, device="cuda"); s = torch.empty(4 * n, dtype=torch.int64, device="cuda")
    ctx.g2_chain_device(p, 0, n, k, d); ctx.fr_random_device(s, n, pyref.SEED_SCALARS, 0)
    ctx.set_profiling(True)
    ms = timed(lambda: ctx.g2_msm_device(p, s, n))
    print(json.dumps({"g2": "plain", "minb": os.environ.get("B200ZK_G2_MINB", "default"), "log_n": log_n, "ms": ms, "phases": ctx.last_msm_phase_ms()}), flush=True)
    h = ctx.g2_bases_from_device(p, n); del p
    ctx.bases_precompute(h, 0)
    ms = timed(lambda: ctx.g2_msm_resident_device(h, s, n))
    print(json.dumps({"g2": "table", "minb": os.environ.get("B200ZK_G2_MINB", "default"), "log_n": log_n, "ms": ms, "phases": ctx.last_msm_phase_ms()}), flush=True)
    ctx.set_profiling(False)
    ctx.bases_free(h); del s
