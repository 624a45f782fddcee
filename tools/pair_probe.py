"""probe: pair-summing rounds x window for the precomputed-table MSM"""
import os, sys, json, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ethrex_b200 as eb, pyref, cpu_oracle as orc
import numpy as np
ctx = eb.Context(0)
k, dd = pyref.chain_scalar(pyref.SEED_POINTS)
def timed(fn, iters=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
for log_n, cs in ((20, (20,)), (24, (20,))):
    n = 1 << log_n
    p = torch.empty(8 * n, dtype=torch.int64, device="cuda"); s = torch.empty(4 * n, dtype=torch.int64, device="cuda")
    ctx.g1_chain_device(p, 0, n, k, dd); ctx.fr_random_device(s, n, pyref.SEED_SCALARS, 0)
    sh = s.cpu().numpy().view(np.uint64).reshape(n, 4)
    dot = orc.chain_dot(sh, k, dd); exp = orc.g1_mul_be(pyref.g1_to_be(pyref.G1_GEN), dot.to_bytes(32, "big"))[1]
    ctx.set_profiling(True)
    for c in cs:
        h = ctx.g1_bases_from_device(p, n); ctx.bases_precompute(h, c)
        for rounds in (0, 1, 2, 3):
            ctx.set_msm_pair_rounds(rounds)
            try:
                out = ctx.g1_msm_resident_device(h, s, n)
                ms = timed(lambda: ctx.g1_msm_resident_device(h, s, n))
                print(json.dumps({"log_n": log_n, "c": c, "rounds": rounds, "ok": out == exp, "ms": round(ms, 3), "phases": {a: round(b, 2) for a, b in ctx.last_msm_phase_ms().items()}}), flush=True)
            except Exception as e:
                print(json.dumps({"log_n": log_n, "c": c, "rounds": rounds, "error": str(e)}), flush=True)
        ctx.bases_free(h)
    ctx.set_msm_pair_rounds(-1)
    # plain (non-table) path with rounds
    for rounds in (0, 1, 2):
        ctx.set_msm_pair_rounds(rounds)
        out = ctx.g1_msm_device(p, s, n)
        ms = timed(lambda: ctx.g1_msm_device(p, s, n))
        print(json.dumps({"log_n": log_n, "plain": True, "rounds": rounds, "ok": out == exp, "ms": round(ms, 3), "phases": {a: round(b, 2) for a, b in ctx.last_msm_phase_ms().items()}}), flush=True)
    ctx.set_msm_pair_rounds(-1)
    del p, s
