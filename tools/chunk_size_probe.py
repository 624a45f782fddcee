"""Exploration probe: bucket-reduction time against the running-sum chunk size (env B200ZK_CHUNK), G1 2^24 table and G2 2^22 table."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import ethrex_b200 as eb
import pyref
ctx = eb.Context(0)
k, d = pyref.chain_scalar(pyref.SEED_POINTS)
for g2, log_n in ((False, 24), (False, 20), (True, 22)):
    n = 1 << log_n
    p = torch.empty((16 if g2 else 8) * n, dtype=torch.int64, device="cuda"); s = torch.empty(4 * n, dtype=torch.int64, device="cuda")
    (ctx.g2_chain_device if g2 else ctx.g1_chain_device)(p, 0, n, k, d); ctx.fr_random_device(s, n, pyref.SEED_SCALARS, 0)
    h = (ctx.g2_bases_from_device if g2 else ctx.g1_bases_from_device)(p, n); del p
    ctx.bases_precompute(h, 0)
    ctx.set_profiling(True)
    run = ctx.g2_msm_resident_device if g2 else ctx.g1_msm_resident_device
    outs = [run(h, s, n) for _ in range(3)]
    ph = ctx.last_msm_phase_ms()
    ctx.set_profiling(False)
    print(json.dumps({"chunk": os.environ.get("B200ZK_CHUNK", "32 (default)"), "g2": g2, "log_n": log_n, "bucket_reduce_ms": ph["bucket_reduce"], "accumulate_ms": ph["accumulate"],
                      "out": outs[-1].hex()[:16]}), flush=True)
    ctx.bases_free(h); del s
