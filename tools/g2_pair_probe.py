"""A/B of the G2 accumulation kernels (B200ZK_G2_PAIR = 0 one thread per slice | 2 | 3 | 4 lane pairs at that many CTAs/SM):
one process per setting (the knob is read once), 2^22 and 2^24 points over window tables; JSON lines."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, time
sys.path[:0] = [%r, %r]
import torch
import ethrex_b200 as eb, pyref
ctx = eb.Context(0)
k, d = pyref.chain_scalar(0xB2000002)
for log_n in (22, 24):
    n = 1 << log_n
    sc = torch.empty(4 * n, dtype=torch.int64, device="cuda"); ctx.fr_random_device(sc, n, 0xB2000001, 0)
    pts = torch.empty(16 * n, dtype=torch.int64, device="cuda"); ctx.g2_chain_device(pts, 0, n, k, d)
    h = ctx.g2_bases_from_device(pts, n); del pts; torch.cuda.empty_cache(); ctx.bases_precompute(h, 0)
    part = torch.zeros(32, dtype=torch.int64, device="cuda")
    out = ctx.g2_msm_resident_device(h, sc, n)
    ctx.set_profiling(True)
    acc = []
    for _ in range(3):
        ctx.g2_msm_partial_resident_device(h, sc, n, part); acc.append(ctx.last_msm_phase_ms())
    ctx.set_profiling(False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): ctx.g2_msm_resident_device(h, sc, n)
    ms = (time.perf_counter() - t0) / 3 * 1e3
    print(json.dumps({"probe": "g2_pair", "knob": os.environ.get("B200ZK_G2_PAIR", "default"), "log_n": log_n, "msm_ms": ms, "phases": acc[-1], "result": out.hex()[:32]}), flush=True)
    ctx.bases_free(h); del sc; torch.cuda.empty_cache()
''' % (ROOT, os.path.join(ROOT, "oracle"))
for knob in ("0", "3", "4"):
    env = dict(os.environ, B200ZK_G2_PAIR=knob)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout)
    if r.returncode:
        sys.stdout.write(json.dumps({"probe": "g2_pair", "knob": knob, "error": r.stderr[-400:]}) + "\n")
    sys.stdout.flush()
