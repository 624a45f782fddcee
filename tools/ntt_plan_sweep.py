"""Sweep of the NTT pass plan at 2^24 (B200ZK_NTT_PLAN is read on every call): forward transform, median of 7.  JSON lines."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch  # noqa: E402

import ethrex_b200 as eb  # noqa: E402

ctx = eb.Context(0)
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << log_n
d = torch.empty(4 * n, dtype=torch.int64, device="cuda")
ctx.fr_random_device(d, n, 3, 0, eb.SCALARS_MONT)
ref = None


def timed(fn, iters=7, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


plans = ["", "8,8,8", "7,8,9", "6,9,9", "9,8,7", "9,9,6", "7,9,8", "8,9,7", "10,7,7", "7,7,10", "6,8,10", "10,8,6", "5,9,10", "11,7,6", "6,7,11", "8,7,9", "9,7,8"] if log_n == 24 else [""]
for tile in ("11", "10", "12"):
    os.environ["B200ZK_NTT_TILE_LOG"] = tile
    for plan in plans:
        if plan:
            os.environ["B200ZK_NTT_PLAN"] = plan
            if max(int(x) for x in plan.split(",")) > int(tile) and tile != "11":
                continue
        else:
            os.environ.pop("B200ZK_NTT_PLAN", None)
        src = d.clone()
        try:
            ctx.fr_ntt_device(src, log_n, 0)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"probe": "ntt_plan", "plan": plan, "tile_log": tile, "error": str(e)[:100]}), flush=True)
            continue
        if ref is None:
            ref = src.clone()
        ok = bool(torch.equal(src, ref))
        ms = timed(lambda: ctx.fr_ntt_device(src, log_n, 0))
        print(json.dumps({"probe": "ntt_plan", "log_n": log_n, "plan": plan or "default", "tile_log": tile, "ms": round(ms, 4), "same_result": ok}), flush=True)
