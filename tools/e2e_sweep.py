"""Sweep of the chunk count (b200zk_set_msm_chunks) and the chunk growth ratio (B200ZK_CHUNK_RATIO) of the pipelined
host-scalar MSM: e2e wall time of
b200zk_g{1,2}_msm_resident at 2^24 with pinned host scalars, resident window tables.  JSON lines."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch  # noqa: E402

import ethrex_b200 as eb  # noqa: E402
import pyref  # noqa: E402


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    groups = sys.argv[2] if len(sys.argv) > 2 else "g1,g2"
    n = 1 << log_n
    torch.cuda.set_device(0)
    ctx = eb.Context(0)
    k, d = pyref.chain_scalar(0xB2000002)
    sc = torch.empty(4 * n, dtype=torch.int64, device="cuda")
    ctx.fr_random_device(sc, n, 0xB2000001, 0)
    hs = torch.empty(4 * n, dtype=torch.int64).pin_memory()
    hs.copy_(sc)
    for g2 in (False, True):
        if ("g2" if g2 else "g1") not in groups:
            continue
        pts = torch.empty((16 if g2 else 8) * n, dtype=torch.int64, device="cuda")
        (ctx.g2_chain_device if g2 else ctx.g1_chain_device)(pts, 0, n, k, d)
        h = (ctx.g2_bases_from_device if g2 else ctx.g1_bases_from_device)(pts, n)
        del pts
        torch.cuda.empty_cache()
        ctx.bases_precompute(h, 0)
        dev = (ctx.g2_msm_resident_device if g2 else ctx.g1_msm_resident_device)
        host = (ctx.g2_msm_resident if g2 else ctx.g1_msm_resident)
        ref = dev(h, sc, n)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            dev(h, sc, n)
        resident_ms = (time.perf_counter() - t0) / 3 * 1e3
        grid = [(K, 1.0) for K in (1, 2, 4, 8)] + [(K, r) for r in (2.0, 3.0, 4.0, 6.0) for K in (2, 3, 4, 5)]
        if len(sys.argv) > 3:  # explicit grid: "K:ratio,K:ratio,..."
            grid = [(int(a.split(":")[0]), float(a.split(":")[1])) for a in sys.argv[3].split(",")]
        for K, ratio in grid:
            ctx.set_msm_chunks(K)
            os.environ["B200ZK_CHUNK_RATIO"] = str(ratio)  # read per call by the library
            assert host(h, hs, n) == ref
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps = 4
            for _ in range(reps):
                host(h, hs, n)
            ms = (time.perf_counter() - t0) / reps * 1e3
            print(json.dumps({"probe": "e2e_chunks", "group": "g2" if g2 else "g1", "log_n": log_n, "chunks": K, "ratio": ratio, "e2e_ms": ms, "resident_ms": resident_ms}), flush=True)
        ctx.set_msm_chunks(0)
        os.environ.pop("B200ZK_CHUNK_RATIO", None)
        ctx.bases_free(h)
        torch.cuda.empty_cache()
    ctx.close()


if __name__ == "__main__":
    main()
