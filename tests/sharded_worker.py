"""torchrun worker of tests/test_gpu_r2.py::test_sharded_g1_g2_msm_two_ranks: BASELINE config 4 in miniature -- one
2^k-point G1 MSM and one G2 MSM point-split over WORLD_SIZE ranks, NCCL all_gather of the XYZZ partials, local fold;
every rank checks the bytes against the closed form of the chain MSM (and rank 0 against the one-GPU result)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import cpu_oracle as orc  # noqa: E402
import pyref  # noqa: E402
import ethrex_b200 as eb  # noqa: E402
from ethrex_b200.dist import msm_sharded, shard_range  # noqa: E402
from helpers import expected_chain_msm_g1, expected_chain_msm_g2  # noqa: E402


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = eb.Context(local)
    n = 1 << log_n
    lo, hi = shard_range(n, rank, world)
    m = hi - lo
    k, d = pyref.chain_scalar(pyref.SEED_POINTS)
    full = torch.empty(4 * n, dtype=torch.int64, device="cuda")
    ctx.fr_random_device(full, n, pyref.SEED_SCALARS, 0)
    s_host = full.cpu().numpy().view(np.uint64).reshape(n, 4)
    for g2 in (False, True):
        w = 16 if g2 else 8
        pts = torch.empty(w * m, dtype=torch.int64, device="cuda")
        (ctx.g2_chain_device if g2 else ctx.g1_chain_device)(pts, lo, m, k, d)
        exp = (expected_chain_msm_g2 if g2 else expected_chain_msm_g1)(s_host, k, d)
        got = msm_sharded(ctx, pts, full[4 * lo:4 * hi], m, g2=g2)
        assert got == exp, f"rank {rank} g2={g2}: sharded MSM differs from the closed form"
        # resident window tables on every rank (the bench's configuration)
        h = (ctx.g2_bases_from_device if g2 else ctx.g1_bases_from_device)(pts, m)
        ctx.bases_precompute(h, 0)
        got = msm_sharded(ctx, None, full[4 * lo:4 * hi], m, g2=g2, handle=h)
        assert got == exp, f"rank {rank} g2={g2}: sharded table MSM differs from the closed form"
        ctx.bases_free(h)
        if rank == 0:  # the one-GPU answer over the whole input
            allp = torch.empty(w * n, dtype=torch.int64, device="cuda")
            (ctx.g2_chain_device if g2 else ctx.g1_chain_device)(allp, 0, n, k, d)
            assert (ctx.g2_msm_device if g2 else ctx.g1_msm_device)(allp, full, n) == exp
            del allp
    dist.barrier()
    print("SHARDED_OK", rank, flush=True)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
