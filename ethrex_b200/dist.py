"""Multi-GPU host logic: one process per GPU, point-split MSM.

Every rank reduces its own shard of bases/scalars to ONE extended-Jacobian partial sum on its GPU
(`b200zk_g{1,2}_msm_partial_device`), the 128/256-byte partials are all-gathered (NCCL over NVLink on the
GPU box, gloo in the CPU tests) and every rank folds them (`b200zk_g{1,2}_fold_partials_device`), so all
ranks return the same bytes.  NCCL has no elliptic-curve reduction operator: all-gather + local fold IS the
"allreduce of partial sums"; it moves world*128 bytes once per MSM and is latency bound.
"""
from __future__ import annotations


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) slice of rank `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def msm_sharded(ctx, d_points, d_scalars, n_local: int, flags: int = 0, g2: bool = False, group=None, handle: int | None = None) -> bytes:
    """MSM over the union of all ranks' shards.  `d_points`/`d_scalars`: this rank's shard (device tensors);
    with `handle` the shard's bases are the resident (possibly precomputed) ones and `d_points` is ignored."""
    import torch
    import torch.distributed as dist
    words = 32 if g2 else 16  # XYZZ partial in int64 words
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    host_scalars = handle is not None and not getattr(d_scalars, "is_cuda", False)
    partial = torch.zeros(words, dtype=torch.int64, device="cuda" if host_scalars else d_scalars.device)
    if host_scalars:  # resident bases + pinned host scalars: the upload is pipelined inside the library
        (ctx.g2_msm_partial_resident if g2 else ctx.g1_msm_partial_resident)(handle, d_scalars, n_local, partial, flags)
    elif handle is not None:
        (ctx.g2_msm_partial_resident_device if g2 else ctx.g1_msm_partial_resident_device)(handle, d_scalars, n_local, partial, flags)
    else:
        (ctx.g2_msm_partial_device if g2 else ctx.g1_msm_partial_device)(d_points, d_scalars, n_local, partial, flags)
    if world == 1:
        gathered = partial
    else:
        gathered = torch.empty(words * world, dtype=torch.int64, device=partial.device)
        dist.all_gather_into_tensor(gathered, partial, group=group)
    return (ctx.g2_fold_partials_device if g2 else ctx.g1_fold_partials_device)(gathered, world, flags)


def ntt_batch_assignment(num_polys: int, rank: int, world: int) -> list[int]:
    """Independent transforms (the 3+3+1 of a Groth16 quotient) are dealt round-robin: replicas, no collective."""
    return [i for i in range(num_polys) if i % world == rank]


def quotient_dealt(ctx, log_n: int, a, b, c, zinv: int, rank: int, world: int, group=None):
    """Groth16 quotient H = (A*B - C)/Z_H on `world` ranks: the three (iNTT, coset NTT) pairs are independent
    transforms and are dealt round-robin (`ntt_batch_assignment`: replicas), each owner
    broadcasts its coset evaluations (n x 32 bytes over NCCL / NVLink), then every rank does the cheap tail itself
    (pointwise quotient + ONE coset iNTT) -- 2 + 1 transforms per rank instead of 7.  a, b, c: Montgomery device
    tensors holding the evaluations on the domain; returns H's coefficients (in `a`) on every rank."""
    import torch.distributed as dist
    from . import _ffi as F
    polys = (a, b, c)
    for i in ntt_batch_assignment(3, rank, world):
        ctx.fr_ntt_device(polys[i], log_n, F.NTT_INVERSE)
        ctx.fr_ntt_device(polys[i], log_n, F.NTT_COSET)
    if world > 1:
        works = [dist.broadcast(polys[i], src=(i % world if group is None else dist.get_global_rank(group, i % world)), group=group, async_op=True) for i in range(3)]
        for wk in works:
            wk.wait()
    ctx.fr_quotient_device(a, b, c, a, 1 << log_n, zinv)
    ctx.fr_ntt_device(a, log_n, F.NTT_INVERSE | F.NTT_COSET)
    return a
