"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in ethrex_b200/dist.py.

The GPU calls are replaced by a stand-in context backed by the CPU oracle -- this file tests the sharding /
all-gather / fold plumbing, not the kernels (tests/test_gpu_parity.py::test_partials_fold_equals_whole and
bench.py --gpus N cover those on real devices)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cpu_oracle as orc
import pyref
from ethrex_b200.dist import msm_sharded, ntt_batch_assignment, shard_range


class OracleCtx:
    """Quacks like ethrex_b200.Context for the three calls msm_sharded makes; partial = affine point as XYZZ with ZZ=ZZZ=1."""

    def __init__(self, pts, scalars):
        self.pts, self.scalars = pts, scalars

    def g1_msm_partial_device(self, d_points, d_scalars, n, d_partial, flags=0):
        be = orc.g1_msm(d_points.numpy().view(np.uint64).reshape(-1, 8), d_scalars.numpy().view(np.uint64).reshape(-1, 4)[:n])
        nat = orc.g1_be_to_native(be).reshape(-1)
        one = orc.fq_to_mont(orc.ints_to_array([1])).reshape(-1)
        zero = np.zeros(4, dtype=np.uint64)
        inf = be == bytes(64)
        xyzz = np.concatenate([nat, zero if inf else one, zero if inf else one])
        d_partial.copy_(torch.from_numpy(xyzz.view(np.int64).copy()))

    def g1_fold_partials_device(self, gathered, count, flags=0):
        acc = bytes(64)
        g = gathered.numpy().view(np.uint64).reshape(count, 16)
        for k in range(count):
            pt = bytes(64) if not g[k, 8:].any() else orc.g1_native_to_be(g[k, :8])
            acc = orc.g1_add_be(acc, pt)[1]
        return acc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        k, d = pyref.chain_scalar(pyref.SEED_POINTS)
        lo, hi = shard_range(n_total, rank, world)
        pts = orc.g1_chain(n_total, k, d)[lo:hi]
        s = orc.rand_fr(pyref.SEED_SCALARS, lo, hi - lo)
        out = msm_sharded(OracleCtx(pts, s), torch.from_numpy(pts.view(np.int64).copy()), torch.from_numpy(s.view(np.int64).copy()), hi - lo)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [1001, 64])
def test_msm_sharded_world2_gloo(n_total):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    k, d = pyref.chain_scalar(pyref.SEED_POINTS)
    expected = orc.g1_msm(orc.g1_chain(n_total, k, d), orc.rand_fr(pyref.SEED_SCALARS, 0, n_total))
    assert res[0] == res[1] == expected


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 1 << 24):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
    assert sorted(sum((ntt_batch_assignment(7, r, 3) for r in range(3)), [])) == list(range(7))


# ---- the multi-GPU proof path (round 2): dealt NTTs + ONE all_gather of the 768-byte blocks + fold -------------------
class OracleProofCtx:
    """Quacks like ethrex_b200.Context for dist.quotient_dealt (fr_ntt_device / fr_quotient_device on CPU tensors) and for the
    commit_partial / fold pair: a block holds A | B1 | B2 | L | H as XYZZ with ZZ = ZZZ = 1 (or all zero for the identity)."""
    NTT_INVERSE, NTT_COSET = 1 << 4, 1 << 5  # ethrex_b200._ffi flag values

    def fr_ntt_device(self, t, log_n, flags):
        oflags = (orc.NTT_INVERSE if flags & self.NTT_INVERSE else 0) | (orc.NTT_COSET if flags & self.NTT_COSET else 0)
        a = orc.fr_ntt(t.numpy().view(np.uint64).reshape(-1, 4), log_n, oflags)
        t.copy_(torch.from_numpy(a.view(np.int64).reshape(-1).copy()))

    def fr_quotient_device(self, a, b, c, out, n, zinv):
        r = orc.fr_quotient(a.numpy().view(np.uint64).reshape(-1, 4), b.numpy().view(np.uint64).reshape(-1, 4), c.numpy().view(np.uint64).reshape(-1, 4), zinv)
        out.copy_(torch.from_numpy(r.view(np.int64).reshape(-1).copy()))


def _g1_xyzz(be: bytes) -> np.ndarray:
    one = orc.fq_to_mont(orc.ints_to_array([1])).reshape(-1)
    if be == bytes(64):
        return np.zeros(16, dtype=np.uint64)
    return np.concatenate([orc.g1_be_to_native(be).reshape(-1), one, one])


def _proof_worker(rank, world, port, log_n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ethrex_b200.dist import quotient_dealt
        from ethrex_b200.groth16 import coset_vanishing_inverse
        n = 1 << log_n
        a = orc.fr_to_mont(orc.rand_fr(11, 0, n)); b = orc.fr_to_mont(orc.rand_fr(12, 0, n)); c = orc.field_mul("fr", a, b)
        ta, tb, tc = (torch.from_numpy(x.view(np.int64).reshape(-1).copy()) for x in (a, b, c))
        h = quotient_dealt(OracleProofCtx(), log_n, ta, tb, tc, coset_vanishing_inverse(log_n), rank, world)
        h_can = orc.fr_from_mont(h.numpy().view(np.uint64).reshape(n, 4))
        # this rank's slice of one proving-key column against its slice of H, as a block; ONE all_gather; fold on every rank
        k, d = pyref.chain_scalar(pyref.SEED_POINTS)
        lo, hi = shard_range(n, rank, world)
        hi_h = min(hi, n - 1)
        part = orc.g1_msm(orc.g1_chain(n, k, d)[lo:hi_h], h_can[lo:hi_h]) if hi_h > lo else bytes(64)
        block = torch.zeros(96, dtype=torch.int64)
        block[80:96] = torch.from_numpy(_g1_xyzz(part).view(np.int64).copy())  # the H slot of A | B1 | B2 | L | H (bytes 640..767)
        gathered = torch.empty(96 * world, dtype=torch.int64)
        dist.all_gather_into_tensor(gathered, block)
        acc = bytes(64)
        g = gathered.numpy().view(np.uint64).reshape(world, 96)
        for r in range(world):
            x = g[r, 80:96]
            acc = orc.g1_add_be(acc, bytes(64) if not x[8:].any() else orc.g1_native_to_be(x[:8]))[1]
        q.put((rank, h_can.tobytes(), acc))
    finally:
        dist.destroy_process_group()


def test_dealt_quotient_and_block_fold_world2_gloo():
    """dist.quotient_dealt on 2 ranks = the single-process quotient; the per-rank H commitments, all-gathered as 768-byte
    blocks and folded, = the whole-column MSM (the host plumbing of SyntheticWrapCircuit.prove_device for world > 1)."""
    from ethrex_b200.groth16 import coset_vanishing_inverse
    log_n, world, port = 8, 2, _free_port()
    n = 1 << log_n
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_proof_worker, args=(r, world, port, log_n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {r: (h, acc) for r, h, acc in (q.get(timeout=180) for _ in range(world))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a = orc.fr_to_mont(orc.rand_fr(11, 0, n)); b = orc.fr_to_mont(orc.rand_fr(12, 0, n)); c = orc.field_mul("fr", a, b)
    cos = [orc.fr_ntt(orc.fr_ntt(p, log_n, orc.NTT_INVERSE), log_n, orc.NTT_COSET) for p in (a, b, c)]
    h = orc.fr_from_mont(orc.fr_ntt(orc.fr_quotient(cos[0], cos[1], cos[2], coset_vanishing_inverse(log_n)), log_n, orc.NTT_INVERSE | orc.NTT_COSET))
    assert res[0][0] == res[1][0] == h.tobytes()
    assert orc.array_to_ints(h)[n - 1] == 0  # the quotient is exact
    k, d = pyref.chain_scalar(pyref.SEED_POINTS)
    whole = orc.g1_msm(orc.g1_chain(n, k, d)[:n - 1], h[:n - 1])
    assert res[0][1] == res[1][1] == whole
