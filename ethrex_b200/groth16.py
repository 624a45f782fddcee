"""Groth16-shaped commitment pipeline on the B200 kernels (BASELINE config #5).

This is the arithmetic the SNARK wrap behind `ProofFormat::Groth16` performs
(/root/reference/crates/prover/src/backend/sp1.rs:97-134 -> gnark; risc0.rs:24-29,71-82 -> risc0-groth16):

    quotient   3 iNTT (A,B,C evaluations -> coefficients), 3 coset NTT, pointwise (a*b - c)/Z_H, 1 coset iNTT
    commit     [A]1 = MSM(pk.A_g1, w)   [B]1 = MSM(pk.B_g1, w)   [B]2 = MSM(pk.B_g2, w)
               [L]1 = MSM(pk.L_g1, w_private)   [H]1 = MSM(pk.H_g1, h)
    assemble   proof = A | B | C with C = [L]1 + [H]1      (EIP-197 byte order, 256 bytes)

The reference's real wrap circuit and proving key live inside the zkVM SDKs and are not in the tree, so the
circuit here is SYNTHETIC: witness and constraint evaluations are derived
deterministically from the serialized program input, with C = A o B on the domain so that the quotient is an
exact polynomial, and the proving key is a set of chain bases.  The STARK stage is excluded and no blinding
is applied; what is measured and parity-checked is exactly the MSM + NTT work of a Groth16 prove.
All arithmetic runs in libb200zk.so; Python only sequences the calls.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field

from . import _ffi as F

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
COSET_GEN = 5  # ark/gnark multiplicative generator of Fr


def _seed64(data: bytes, tag: bytes) -> int:
    return int.from_bytes(hashlib.sha256(tag + data).digest()[:8], "little")


def _chain_kd(tag: bytes) -> tuple[int, int]:
    h = hashlib.sha256(b"b200zk-pk-" + tag).digest()
    return (int.from_bytes(h[:16], "little") | 1), (int.from_bytes(h[16:], "little") | 1)


@dataclass
class ProvingKey:
    """Resident (precomputed) proving-key columns: handles into the context + the chain scalars that define them."""
    log_n: int
    handles: dict = field(default_factory=dict)   # name -> b200zk bases handle
    chains: dict = field(default_factory=dict)    # name -> (k, d, is_g2)


class SyntheticWrapCircuit:
    """Domain size 2^log_n; `n` witness entries; H has n-1 coefficients."""
    QUERIES = (("a_g1", False), ("b_g1", False), ("b_g2", True), ("l_g1", False), ("h_g1", False))

    def __init__(self, ctx, log_n: int, precompute: bool = True, g2: bool = True, rank: int = 0, world: int = 1):
        """rank/world: every rank keeps only its contiguous shard [lo, hi) of each proving-key column (point-split
        MSMs, ethrex_b200.dist.msm_sharded); the NTTs of the quotient are cheap and computed on every rank."""
        import torch
        from .dist import shard_range
        self.ctx, self.log_n, self.n = ctx, log_n, 1 << log_n
        self.rank, self.world = rank, world
        self.lo, self.hi = shard_range(self.n, rank, world)
        self.pk = ProvingKey(log_n)
        m = self.hi - self.lo
        for name, is_g2 in self.QUERIES:
            if is_g2 and not g2:
                continue
            k, d = _chain_kd(name.encode())
            pts = torch.empty((16 if is_g2 else 8) * m, dtype=torch.int64, device="cuda")
            (ctx.g2_chain_device if is_g2 else ctx.g1_chain_device)(pts, self.lo, m, k, d)
            h = (ctx.g2_bases_from_device if is_g2 else ctx.g1_bases_from_device)(pts, m)
            del pts
            if precompute:
                ctx.bases_precompute(h, 0)
            self.pk.handles[name] = h
            self.pk.chains[name] = (k, d, is_g2)
        # Z_H on the coset h*<w> is the constant h^n - 1
        self.zinv = pow((pow(COSET_GEN, self.n, R_MOD) - 1) % R_MOD, -1, R_MOD)

    def close(self):
        for h in self.pk.handles.values():
            self.ctx.bases_free(h)
        self.pk.handles.clear()

    # ---- witness / constraint evaluations from the serialized input (deterministic) ----
    def assign(self, serialized_input: bytes):
        """Returns device tensors (witness canonical, a/b/c evaluations Montgomery)."""
        import torch
        n, ctx = self.n, self.ctx
        w = torch.empty(4 * n, dtype=torch.int64, device="cuda")
        a = torch.empty(4 * n, dtype=torch.int64, device="cuda")
        b = torch.empty(4 * n, dtype=torch.int64, device="cuda")
        c = torch.empty(4 * n, dtype=torch.int64, device="cuda")
        ctx.fr_random_device(w, n, _seed64(serialized_input, b"witness"), 0)
        ctx.fr_random_device(a, n, _seed64(serialized_input, b"A"), 0, F.SCALARS_MONT)
        ctx.fr_random_device(b, n, _seed64(serialized_input, b"B"), 0, F.SCALARS_MONT)
        ctx.field_mul_device(a, b, c, n, 1)  # C = A o B on the domain: the R1CS is satisfied
        return w, a, b, c

    # ---- the hot path ----
    def quotient(self, a, b, c):
        """H coefficients (Montgomery) from the A,B,C evaluations; a is overwritten with H."""
        ctx, k = self.ctx, self.log_n
        for poly in (a, b, c):
            ctx.fr_ntt_device(poly, k, F.NTT_INVERSE)   # evaluations -> coefficients
            ctx.fr_ntt_device(poly, k, F.NTT_COSET)     # coefficients -> evaluations on the coset
        ctx.fr_quotient_device(a, b, c, a, self.n, self.zinv)
        ctx.fr_ntt_device(a, k, F.NTT_INVERSE | F.NTT_COSET)
        return a

    def commit(self, w, h_coeffs, msm=None):
        """The five MSMs.  `msm(name, scalars, n, flags)` lets the multi-GPU driver substitute a sharded MSM."""
        ctx, n = self.ctx, self.n

        def local(name, scalars, count, flags):
            hnd = self.pk.handles[name]
            is_g2 = self.pk.chains[name][2]
            if self.world > 1:  # this rank's slice of the scalars against its shard of the column, then all-gather + fold
                from .dist import msm_sharded
                hi = min(self.hi, count)
                m = max(0, hi - self.lo)
                return msm_sharded(ctx, None, scalars[4 * self.lo: 4 * (self.lo + max(m, 1))], m, flags, g2=is_g2, handle=hnd)
            if is_g2:
                return ctx.g2_msm_resident_device(hnd, scalars, count, flags)
            return ctx.g1_msm_resident_device(hnd, scalars, count, flags)

        run = msm or local
        out = {"a_g1": run("a_g1", w, n, 0), "b_g1": run("b_g1", w, n, 0)}
        if "b_g2" in self.pk.handles:
            out["b_g2"] = run("b_g2", w, n, 0)
        out["l_g1"] = run("l_g1", w, n, 0)
        out["h_g1"] = run("h_g1", h_coeffs, n - 1, F.SCALARS_MONT)
        return out

    def assemble(self, commitments) -> bytes:
        """proof = A (64) | B (128, x_im|x_re|y_im|y_re) | C (64), C = L + H folded on the device."""
        one = (1).to_bytes(32, "little")
        # C = 1*L + 1*H: a 2-point MSM through the host entry point (validates the two encodings on the way in)
        c_pt = self.ctx.g1_msm(commitments["l_g1"] + commitments["h_g1"], one + one, 2, F.POINTS_BE)
        b2 = commitments.get("b_g2", bytes(128))
        return commitments["a_g1"] + b2 + c_pt

    def prove(self, serialized_input: bytes, msm=None) -> bytes:
        w, a, b, c = self.assign(serialized_input)
        h = self.quotient(a, b, c)
        return self.assemble(self.commit(w, h, msm))
