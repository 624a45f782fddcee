"""EIP-4844 blob commitments on the B200 kernels: the host-side mirror of the reference's `crypto::kzg` surface
(/root/reference/crates/common/crypto/kzg.rs:259-293, used by /root/reference/crates/common/types/blobs_bundle.rs:90-118
and the L2 committer, crates/l2/sequencer/l1_committer.rs:1488-1521).

    blob_to_kzg_commitment(blob)               one 4096-point BLS12-381 G1 MSM over the Lagrange-form trusted setup
    compute_kzg_proof(blob, z)                 p(z) by the barycentric formula + the quotient's commitment (a second MSM)
    compute_blob_kzg_proof(blob, commitment)   the same at the Fiat-Shamir challenge of EIP-4844
    blob_to_kzg_commitment_and_proof(blob)     what `BlobsBundle::create_from_blobs` calls per blob (wrapper version 0)

The two MSMs run on the GPU (libb200zk.so, `b200zk_kzg_blob_to_commitment` / `b200zk_bls12_381_g1_msm_resident`); the scalar-
field bookkeeping around them (4096 modular inverses, one SHA-256) is host work, exactly as it is CPU work in c-kzg.  The
trusted setup is an INPUT (4096 compressed G1 points in c-kzg's g1_lagrange_brp order): the reference gets it from inside the
c-kzg / kzg-rs crates, which are not in the tree, so no setup is bundled here.  Cell proofs (wrapper version 1) are out of scope.
"""
from __future__ import annotations

import hashlib

BLS_MODULUS = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
FIELD_ELEMENTS_PER_BLOB = 4096
BYTES_PER_BLOB = 32 * FIELD_ELEMENTS_PER_BLOB
FIAT_SHAMIR_PROTOCOL_DOMAIN = b"FSBLOBVERIFY_V1_"
PRIMITIVE_ROOT_OF_UNITY = 7


def _bit_reverse(i: int, bits: int) -> int:
    return int(format(i, f"0{bits}b")[::-1], 2)


_ROOTS_BRP = None


def roots_of_unity_brp():
    """the 4096 roots of unity in bit-reversed order (the order blobs list their evaluations in)"""
    global _ROOTS_BRP
    if _ROOTS_BRP is None:
        w = pow(PRIMITIVE_ROOT_OF_UNITY, (BLS_MODULUS - 1) // FIELD_ELEMENTS_PER_BLOB, BLS_MODULUS)
        nat = [1] * FIELD_ELEMENTS_PER_BLOB
        for i in range(1, FIELD_ELEMENTS_PER_BLOB):
            nat[i] = nat[i - 1] * w % BLS_MODULUS
        _ROOTS_BRP = [nat[_bit_reverse(i, 12)] for i in range(FIELD_ELEMENTS_PER_BLOB)]
    return _ROOTS_BRP


class KzgSettings:
    """The trusted setup resident in HBM (the reference's `c_kzg::ethereum_kzg_settings(KZG_PRECOMPUTE)`, kzg.rs:262)."""

    def __init__(self, ctx, g1_lagrange_brp: bytes, precompute: bool = True):
        if len(g1_lagrange_brp) != 48 * FIELD_ELEMENTS_PER_BLOB:
            raise ValueError("the setup is 4096 compressed G1 points (48 bytes each) in g1_lagrange_brp order")
        self.ctx = ctx
        self.handle = ctx.bls12_381_g1_bases_upload(g1_lagrange_brp, FIELD_ELEMENTS_PER_BLOB)
        if precompute:
            ctx.bases_precompute(self.handle, 0)

    def close(self):
        if self.handle:
            self.ctx.bases_free(self.handle)
            self.handle = 0

    # ---- kzg.rs:259-272
    def blob_to_kzg_commitment(self, blob: bytes) -> bytes:
        if len(blob) != BYTES_PER_BLOB:
            raise ValueError("a blob is 131072 bytes")
        return self.ctx.kzg_blob_to_commitment(self.handle, blob)[0]

    def blobs_to_kzg_commitments(self, blobs) -> list:
        return self.ctx.kzg_blob_to_commitment(self.handle, b"".join(blobs)) if blobs else []

    def compute_kzg_proof(self, blob: bytes, z: int):
        """-> (proof 48 bytes, y = p(z)): the quotient (p(x) - y) / (x - z) in evaluation form, committed with one MSM"""
        poly = [int.from_bytes(blob[32 * i:32 * i + 32], "big") for i in range(FIELD_ELEMENTS_PER_BLOB)]
        if any(v >= BLS_MODULUS for v in poly) or not 0 <= z < BLS_MODULUS:
            raise ValueError("field element out of range")
        roots = roots_of_unity_brp()
        r = BLS_MODULUS
        if z in roots:
            m = roots.index(z)
            y = poly[m]
            # q_m = sum_{i != m} (p_i - y) w_i / (z (z - w_i));  q_i = (p_i - y) / (w_i - z) elsewhere
            q = [0] * FIELD_ELEMENTS_PER_BLOB
            zinv = pow(z, -1, r)
            for i, w in enumerate(roots):
                if i == m:
                    continue
                d = pow((w - z) % r, -1, r)
                q[i] = (poly[i] - y) * d % r
                q[m] = (q[m] + (poly[i] - y) * w % r * zinv % r * pow((z - w) % r, -1, r)) % r
        else:
            # barycentric evaluation: p(z) = (z^n - 1)/n * sum_i p_i w_i / (z - w_i)
            inv = [pow((z - w) % r, -1, r) for w in roots]
            acc = sum(p * w % r * d for p, w, d in zip(poly, roots, inv)) % r
            y = (pow(z, FIELD_ELEMENTS_PER_BLOB, r) - 1) * pow(FIELD_ELEMENTS_PER_BLOB, -1, r) % r * acc % r
            q = [(y - p) * d % r for p, d in zip(poly, inv)]  # (p_i - y)/(w_i - z) = (y - p_i)/(z - w_i)
        scalars = b"".join(v.to_bytes(32, "big") for v in q)
        return self.ctx.bls12_381_g1_msm_resident(self.handle, scalars, FIELD_ELEMENTS_PER_BLOB), y

    @staticmethod
    def compute_challenge(blob: bytes, commitment: bytes) -> int:
        """EIP-4844 compute_challenge: hash_to_bls_field(domain | degree (16 B BE) | blob | commitment)"""
        data = FIAT_SHAMIR_PROTOCOL_DOMAIN + FIELD_ELEMENTS_PER_BLOB.to_bytes(16, "big") + blob + commitment
        return int.from_bytes(hashlib.sha256(data).digest(), "big") % BLS_MODULUS

    def compute_blob_kzg_proof(self, blob: bytes, commitment: bytes) -> bytes:
        return self.compute_kzg_proof(blob, self.compute_challenge(blob, commitment))[0]

    def blob_to_kzg_commitment_and_proof(self, blob: bytes):
        c = self.blob_to_kzg_commitment(blob)
        return c, self.compute_blob_kzg_proof(blob, c)
