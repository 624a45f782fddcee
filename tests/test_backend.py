"""B200Backend: the ProverBackend mirror (CPU-side behaviour) and, on a GPU, the Groth16-shaped pipeline
against the same pipeline computed with the CPU oracle."""
import numpy as np
import pytest

import cpu_oracle as orc
import pyref
import ethrex_b200 as eb
from ethrex_b200.backend import B200Backend, BackendType, ProofFormat, ProverType, serialize_program_input
from ethrex_b200.groth16 import COSET_GEN, SyntheticWrapCircuit, _chain_kd, _seed64


def test_backend_without_circuit_mirrors_reference_errors():
    """error behaviour of a ProverBackend without its SDK (backend/mod.rs:81-147, error.rs:3-51)."""
    b = B200Backend()
    assert b.prover_type() is ProverType.EXEC
    assert BackendType.from_str("B200") is BackendType.B200 and BackendType.from_str("exec") is BackendType.EXEC
    with pytest.raises(ValueError, match="Invalid backend"):
        BackendType.from_str("sp2")
    with pytest.raises(eb.B200Error) as e:
        b.prove({"blocks": []})
    assert e.value.kind == "NotImplemented"
    with pytest.raises(eb.B200Error) as e:
        b.prove(b"x", ProofFormat.COMPRESSED)
    assert e.value.kind == "NotImplemented"
    with pytest.raises(eb.B200Error, match="Verify not implemented for this backend"):
        b.verify(None)
    with pytest.raises(eb.B200Error) as e:
        b.serialize_input({"x": object()})
    assert e.value.kind == "Serialization"
    assert serialize_program_input({"b": 1, "a": [2]}) == b'{"a":[2],"b":1}'
    s, dt = b.serialize_input_timed(b"abc")
    assert s == b"abc" and dt >= 0
    assert b.execute_timed(b"abc") >= 0


@pytest.mark.gpu
def test_groth16_pipeline_matches_oracle(ctx):
    """config #5 at 2^10: quotient (7 NTTs + pointwise) and the five MSMs, byte for byte."""
    log_n, n = 10, 1 << 10
    circuit = SyntheticWrapCircuit(ctx, log_n, precompute=True)
    backend = B200Backend(ctx, circuit)
    try:
        inp = {"blocks": [1, 2, 3], "elasticity_multiplier": 2}
        proof, dt = backend.prove_timed(inp, ProofFormat.GROTH16)
        out = backend.to_proof_bytes(proof, ProofFormat.GROTH16)
        assert out.prover_type() is ProverType.SP1 and len(out.proof_bytes.proof) == 256
        # ---- the same pipeline on the CPU oracle
        ser = backend.serialize_input(inp)
        w = orc.rand_fr(_seed64(ser, b"witness"), 0, n)
        a = orc.fr_to_mont(orc.rand_fr(_seed64(ser, b"A"), 0, n))
        b = orc.fr_to_mont(orc.rand_fr(_seed64(ser, b"B"), 0, n))
        c = orc.field_mul("fr", a, b)
        cos = []
        for poly in (a, b, c):
            coeff = orc.fr_ntt(poly, log_n, orc.NTT_INVERSE)
            cos.append(orc.array_to_ints(orc.fr_from_mont(orc.fr_ntt(coeff, log_n, orc.NTT_COSET))))
        zinv = pow((pow(COSET_GEN, n, pyref.R) - 1) % pyref.R, -1, pyref.R)
        hq = [((x * y - z) * zinv) % pyref.R for x, y, z in zip(*cos)]
        h = orc.fr_ntt(orc.fr_to_mont(orc.ints_to_array(hq)), log_n, orc.NTT_INVERSE | orc.NTT_COSET)
        h_can = orc.fr_from_mont(h)
        assert orc.array_to_ints(h_can)[n - 1] == 0  # deg H < n-1: the quotient is exact
        exp = {}
        for name, is_g2 in SyntheticWrapCircuit.QUERIES:
            k, d = _chain_kd(name.encode())
            sc, cnt = (h_can, n - 1) if name == "h_g1" else (w, n)
            if is_g2:
                exp[name] = orc.g2_msm(orc.g2_chain(n, k, d)[:cnt], sc[:cnt])
            else:
                exp[name] = orc.g1_msm(orc.g1_chain(n, k, d)[:cnt], sc[:cnt])
        c_pt = orc.g1_add_be(exp["l_g1"], exp["h_g1"])[1]
        # the one-call path (b200zk_groth16_commit) reports A, B1, B2 and C = L + H
        assert proof.commitments == {"a_g1": exp["a_g1"], "b_g1": exp["b_g1"], "b_g2": exp["b_g2"], "c_g1": c_pt}
        assert proof.proof == exp["a_g1"] + exp["b_g2"] + c_pt
        # the separate-call path (five read-back MSMs) reports every commitment and lands on the same proof
        proof2, cm = circuit.prove_separate(ser)
        for name in exp:
            assert cm[name] == exp[name], name
        assert proof2 == proof.proof
        # deterministic: same input, same proof; different input, different proof
        assert backend.prove(inp).proof == proof.proof
        assert backend.prove({"blocks": [9]}).proof != proof.proof
    finally:
        circuit.close()


def test_groth16_verifier_host_logic_without_a_gpu():
    """Groth16Verifier's host side (negating A, the public-input combination, the pair order) against the toy
    instance's own verifier calldata, with the MSM call served by the oracle; and the Verification error path of
    B200Backend.verify with a stub verifier.  The device half runs in test_gpu_parity.py."""
    from ethrex_b200.backend import B200ProveOutput
    from ethrex_b200.groth16 import Groth16Verifier
    from groth16_toy import ToyGroth16, _g1

    class OracleMsm:
        def g1_msm(self, pts, sc, n, flags):
            acc = None
            for i in range(n):
                acc = pyref.g1_add(acc, pyref.g1_mul(int.from_bytes(sc[32 * i:32 * i + 32], "big"), pyref.g1_from_be(pts[64 * i:64 * i + 64])))
            return pyref.g1_to_be(acc)

    toy = ToyGroth16(3)
    ver = Groth16Verifier(OracleMsm(), toy.vk_alpha_g1, toy.vk_beta_g2, toy.vk_gamma_g2, toy.vk_delta_g2, [_g1(s) for s in toy.ic])
    for x in (0, 12345, pyref.R - 1):
        proof = toy.expected_proof(toy.assign(x))
        assert ver.calldata(proof, [x]) == toy.verifier_calldata(proof, x)
    with pytest.raises(ValueError):
        ver.calldata(proof[:255], [1])
    # non-canonical encodings are rejected, never reduced: A = (x, y + p), public input x + r, negative input
    x = 12345
    proof = toy.expected_proof(toy.assign(x))
    ay = int.from_bytes(proof[32:64], "big")
    assert ver.calldata(proof[:32] + (ay + pyref.P).to_bytes(32, "big") + proof[64:], [x]) is None
    assert ver.calldata((pyref.P).to_bytes(32, "big") + proof[32:], [x]) is None
    assert ver.calldata(proof, [x + pyref.R]) is None and ver.calldata(proof, [-1]) is None
    assert ver.calldata(proof, [x]) is not None

    class Stub:
        def __init__(self, answer): self.answer = answer
        def verify(self, proof, public_inputs): return self.answer

    out = B200ProveOutput(ProverType.SP1, bytes(256), {})
    B200Backend(None, verifier=Stub(True)).verify(out, [1])
    with pytest.raises(eb.B200Error) as e:
        B200Backend(None, verifier=Stub(False)).verify(out, [1])
    assert e.value.kind == "Verification" and str(e.value).startswith("Verification error")
    with pytest.raises(eb.B200Error) as e:
        B200Backend(None).verify(out, [1])
    assert e.value.kind == "NotImplemented"


def test_timed_log_line_has_the_reference_fields():
    """`--timed` log line of /root/reference/crates/prover/src/prover.rs:106-118: id, proving_time_s, proving_time_ms."""
    from ethrex_b200.backend import log_proved
    line = log_proved(7, 1.2345)
    assert "id=7" in line and "proving_time_s=1 " in line and "proving_time_ms=1234 " in line and "Proved payload #7 in 1.23s" in line


def test_host_buffer_length_guard():
    from ethrex_b200.context import _need
    _need(bytes(64), 64, "x")
    _need(np.zeros(8, dtype=np.uint64), 64, "x")
    with pytest.raises(eb.B200Error, match="holds 63 bytes, the call needs 64"):
        _need(bytes(63), 64, "x")
    with pytest.raises(eb.B200Error):
        _need(np.zeros(7, dtype=np.uint64), 64, "x")


def test_kzg_host_bookkeeping_without_a_gpu():
    """ethrex_b200/kzg.py's scalar-field work (barycentric evaluation, quotient in evaluation form, the in-domain special
    case, the Fiat-Shamir challenge) with the two MSMs served in the exponent by a stand-in context: the commitment must be
    p(tau) G and the proof ((p(tau) - p(z)) / (tau - z)) G for a synthetic Lagrange setup (device half: tests/test_gpu_bls.py)."""
    import bls_ref as bls
    from ethrex_b200.kzg import KzgSettings, roots_of_unity_brp
    tau = 0x123456789ABCDEF0FEDCBA9876543210 % bls.R
    lag = bls.lagrange_setup_scalars(tau)

    class ExponentCtx:
        def bls12_381_g1_bases_upload(self, pts, n, flags=0): return 1
        def bases_precompute(self, h, c): pass
        def bases_free(self, h): pass
        def _dot(self, raw):
            return sum(int.from_bytes(raw[32 * i:32 * i + 32], "big") * l for i, l in enumerate(lag)) % bls.R
        def kzg_blob_to_commitment(self, h, blobs): return [bls.compress(bls.mul(self._dot(blobs[k * 131072:(k + 1) * 131072]), bls.G1)) for k in range(len(blobs) // 131072)]
        def bls12_381_g1_msm_resident(self, h, scalars, n, flags=0): return bls.compress(bls.mul(self._dot(scalars), bls.G1))

    st = KzgSettings(ExponentCtx(), bytes(48 * 4096), precompute=True)
    rng = np.random.default_rng(7)
    vals = [int.from_bytes(rng.bytes(32), "big") % bls.R for _ in range(4096)]
    blob = b"".join(v.to_bytes(32, "big") for v in vals)
    p_tau = sum(v * l for v, l in zip(vals, lag)) % bls.R
    c, proof = st.blob_to_kzg_commitment_and_proof(blob)
    assert c == bls.compress(bls.mul(p_tau, bls.G1))
    z = st.compute_challenge(blob, c)
    proof_z, y = st.compute_kzg_proof(blob, z)
    assert proof_z == proof
    # y = p(z): check against the definition through the Lagrange basis at z
    lag_z = bls.lagrange_setup_scalars(z)
    assert y == sum(v * l for v, l in zip(vals, lag_z)) % bls.R
    assert proof == bls.compress(bls.mul((p_tau - y) * pow((tau - z) % bls.R, -1, bls.R) % bls.R, bls.G1))
    zr = roots_of_unity_brp()[1234]
    proof_r, yr = st.compute_kzg_proof(blob, zr)
    assert yr == vals[1234]
    assert proof_r == bls.compress(bls.mul((p_tau - yr) * pow((tau - zr) % bls.R, -1, bls.R) % bls.R, bls.G1))
    with pytest.raises(ValueError):
        st.compute_kzg_proof(bls.R.to_bytes(32, "big") + blob[32:], 5)
