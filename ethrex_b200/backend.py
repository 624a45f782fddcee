"""`B200Backend`: the Python mirror of `rust/ethrex-backend/src/b200.rs`, i.e. of the `ProverBackend` trait
(/root/reference/crates/prover/src/backend/mod.rs:81-147): same six methods plus the `*_timed` defaults, same
error variants (`BackendError`, error.rs:3-51), same output types (`ProverOutput::Proof(ProofBytes{..})`,
/root/reference/crates/common/types/prover.rs:50-70,103-110).

`prove(input, ProofFormat.GROTH16)` runs the wrap's commitment arithmetic -- 7 NTTs, a pointwise quotient,
4 G1 MSMs and 1 G2 MSM -- on the GPU over a `SyntheticWrapCircuit` (the real wrap circuit belongs to the zkVM
SDKs and is not in the reference tree; the STARK stage is out of scope).  Without a circuit it raises
`NotImplemented`, like a backend compiled without its SDK.
"""
from __future__ import annotations

import enum
import hashlib
import json
import time
from dataclasses import dataclass

from .errors import B200Error


class ProofFormat(enum.Enum):  # crates/common/types/prover.rs:103-110 (Groth16 is the default)
    GROTH16 = "Groth16"
    COMPRESSED = "Compressed"


class ProverType(enum.Enum):  # crates/common/types/prover.rs:7-12
    EXEC = "Exec"
    RISC0 = "RISC0"
    SP1 = "SP1"
    TDX = "TDX"


@dataclass
class ProofBytes:  # prover.rs:50-54
    prover_type: ProverType
    proof: bytes


@dataclass
class ProverOutput:  # prover.rs:63-70, the `Proof` variant
    proof_bytes: ProofBytes

    def prover_type(self) -> ProverType:
        return self.proof_bytes.prover_type


@dataclass
class B200ProveOutput:
    prover_type: ProverType
    proof: bytes
    commitments: dict


class BackendType(enum.Enum):  # backend/mod.rs:41-55, with the new variant
    EXEC = "exec"
    B200 = "b200"

    @classmethod
    def from_str(cls, s: str) -> "BackendType":  # backend/mod.rs:58-76
        try:
            return cls(s.lower())
        except ValueError:
            raise ValueError("Invalid backend") from None


def serialize_program_input(program_input) -> bytes:
    """Stand-in for `rkyv::to_bytes(ProgramInput)` (sp1.rs:145-150): bytes pass through, anything else is
    canonical JSON.  The Rust shim uses rkyv; the pipeline only needs *a* deterministic byte string."""
    if isinstance(program_input, (bytes, bytearray, memoryview)):
        return bytes(program_input)
    try:
        return json.dumps(program_input, sort_keys=True, separators=(",", ":")).encode()
    except (TypeError, ValueError) as e:
        raise B200Error.serialization(e) from None


class B200Backend:
    def __init__(self, ctx=None, circuit=None, prover_type: ProverType = ProverType.SP1, verifier=None):
        """verifier: an ethrex_b200.groth16.Groth16Verifier for the circuit's verifying key (optional; without it
        `verify` answers like a backend built without its SDK's verifier)."""
        self.ctx, self.circuit, self._prover_type, self.verifier = ctx, circuit, prover_type, verifier

    # ---- ProverBackend ----
    def prover_type(self) -> ProverType:
        return self._prover_type if self.circuit is not None else ProverType.EXEC

    def serialize_input(self, program_input) -> bytes:
        return serialize_program_input(program_input)

    def execute(self, program_input) -> None:
        """The guest program re-executes the batch on the CPU (ExecBackend, exec.rs:25-54); that EVM path is out
        of scope here, so `execute` only checks that the input serializes."""
        self.serialize_input(program_input)

    def prove(self, program_input, proof_format: ProofFormat = ProofFormat.GROTH16, msm=None) -> B200ProveOutput:
        serialized = self.serialize_input(program_input)
        self.execute(program_input)
        if proof_format is ProofFormat.COMPRESSED:
            raise B200Error.not_implemented("b200 backend accelerates the BN254 wrap only; Compressed (STARK) proofs come from the zkVM backend")
        if self.circuit is None:
            raise B200Error.not_implemented("b200 backend built without a wrap circuit: ProofFormat::Groth16 needs a zkVM SDK's proving key")
        try:
            if msm is not None or not hasattr(self.circuit, "prove_device"):
                proof, commitments = self.circuit.prove_separate(serialized, msm)
            else:  # ONE C-ABI call per proof (b200zk_groth16_commit), like rust/ethrex-backend/src/b200.rs
                proof, b_g1 = self.circuit.prove_device(serialized)
                commitments = {"a_g1": proof[:64], "b_g2": proof[64:192], "c_g1": proof[192:], "b_g1": b_g1}
        except B200Error:
            raise
        except Exception as e:  # noqa: BLE001  (mirror of `.map_err(BackendError::proving)`)
            raise B200Error.proving(e) from e
        return B200ProveOutput(self.prover_type(), proof, commitments)

    def verify(self, proof: B200ProveOutput, public_inputs=None) -> None:
        """ProverBackend::verify (backend/mod.rs:116-117): Ok(()) or BackendError::Verification."""
        if self.verifier is None or public_inputs is None:
            raise B200Error.verify_not_supported()
        if not self.verifier.verify(proof.proof, public_inputs):
            raise B200Error.verification("Groth16 pairing check failed")

    def to_proof_bytes(self, proof: B200ProveOutput, proof_format: ProofFormat = ProofFormat.GROTH16) -> ProverOutput:
        if proof_format is not ProofFormat.GROTH16:
            raise B200Error.proof_conversion("only Groth16 proofs are produced")
        return ProverOutput(ProofBytes(proof.prover_type, proof.proof))

    # ---- default *_timed methods (backend/mod.rs:94-146) ----
    def serialize_input_timed(self, program_input):
        t0 = time.perf_counter()
        s = self.serialize_input(program_input)
        return s, time.perf_counter() - t0

    def execute_timed(self, program_input) -> float:
        t0 = time.perf_counter()
        self.execute(program_input)
        return time.perf_counter() - t0

    def prove_timed(self, program_input, proof_format: ProofFormat = ProofFormat.GROTH16, msm=None):
        t0 = time.perf_counter()
        proof = self.prove(program_input, proof_format, msm)
        return proof, time.perf_counter() - t0


def log_proved(payload_id: int, elapsed_s: float, logger=None) -> str:
    """The reference's `--timed` log line (`Prover::poll_endpoints`, /root/reference/crates/prover/src/prover.rs:106-118):
    fields `id`, `proving_time_s`, `proving_time_ms` and the message `Proved payload #<id> in <elapsed>`, kept
    identical so that the reference's log scrapers (docs/l2 benchmarks) read this backend's timings unchanged."""
    line = f"id={payload_id} proving_time_s={int(elapsed_s)} proving_time_ms={int(elapsed_s * 1000)} Proved payload #{payload_id} in {elapsed_s:.2f}s"
    if logger is not None:
        logger.info(line)
    return line


def proof_digest(proof: bytes) -> str:
    return hashlib.sha256(proof).hexdigest()
