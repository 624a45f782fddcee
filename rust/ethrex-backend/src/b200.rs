//! `B200Backend`: a `ProverBackend` (`crates/prover/src/backend/mod.rs:81-147`) whose BN254 arithmetic -- the
//! G1/G2 multi-scalar multiplications and Fr NTTs of the Groth16 wrap that the reference reaches only inside
//! third-party SDKs when the coordinator asks for `ProofFormat::Groth16`
//! (`crates/l2/sequencer/proof_coordinator.rs:252-256`, `crates/prover/src/backend/sp1.rs:97-134`) -- runs on
//! a B200 through `libb200zk.so`.
//!
//! Scope: this backend does not introduce a new proof system.  It reports the
//! `ProverType` of the zkVM whose on-chain Groth16 verifier it targets and replaces the *commitment
//! arithmetic* of that zkVM's wrap stage.  The wrap circuit itself (R1CS + proving key) belongs to the
//! zkVM SDK and is injected through [`WrapCircuit`]; without one, `prove` returns
//! `BackendError::NotImplemented` for `ProofFormat::Groth16`, exactly like a backend built without its SDK.
//! Drop this file in as `crates/prover/src/backend/b200.rs` (wiring in INTEGRATION.md).
use std::time::{Duration, Instant};

use ethrex_common::types::prover::{ProofBytes, ProofFormat, ProverOutput, ProverType};
use ethrex_guest_program::input::ProgramInput;
use rkyv::rancor::Error;
use tracing::info;

use ethrex_prover::backend::{BackendError, ExecBackend, ProverBackend};

use crate::ffi;
use b200zk_sys::{B200ZK_NTT_COSET, B200ZK_NTT_INVERSE};

/// What the zkVM SDK has to provide for the Groth16 wrap: the witness vector and the proving-key columns.
/// All buffers use the library's native formats (ark-ff Montgomery limbs for points, canonical limbs for
/// scalars), i.e. the in-memory form the arkworks-based wrap provers already hold them in.
pub trait WrapCircuit: Send + Sync {
    /// `ProverType` whose on-chain verifier checks the produced proof (`crates/common/types/prover.rs:7-23`).
    fn prover_type(&self) -> ProverType;
    /// log2 of the evaluation-domain size.
    fn domain_log2(&self) -> u32;
    /// Full witness assignment (public inputs first), canonical 32-byte little-endian scalars.
    fn witness(&self, serialized_input: &[u8]) -> Result<Vec<u8>, BackendError>;
    /// Proving-key columns as native affine points: A (G1), B (G1), B (G2), K/L (G1), H (G1).
    fn pk_a_g1(&self) -> &[u8];
    fn pk_b_g1(&self) -> &[u8];
    fn pk_b_g2(&self) -> &[u8];
    fn pk_l_g1(&self) -> &[u8];
    fn pk_h_g1(&self) -> &[u8];
    /// Evaluations of A*w, B*w, C*w over the domain (Montgomery limbs), from which the quotient H is built.
    fn abc_evaluations(&self, witness: &[u8]) -> Result<[Vec<u8>; 3], BackendError>;
    /// (a .* b - c) ./ Z_H on the coset, element-wise, done by the SDK's field code.
    fn quotient_on_coset(&self, abc_coset: &mut [Vec<u8>; 3]) -> Result<Vec<u8>, BackendError>;
    /// Assemble (A, B, C) with the blinding terms and the verifier's byte order.
    fn assemble(&self, commitments: &Groth16Commitments) -> Result<Vec<u8>, BackendError>;
    /// The ecpairing calldata of the Groth16 verification equation for `proof`
    /// (`-A | B | alpha | beta | IC(public inputs) | gamma | C | delta`, 4 x 192 bytes), when the SDK exposes its
    /// verifying key; `None` makes `verify` answer "not implemented", like the reference's default.
    fn verifier_calldata(&self, _proof: &[u8]) -> Option<Vec<u8>> {
        None
    }
}

/// The five commitments of a Groth16 proof before blinding, EIP-196/197 encoded.
pub struct Groth16Commitments {
    pub a_g1: [u8; 64],
    pub b_g1: [u8; 64],
    pub b_g2: [u8; 128],
    pub l_g1: [u8; 64],
    pub h_g1: [u8; 64],
}

pub struct B200ProveOutput {
    pub prover_type: ProverType,
    pub proof: Vec<u8>,
}

#[derive(Default)]
pub struct B200Backend {
    circuit: Option<Box<dyn WrapCircuit>>,
}

impl B200Backend {
    pub fn new() -> Self {
        Self { circuit: None }
    }

    pub fn with_circuit(circuit: Box<dyn WrapCircuit>) -> Self {
        Self { circuit: Some(circuit) }
    }

    /// The hot path: 3 iNTT + 3 coset NTT + 1 coset iNTT, then 4 G1 MSMs and 1 G2 MSM on the GPU.
    fn commit(circuit: &dyn WrapCircuit, serialized: &[u8]) -> Result<Groth16Commitments, BackendError> {
        let gpu = ffi::global()?;
        let mut gpu = gpu.lock().map_err(|_| BackendError::proving("b200zk context poisoned"))?;
        let k = circuit.domain_log2();
        let witness = circuit.witness(serialized)?;
        let mut abc = circuit.abc_evaluations(&witness)?;
        for poly in abc.iter_mut() {
            gpu.fr_ntt(poly, k, B200ZK_NTT_INVERSE, None)?; // evaluations -> coefficients
            gpu.fr_ntt(poly, k, B200ZK_NTT_COSET, None)?; // coefficients -> coset evaluations
        }
        let mut h = circuit.quotient_on_coset(&mut abc)?;
        gpu.fr_ntt(&mut h, k, B200ZK_NTT_INVERSE | B200ZK_NTT_COSET, None)?;
        Ok(Groth16Commitments {
            a_g1: gpu.g1_msm(circuit.pk_a_g1(), &witness, 0)?,
            b_g1: gpu.g1_msm(circuit.pk_b_g1(), &witness, 0)?,
            b_g2: gpu.g2_msm(circuit.pk_b_g2(), &witness, 0)?,
            l_g1: gpu.g1_msm(circuit.pk_l_g1(), &witness, 0)?,
            h_g1: gpu.g1_msm(circuit.pk_h_g1(), &h, b200zk_sys::B200ZK_SCALARS_MONT)?,
        })
    }
}

impl ProverBackend for B200Backend {
    type ProofOutput = B200ProveOutput;
    type SerializedInput = Vec<u8>;

    fn prover_type(&self) -> ProverType {
        self.circuit.as_ref().map_or(ProverType::Exec, |c| c.prover_type())
    }

    fn serialize_input(&self, input: &ProgramInput) -> Result<Self::SerializedInput, BackendError> {
        // same wire form the zkVM backends feed their guests (sp1.rs:145-150)
        let bytes = rkyv::to_bytes::<Error>(input).map_err(BackendError::serialization)?;
        Ok(bytes.to_vec())
    }

    fn execute(&self, input: ProgramInput) -> Result<(), BackendError> {
        ExecBackend::new().execute(input)
    }

    fn prove(&self, input: ProgramInput, format: ProofFormat) -> Result<Self::ProofOutput, BackendError> {
        let serialized = self.serialize_input(&input)?;
        // the guest program must accept the batch before anything is committed to
        ExecBackend::new().execute(input)?;
        match (format, self.circuit.as_deref()) {
            (ProofFormat::Groth16, Some(circuit)) => {
                let commitments = Self::commit(circuit, &serialized)?;
                Ok(B200ProveOutput { prover_type: circuit.prover_type(), proof: circuit.assemble(&commitments)? })
            }
            (ProofFormat::Groth16, None) => Err(BackendError::not_implemented(
                "b200 backend built without a wrap circuit: ProofFormat::Groth16 needs a zkVM SDK's proving key",
            )),
            (ProofFormat::Compressed, _) => Err(BackendError::not_implemented(
                "b200 backend accelerates the BN254 wrap only; Compressed (STARK) proofs come from the zkVM backend",
            )),
        }
    }

    /// The pairing check of the Groth16 verification equation, on the device (`b200zk_bn254_pairing_check_batch`).
    fn verify(&self, proof: &Self::ProofOutput) -> Result<(), BackendError> {
        let calldata = self
            .circuit
            .as_ref()
            .and_then(|c| c.verifier_calldata(&proof.proof))
            .ok_or_else(BackendError::verify_not_supported)?;
        let mut gpu = ffi::global()?.lock().map_err(|e| BackendError::verification(e.to_string()))?;
        match gpu.bn254_pairing_check_batch(&[calldata.as_slice()])?.first() {
            Some(Ok(true)) => Ok(()),
            Some(Ok(false)) => Err(BackendError::verification("Groth16 pairing check failed")),
            Some(Err(status)) => Err(BackendError::verification(format!("malformed proof point: {status:?}"))),
            None => Err(BackendError::verification("b200zk returned no result")),
        }
    }

    fn to_proof_bytes(&self, proof: Self::ProofOutput, _format: ProofFormat) -> Result<ProverOutput, BackendError> {
        Ok(ProverOutput::Proof(ProofBytes { prover_type: proof.prover_type, proof: proof.proof }))
    }

    fn prove_timed(&self, input: ProgramInput, format: ProofFormat) -> Result<(Self::ProofOutput, Duration), BackendError> {
        let start = Instant::now();
        let proof = self.prove(input, format)?;
        let elapsed = start.elapsed();
        info!("b200 backend proved in {:.2?}", elapsed);
        Ok((proof, elapsed))
    }
}
