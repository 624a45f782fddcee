"""Error type mirroring `BackendError` (/root/reference/crates/prover/src/backend/error.rs:3-51):
variants Serialization / Execution / Proving / Verification / ProofConversion / NotImplemented, each
carrying a message; constructors are lower-case class methods like the Rust helper fns."""
from __future__ import annotations


class B200Error(Exception):
    KINDS = ("Serialization", "Execution", "Proving", "Verification", "ProofConversion", "NotImplemented")

    def __init__(self, kind: str, message: str, status: int | None = None):
        assert kind in self.KINDS
        self.kind, self.message, self.status = kind, message, status
        # same Display strings as error.rs:4-20
        prefix = {
            "Serialization": "Serialization error",
            "Execution": "Execution error",
            "Proving": "Proving error",
            "Verification": "Verification error",
            "ProofConversion": "Proof conversion error",
            "NotImplemented": "Not implemented",
        }[kind]
        super().__init__(f"{prefix}: {message}")

    @classmethod
    def serialization(cls, m, status=None): return cls("Serialization", str(m), status)
    @classmethod
    def execution(cls, m, status=None): return cls("Execution", str(m), status)
    @classmethod
    def proving(cls, m, status=None): return cls("Proving", str(m), status)
    @classmethod
    def verification(cls, m, status=None): return cls("Verification", str(m), status)
    @classmethod
    def proof_conversion(cls, m, status=None): return cls("ProofConversion", str(m), status)
    @classmethod
    def not_implemented(cls, m, status=None): return cls("NotImplemented", str(m), status)
    @classmethod
    def verify_not_supported(cls): return cls("NotImplemented", "Verify not implemented for this backend")


class NoDeviceError(B200Error):
    def __init__(self, message: str):
        super().__init__("Proving", message, 6)


def status_to_error(status: int, message: str) -> B200Error:
    """C status -> BackendError variant (the mapping the Rust shim applies)."""
    if status in (2, 3, 4):  # not in field / not on curve / invalid argument: malformed input
        return B200Error.serialization(message, status)
    if status == 8:
        return B200Error.not_implemented(message, status)
    return B200Error.proving(message, status)
