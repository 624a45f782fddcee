//! `B200Crypto`: the three BN254 calls of the reference's `Crypto` trait
//! (`crates/common/crypto/provider.rs:201-330`) on the GPU.  The trait is one item per call; a provider that wants
//! throughput collects the items of a block (or of the batch being proved) and calls the `*_batch` wrappers of
//! [`crate::ffi::B200zk`] directly -- the single-item methods below are the drop-in form.
//!
//! Error mapping follows the reference: the levm wrappers reject coordinates >= p before the curve call
//! (`crates/vm/levm/src/precompiles.rs:801-820`, `PrecompileError::CoordinateExceedsFieldModulus`), the provider
//! reports points off the curve as `CryptoError::InvalidPoint`.
use ethrex_crypto::{Crypto, CryptoError};

use crate::ffi::{global, ItemStatus};

#[derive(Debug, Default, Clone, Copy)]
pub struct B200Crypto;

fn device_error<E: std::fmt::Display>(e: E) -> CryptoError {
    CryptoError::Other(e.to_string())
}

fn item_error(status: ItemStatus, what: &'static str) -> CryptoError {
    match status {
        ItemStatus::NotInField => CryptoError::InvalidInput("coordinate exceeds the field modulus"),
        _ => CryptoError::InvalidPoint(what),
    }
}

impl Crypto for B200Crypto {
    fn bn254_g1_add(&self, p1: &[u8], p2: &[u8]) -> Result<[u8; 64], CryptoError> {
        let (a, b) = (p1.get(..64).ok_or(CryptoError::InvalidInput("G1 point must be 64 bytes"))?, p2.get(..64).ok_or(CryptoError::InvalidInput("G1 point must be 64 bytes"))?);
        let mut gpu = global().map_err(device_error)?.lock().map_err(device_error)?;
        let (out, st) = gpu.bn254_g1_add_batch(a, b).map_err(device_error)?;
        match st.first().copied() {
            Some(ItemStatus::Ok | ItemStatus::OkIdentity) => <[u8; 64]>::try_from(out.as_slice()).map_err(device_error),
            Some(bad) => Err(item_error(bad, "G1 point not on curve")),
            None => Err(CryptoError::Other("b200zk returned no status".to_string())),
        }
    }

    fn bn254_g1_mul(&self, point: &[u8], scalar: &[u8]) -> Result<[u8; 64], CryptoError> {
        let (p, k) = (point.get(..64).ok_or(CryptoError::InvalidInput("invalid input length"))?, scalar.get(..32).ok_or(CryptoError::InvalidInput("invalid input length"))?);
        let mut gpu = global().map_err(device_error)?.lock().map_err(device_error)?;
        let (out, st) = gpu.bn254_g1_mul_batch(p, k).map_err(device_error)?;
        match st.first().copied() {
            Some(ItemStatus::Ok | ItemStatus::OkIdentity) => <[u8; 64]>::try_from(out.as_slice()).map_err(device_error),
            Some(bad) => Err(item_error(bad, "G1 point not on curve")),
            None => Err(CryptoError::Other("b200zk returned no status".to_string())),
        }
    }

    fn bn254_pairing_check(&self, pairs: &[(&[u8], &[u8])]) -> Result<bool, CryptoError> {
        let mut calldata = Vec::with_capacity(pairs.len().saturating_mul(192));
        for (g1, g2) in pairs {
            calldata.extend_from_slice(g1.get(..64).ok_or(CryptoError::InvalidInput("G1 must be 64 bytes"))?);
            calldata.extend_from_slice(g2.get(..128).ok_or(CryptoError::InvalidInput("G2 must be 128 bytes"))?);
        }
        let mut gpu = global().map_err(device_error)?.lock().map_err(device_error)?;
        let res = gpu.bn254_pairing_check_batch(&[calldata.as_slice()]).map_err(device_error)?;
        match res.first() {
            Some(Ok(v)) => Ok(*v),
            Some(Err(bad)) => Err(item_error(*bad, "G1/G2 not on BN254 curve")),
            None => Err(CryptoError::Other("b200zk returned no result".to_string())),
        }
    }
}
