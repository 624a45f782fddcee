//! Safe wrapper over the raw C ABI (`b200zk-sys`).  One process-global context, like the reference's
//! `static PROVER_SETUP: OnceLock<ProverSetup>` (`crates/prover/src/backend/sp1.rs:30,93-95`).
//! Obeys the prover crate's lint policy (`crates/prover/Cargo.toml:72-80`): no unwrap / expect / panic /
//! indexing / `as`.
use std::ffi::CStr;
use std::ptr::NonNull;
use std::sync::{Mutex, OnceLock};

use b200zk_sys as sys;
use ethrex_prover::backend::BackendError;

/// Owned `b200zk_ctx*`.  The library serialises work on its own stream; the `Mutex` in [`global`] makes the
/// handle usable from the prover actor's blocking thread (`crates/prover/src/prover.rs:240-251`) or any other.
pub struct B200zk {
    ctx: NonNull<sys::b200zk_ctx>,
}

// SAFETY: the context owns only device resources and is never aliased outside the Mutex in `global()`.
unsafe impl Send for B200zk {}

static GLOBAL: OnceLock<Result<Mutex<B200zk>, String>> = OnceLock::new();

/// Lazily initialised process-global context on the device selected by `CUDA_VISIBLE_DEVICES`.
pub fn global() -> Result<&'static Mutex<B200zk>, BackendError> {
    GLOBAL
        .get_or_init(|| B200zk::new(0).map(Mutex::new).map_err(|e| e.to_string()))
        .as_ref()
        .map_err(BackendError::proving)
}

fn status_message(ctx: Option<&B200zk>, status: i32) -> String {
    // SAFETY: both functions return NUL-terminated strings owned by the library / the context.
    let base = unsafe { CStr::from_ptr(sys::b200zk_strerror(status)) }.to_string_lossy().into_owned();
    match ctx {
        Some(c) => {
            let detail = unsafe { CStr::from_ptr(sys::b200zk_last_error(c.ctx.as_ptr())) }.to_string_lossy().into_owned();
            format!("b200zk status {status}: {base} ({detail})")
        }
        None => format!("b200zk status {status}: {base}"),
    }
}

/// C status -> `BackendError` (the table INTEGRATION.md documents): malformed input is a serialization
/// error, everything the device reports is a proving error.
fn check(ctx: &B200zk, status: i32) -> Result<bool, BackendError> {
    match status {
        sys::B200ZK_OK => Ok(false),
        sys::B200ZK_OK_INFINITY => Ok(true),
        sys::B200ZK_ERR_NOT_IN_FIELD | sys::B200ZK_ERR_NOT_ON_CURVE | sys::B200ZK_ERR_INVALID_ARG => {
            Err(BackendError::serialization(status_message(Some(ctx), status)))
        }
        sys::B200ZK_ERR_UNSUPPORTED => Err(BackendError::not_implemented(status_message(Some(ctx), status))),
        other => Err(BackendError::proving(status_message(Some(ctx), other))),
    }
}

impl B200zk {
    pub fn new(device: i32) -> Result<Self, BackendError> {
        let mut raw: *mut sys::b200zk_ctx = std::ptr::null_mut();
        // SAFETY: `raw` is a valid out-pointer.
        let status = unsafe { sys::b200zk_init(device, &mut raw) };
        match NonNull::new(raw) {
            Some(ctx) if status == sys::B200ZK_OK => Ok(Self { ctx }),
            _ => Err(BackendError::proving(status_message(None, status))),
        }
    }

    /// sum_i scalars[i] * bases[i] over BN254 G1.  `points`: n x 64 bytes, `scalars`: n x 32 bytes, formats per
    /// `flags` (see include/b200zk.h).  Returns the 64-byte EIP-196 encoding.
    pub fn g1_msm(&mut self, points: &[u8], scalars: &[u8], flags: u32) -> Result<[u8; 64], BackendError> {
        let n = scalars.len() / 32;
        if points.len() / 64 < n {
            return Err(BackendError::serialization("g1_msm: fewer points than scalars"));
        }
        let mut out = [0u8; 64];
        // SAFETY: lengths checked above; buffers outlive the (synchronous) call.
        let status = unsafe {
            sys::b200zk_g1_msm(self.ctx.as_ptr(), points.as_ptr().cast(), scalars.as_ptr().cast(), n, flags, out.as_mut_ptr())
        };
        check(self, status)?;
        Ok(out)
    }

    pub fn g2_msm(&mut self, points: &[u8], scalars: &[u8], flags: u32) -> Result<[u8; 128], BackendError> {
        let n = scalars.len() / 32;
        if points.len() / 128 < n {
            return Err(BackendError::serialization("g2_msm: fewer points than scalars"));
        }
        let mut out = [0u8; 128];
        // SAFETY: as above.
        let status = unsafe {
            sys::b200zk_g2_msm(self.ctx.as_ptr(), points.as_ptr().cast(), scalars.as_ptr().cast(), n, flags, out.as_mut_ptr())
        };
        check(self, status)?;
        Ok(out)
    }

    /// In-place NTT of `data` (2^log_n x 32 bytes).
    pub fn fr_ntt(&mut self, data: &mut [u8], log_n: u32, flags: u32, coset_gen: Option<&[u8; 32]>) -> Result<(), BackendError> {
        let want = 32usize.checked_shl(log_n).ok_or_else(|| BackendError::serialization("fr_ntt: log_n too large"))?;
        if data.len() != want {
            return Err(BackendError::serialization("fr_ntt: buffer length is not 32 * 2^log_n"));
        }
        let cg = coset_gen.map_or(std::ptr::null(), |g| g.as_ptr());
        // SAFETY: length checked above.
        let status = unsafe { sys::b200zk_fr_ntt(self.ctx.as_ptr(), data.as_mut_ptr().cast(), log_n, flags, cg) };
        check(self, status).map(|_| ())
    }

    /// Uploads a proving-key column once; later proofs only ship scalars.
    pub fn g1_bases_upload(&mut self, points: &[u8], flags: u32) -> Result<u64, BackendError> {
        let mut handle = 0u64;
        // SAFETY: slice is valid for points.len() bytes.
        let status = unsafe { sys::b200zk_g1_bases_upload(self.ctx.as_ptr(), points.as_ptr().cast(), points.len() / 64, flags, &mut handle) };
        check(self, status)?;
        Ok(handle)
    }

    pub fn g2_bases_upload(&mut self, points: &[u8], flags: u32) -> Result<u64, BackendError> {
        let mut handle = 0u64;
        // SAFETY: slice is valid for points.len() bytes.
        let status = unsafe { sys::b200zk_g2_bases_upload(self.ctx.as_ptr(), points.as_ptr().cast(), points.len() / 128, flags, &mut handle) };
        check(self, status)?;
        Ok(handle)
    }

    /// One-off per proving-key column: expand the resident bases into their window multiples (`window_bits` 0 = automatic).
    pub fn bases_precompute(&mut self, handle: u64, window_bits: u32) -> Result<(), BackendError> {
        // SAFETY: plain value arguments.
        let status = unsafe { sys::b200zk_bases_precompute(self.ctx.as_ptr(), handle, window_bits) };
        check(self, status).map(|_| ())
    }

    pub fn bases_free(&mut self, handle: u64) -> Result<(), BackendError> {
        // SAFETY: plain value arguments.
        let status = unsafe { sys::b200zk_bases_free(self.ctx.as_ptr(), handle) };
        check(self, status).map(|_| ())
    }

    /// Selects the 2^28-th root of unity the NTT domains derive from (`None` = ark/gnark default; halo2curves'
    /// value comes from `b200zk_ntt_root_preset(1, ..)`), for wraps built on another FFT convention (`openvm.rs:52-56`).
    pub fn set_ntt_root(&mut self, root_le: Option<&[u8; 32]>) -> Result<(), BackendError> {
        let p = root_le.map_or(std::ptr::null(), |g| g.as_ptr());
        // SAFETY: NULL or 32 valid bytes.
        let status = unsafe { sys::b200zk_set_ntt_root(self.ctx.as_ptr(), p) };
        check(self, status).map(|_| ())
    }

    /// The whole Groth16 prove arithmetic (quotient NTTs, five MSMs over the RESIDENT proving key, C = L + H) in one
    /// call with one synchronisation.  `witness`: canonical LE scalars; `a`, `b`, `c`: (A z), (B z), (C z) on the
    /// domain, Montgomery LE, 2^log_n elements each.  Returns (A | B2 | C, [B]1).
    pub fn groth16_commit(&mut self, pk: &sys::b200zk_groth16_pk, witness: &[u8], a: &mut [u8], b: &mut [u8], c: &mut [u8]) -> Result<([u8; 256], [u8; 64]), BackendError> {
        let n_bytes = 32usize.checked_shl(pk.log_n).ok_or_else(|| BackendError::serialization("groth16_commit: log_n too large"))?;
        if a.len() != n_bytes || b.len() != n_bytes || c.len() != n_bytes {
            return Err(BackendError::serialization("groth16_commit: evaluation vectors must hold 2^log_n elements"));
        }
        // the witness must reach the end of every column that multiplies it (columns 0..3)
        let mut wit_end = 0u64;
        for ((h, cnt), off) in pk.handle.iter().zip(pk.count.iter()).zip(pk.offset.iter()).take(4) {
            if *h != 0 {
                wit_end = wit_end.max(off.checked_add(*cnt).ok_or_else(|| BackendError::serialization("groth16_commit: column range overflows"))?);
            }
        }
        let have = u64::try_from(witness.len() / 32).map_err(BackendError::serialization)?;
        if have < wit_end {
            return Err(BackendError::serialization(format!("groth16_commit: witness holds {have} scalars, the proving key multiplies {wit_end}")));
        }
        let mut proof = [0u8; 256];
        let mut b1 = [0u8; 64];
        // SAFETY: lengths checked above; host buffers outlive the synchronous call; NULL stream = the context's own.
        let status = unsafe {
            sys::b200zk_groth16_commit(self.ctx.as_ptr(), pk, witness.as_ptr().cast(), a.as_mut_ptr().cast(), b.as_mut_ptr().cast(), c.as_mut_ptr().cast(), 0,
                                       std::ptr::null_mut(), proof.as_mut_ptr(), b1.as_mut_ptr())
        };
        check(self, status)?;
        Ok((proof, b1))
    }

    pub fn g1_msm_resident(&mut self, handle: u64, scalars: &[u8], flags: u32) -> Result<[u8; 64], BackendError> {
        let mut out = [0u8; 64];
        // SAFETY: slice valid; n derived from its length.
        let status = unsafe {
            sys::b200zk_g1_msm_resident(self.ctx.as_ptr(), handle, scalars.as_ptr().cast(), scalars.len() / 32, flags, out.as_mut_ptr())
        };
        check(self, status)?;
        Ok(out)
    }
}

/// Per-item outcome of the batched precompile calls (include/b200zk.h: 0 ok, 1 ok-identity, 2 coordinate >= p,
/// 3 not on the curve / not in the subgroup).
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum ItemStatus {
    Ok,
    OkIdentity,
    NotInField,
    NotOnCurve,
}

impl ItemStatus {
    fn from_code(code: u8) -> Self {
        match code {
            0 => Self::Ok,
            1 => Self::OkIdentity,
            2 => Self::NotInField,
            _ => Self::NotOnCurve,
        }
    }
}

impl B200zk {
    /// `count` independent ecAdd items: `a`, `b` = count x 64 bytes.  Returns (count x 64 result bytes, per-item status).
    pub fn bn254_g1_add_batch(&mut self, a: &[u8], b: &[u8]) -> Result<(Vec<u8>, Vec<ItemStatus>), BackendError> {
        if a.len() != b.len() || a.len() % 64 != 0 {
            return Err(BackendError::serialization("bn254_g1_add_batch: inputs must be equal multiples of 64 bytes"));
        }
        let count = a.len() / 64;
        let mut out = vec![0u8; a.len()];
        let mut st = vec![0u8; count];
        // SAFETY: all four buffers hold `count` items of the documented sizes and outlive the synchronous call.
        let status = unsafe { sys::b200zk_bn254_g1_add_batch(self.ctx.as_ptr(), a.as_ptr(), b.as_ptr(), count, out.as_mut_ptr(), st.as_mut_ptr()) };
        check(self, status)?;
        Ok((out, st.into_iter().map(ItemStatus::from_code).collect()))
    }

    /// `count` independent ecMul items: `points` = count x 64 bytes, `scalars` = count x 32 bytes (big-endian).
    pub fn bn254_g1_mul_batch(&mut self, points: &[u8], scalars: &[u8]) -> Result<(Vec<u8>, Vec<ItemStatus>), BackendError> {
        if points.len() % 64 != 0 || scalars.len() != points.len() / 2 {
            return Err(BackendError::serialization("bn254_g1_mul_batch: need 64 bytes of point and 32 of scalar per item"));
        }
        let count = points.len() / 64;
        let mut out = vec![0u8; points.len()];
        let mut st = vec![0u8; count];
        // SAFETY: as above.
        let status = unsafe { sys::b200zk_bn254_g1_mul_batch(self.ctx.as_ptr(), points.as_ptr(), scalars.as_ptr(), count, out.as_mut_ptr(), st.as_mut_ptr()) };
        check(self, status)?;
        Ok((out, st.into_iter().map(ItemStatus::from_code).collect()))
    }

    /// Several ecPairing checks in one launch.  `checks[i]` is the precompile's calldata (k x 192 bytes).
    /// Returns, per check, `Ok(true/false)` or the input error.
    pub fn bn254_pairing_check_batch(&mut self, checks: &[&[u8]]) -> Result<Vec<Result<bool, ItemStatus>>, BackendError> {
        let mut blob = Vec::new();
        let mut offsets = Vec::with_capacity(checks.len().saturating_add(1));
        offsets.push(0u32);
        for cd in checks {
            if cd.len() % 192 != 0 {
                return Err(BackendError::serialization("bn254_pairing_check_batch: calldata must be a multiple of 192 bytes"));
            }
            blob.extend_from_slice(cd);
            let pairs = u32::try_from(blob.len() / 192).map_err(|_| BackendError::serialization("bn254_pairing_check_batch: too many pairs"))?;
            offsets.push(pairs);
        }
        let count = checks.len();
        let mut res = vec![0u8; count];
        let mut st = vec![0u8; count];
        // SAFETY: `offsets` has count + 1 entries, `blob` holds offsets[count] pairs, outputs hold `count` bytes.
        let status = unsafe {
            sys::b200zk_bn254_pairing_check_batch(self.ctx.as_ptr(), blob.as_ptr(), offsets.as_ptr(), count, res.as_mut_ptr(), st.as_mut_ptr())
        };
        check(self, status)?;
        Ok(res
            .into_iter()
            .zip(st)
            .map(|(r, s)| match ItemStatus::from_code(s) {
                ItemStatus::Ok | ItemStatus::OkIdentity => Ok(r == 1),
                bad => Err(bad),
            })
            .collect())
    }
}

impl Drop for B200zk {
    fn drop(&mut self) {
        // SAFETY: ctx came from b200zk_init and is dropped exactly once.
        unsafe { sys::b200zk_destroy(self.ctx.as_ptr()) }
    }
}
