/* b200zk.h -- frozen C ABI of libb200zk.so, the B200 (sm_100a) BN254 MSM + Fr NTT backend.
 *
 * This is the drop-in boundary for ethrex's L2 prover hot path: the entry points a
 * `crates/prover/src/backend/b200.rs` ProverBackend implementation binds through `unsafe extern "C"`
 * (the binding a maintainer adds is shown in INTEGRATION.md and rust/b200zk-sys/src/lib.rs).
 *
 * Conventions follow the in-tree C-ABI precedent
 *   /root/reference/crates/guest-program/src/crypto/zisk.rs:5-64   (declarations)
 *   /root/reference/crates/guest-program/src/crypto/zisk.rs:144-172 (status codes 0/1/2/3)
 * i.e. caller-owned buffers, plain pointers and sizes, small integer status, nothing allocated across the
 * boundary except opaque handles.  No torch / CUDA types appear in any signature: device pointers and
 * streams travel as `void*` (a `cudaStream_t` is a pointer).
 *
 * Byte formats
 *   "BE"      32-byte big-endian canonical field elements, the EIP-196/197 wire format of
 *             /root/reference/crates/common/crypto/provider.rs:201-330: G1 = x|y (64 B), (0,0) = identity;
 *             G2 = x_im|x_re|y_im|y_re (128 B); coordinates >= p are rejected like
 *             /root/reference/crates/vm/levm/src/precompiles.rs:801-820.
 *   "native"  little-endian limbs (bytes of 4 x u64 == 8 x u32 on a little-endian host).  Field elements of
 *             points and NTT data are in Montgomery form with R = 2^256 -- the in-memory form of ark-ff
 *             0.5.0 Fp256<MontBackend> that the SNARK wrap behind ProofFormat::Groth16
 *             (/root/reference/crates/prover/src/backend/sp1.rs:97-134, risc0.rs:24-29) holds its proving
 *             key in.  G1 affine = x|y (64 B), G2 affine = x.c0|x.c1|y.c0|y.c1 (128 B), (0,..,0) = identity.
 *             Scalars are canonical (non-Montgomery) 256-bit integers (ark `into_bigint()`), reduced mod r by
 *             the library, unless B200ZK_SCALARS_MONT is set.
 *
 * Semantics
 *   MSM  = ark_ec::VariableBaseMSM::msm (ark-ec 0.5.0, /root/reference/Cargo.lock:978): sum_i s_i * P_i,
 *          returned as the affine point -- the value is algorithm independent, so "bit exact" means equal
 *          (x, y).
 *   NTT  = ark_poly::Radix2EvaluationDomain (ark-poly 0.5.0, /root/reference/Cargo.lock:1140), equal to
 *          gnark-crypto's bn254 fr/fft: natural order in and out, out[k] = sum_j a[j] w^{jk},
 *          w = 5^((r-1)/2^28)^(2^(28-log_n)); inverse uses w^-1 and scales by n^-1; coset pre-multiplies a[j]
 *          by h^j (forward) / post-multiplies by h^-j (inverse), h = 5 unless given.
 *
 * Threading and streams
 *   Every entry point switches to the context's device for the duration of the call (and restores the caller's),
 *   so a context may be used from any thread and next to contexts on other devices.  A context owns ONE set of
 *   grow-only scratch buffers: drive it from one stream at a time (or order the streams with events yourself).
 *   Cached tables (NTT twiddles, coset powers) carry a build event that consumers on other streams wait on.
 *
 * There is NO CPU fallback anywhere behind this interface: without a CUDA device b200zk_init fails with
 * B200ZK_ERR_NO_DEVICE and every other call fails with B200ZK_ERR_INVALID_ARG on the NULL context.
 */
#ifndef B200ZK_H
#define B200ZK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200ZK_ABI_VERSION 2 /* v2: async MSM forms append a 4-byte is-infinity word; groth16 + ntt-root entry points */

typedef struct b200zk_ctx b200zk_ctx;

/* status codes: 0..3 are the ZisK table (zisk.rs:144-172); the rest are this library's own failures */
enum {
  B200ZK_OK = 0,
  B200ZK_OK_INFINITY = 1,       /* success, result is the identity */
  B200ZK_ERR_NOT_IN_FIELD = 2,  /* a BE coordinate is >= p */
  B200ZK_ERR_NOT_ON_CURVE = 3,  /* a BE point does not satisfy the curve equation */
  B200ZK_ERR_INVALID_ARG = 4,
  B200ZK_ERR_CUDA = 5,
  B200ZK_ERR_NO_DEVICE = 6,
  B200ZK_ERR_OOM = 7,
  B200ZK_ERR_UNSUPPORTED = 8
};

/* flags */
enum {
  B200ZK_POINTS_BE = 1u << 0,       /* points are EIP-196/197 big-endian bytes (validated); default native */
  B200ZK_SCALARS_BE = 1u << 1,      /* scalars are 32-byte big-endian; default little-endian limbs */
  B200ZK_SCALARS_MONT = 1u << 2,    /* scalars are Montgomery-form limbs (ark Fr in-memory form) */
  B200ZK_OUT_NATIVE = 1u << 3,      /* MSM result as native affine limbs instead of BE bytes */
  B200ZK_NTT_INVERSE = 1u << 4,
  B200ZK_NTT_COSET = 1u << 5,
  B200ZK_NTT_CANONICAL = 1u << 6,   /* NTT data are canonical little-endian limbs (converted on device) */
  B200ZK_NTT_BE = 1u << 7,          /* NTT data are 32-byte big-endian canonical values */
  B200ZK_G16_INPUTS_DEVICE = 1u << 8, /* b200zk_groth16_commit*: witness and evaluation buffers are DEVICE pointers (used in place) */
  B200ZK_G16_H_COEFFS = 1u << 9,    /* b200zk_groth16_commit*: a_evals already holds the quotient's coefficients (Montgomery); skip the NTTs */
  B200ZK_SCALARS_RAW = 1u << 10,    /* scalars are plain 256-bit integers < 2^255, NOT reduced mod the BN254 group order (BLS12-381 calls) */
  B200ZK_POINTS_COMPRESSED = 1u << 11 /* BLS12-381 G1 points in the 48-byte compressed ZCash / IETF format (the trusted setup's form) */
};

/* ---- lifecycle (ProverBackend::new / process-global OnceLock, cf. sp1.rs:30,93-95) ------------------- */
int b200zk_abi_version(void);
int b200zk_device_count(void);
int b200zk_init(int device, b200zk_ctx** out);
void b200zk_destroy(b200zk_ctx* ctx);
const char* b200zk_strerror(int status);
const char* b200zk_last_error(const b200zk_ctx* ctx); /* detail of the last failure on this context */
/* number of kernels this context has launched since init (bench.py's gpu_launches claim) */
uint64_t b200zk_launch_count(const b200zk_ctx* ctx);
int b200zk_synchronize(b200zk_ctx* ctx);

/* ---- host-buffer entry points: what the Rust backend calls.  Copies are part of the call. -------------- */
/* replaces ark_ec::VariableBaseMSM::msm for G1Affine / G2Affine */
int b200zk_g1_msm(b200zk_ctx* ctx, const void* points, const void* scalars, size_t n, uint32_t flags,
                  uint8_t out[64]);
int b200zk_g2_msm(b200zk_ctx* ctx, const void* points, const void* scalars, size_t n, uint32_t flags,
                  uint8_t out[128]);
/* replaces ark_poly::Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place (row a8);
 * n = 2^log_n elements of 32 bytes, transformed in place in the caller's buffer. coset_gen: 32-byte
 * canonical value in the same endianness family as the data (LE limbs, or BE with B200ZK_NTT_BE); NULL = 5 */
int b200zk_fr_ntt(b200zk_ctx* ctx, void* data, uint32_t log_n, uint32_t flags, const uint8_t* coset_gen);

/* ---- resident bases: the proving key (SRS) lives in HBM across proofs ---------------------------------- */
int b200zk_g1_bases_upload(b200zk_ctx* ctx, const void* points, size_t n, uint32_t flags, uint64_t* handle);
int b200zk_g2_bases_upload(b200zk_ctx* ctx, const void* points, size_t n, uint32_t flags, uint64_t* handle);
/* same, from points already in device memory (native format); the library keeps its own copy */
int b200zk_g1_bases_from_device(b200zk_ctx* ctx, const void* d_points, size_t n, void* stream, uint64_t* handle);
int b200zk_g2_bases_from_device(b200zk_ctx* ctx, const void* d_points, size_t n, void* stream, uint64_t* handle);
/* One-off, when the proving key is loaded: replace the resident bases by a table of their window multiples
 * 2^(c*w) * P_i, w = 0..ceil(255/c)-1 (W times the memory).  All windows then share ONE bucket set: larger
 * windows pay off (c = 22 -> 12 n additions instead of 15 n at 2^24) and the final Horner pass disappears.
 * window_bits = 0 picks c from n.  Results are unchanged (same group element). */
int b200zk_bases_precompute(b200zk_ctx* ctx, uint64_t handle, uint32_t window_bits);
int b200zk_bases_free(b200zk_ctx* ctx, uint64_t handle);
/* MSM of the first n resident bases against host scalars */
int b200zk_g1_msm_resident(b200zk_ctx* ctx, uint64_t handle, const void* scalars, size_t n, uint32_t flags,
                           uint8_t out[64]);
int b200zk_g2_msm_resident(b200zk_ctx* ctx, uint64_t handle, const void* scalars, size_t n, uint32_t flags,
                           uint8_t out[128]);
/* resident bases against scalars already in device memory (`stream` as for the *_device calls below) */
int b200zk_g1_msm_resident_device(b200zk_ctx* ctx, uint64_t handle, const void* d_scalars, size_t n,
                                  uint32_t flags, void* stream, uint8_t out[64]);
int b200zk_g2_msm_resident_device(b200zk_ctx* ctx, uint64_t handle, const void* d_scalars, size_t n,
                                  uint32_t flags, void* stream, uint8_t out[128]);

/* ---- device-pointer entry points: inputs already in HBM (native formats only) ------------------------- */
/* `stream`: a cudaStream_t passed as void*; NULL = the context's own (non-blocking) stream, so pass
 * cudaStreamLegacy ((void*)0x1) to mean the legacy default stream.  The *_device calls enqueue
 * every kernel on that stream, then copy the 64/128-byte result to `out` and wait for it. */
int b200zk_g1_msm_device(b200zk_ctx* ctx, const void* d_points, const void* d_scalars, size_t n,
                         uint32_t flags, void* stream, uint8_t out[64]);
int b200zk_g2_msm_device(b200zk_ctx* ctx, const void* d_points, const void* d_scalars, size_t n,
                         uint32_t flags, void* stream, uint8_t out[128]);
/* fully asynchronous forms: the encoded result (BE or native per flags) is written to device memory, followed by a
 * 32-bit word that is 1 when the result is the identity (what the synchronous forms return as status 1):
 * d_out68 = 64 + 4 bytes, d_out132 = 128 + 4 bytes */
int b200zk_g1_msm_device_async(b200zk_ctx* ctx, const void* d_points, const void* d_scalars, size_t n,
                               uint32_t flags, void* stream, void* d_out68);
int b200zk_g2_msm_device_async(b200zk_ctx* ctx, const void* d_points, const void* d_scalars, size_t n,
                               uint32_t flags, void* stream, void* d_out132);
int b200zk_fr_ntt_device(b200zk_ctx* ctx, void* d_data, uint32_t log_n, uint32_t flags,
                         const uint8_t* coset_gen, void* stream);
/* The primitive 2^28-th root of unity g of Fr that every domain generator derives from (w_n = g^(2^(28-log_n))) is a
 * parameter of the context.  root_le = canonical little-endian 32 bytes, NULL = back to the
 * default.  Rejected with B200ZK_ERR_INVALID_ARG unless g^(2^28) = 1 and g^(2^27) != 1.
 *   preset 0  ark-poly 0.5.0 / gnark-crypto bn254 fr:  5^((r-1)/2^28) = 0x2a3c09f0a58a7e8500e0a7eb8ef62abc402d111e41112ed49bd61b6e725b19f0
 *             (the SP1 / RISC0 Groth16 wraps, /root/reference/crates/prover/src/backend/sp1.rs:97-134, risc0.rs:24-29)   [default]
 *   preset 1  halo2curves-axiom 0.7.2 bn256::Fr:       7^((r-1)/2^28) = 0x03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c
 *             (the OpenVM halo2-KZG wrap, /root/reference/crates/prover/src/backend/openvm.rs:52-56) */
int b200zk_set_ntt_root(b200zk_ctx* ctx, const uint8_t* root_le);
int b200zk_ntt_root_preset(int preset, uint8_t root_le_out[32]);

/* ---- multi-GPU: one process per GPU, point-split MSM ------------------------------------ */
/* Each rank reduces its shard to ONE partial sum in extended-Jacobian XYZZ form (G1: 128 B, G2: 256 B,
 * native Montgomery limbs) left in device memory; the host side all-gathers the partials over NCCL and
 * every rank folds them with b200zk_g{1,2}_fold_partials_device -- NCCL has no elliptic-curve reduce op, so
 * this pair is the "allreduce of partial sums". */
int b200zk_g1_msm_partial_device(b200zk_ctx* ctx, const void* d_points, const void* d_scalars, size_t n,
                                 uint32_t flags, void* stream, void* d_partial128);
int b200zk_g2_msm_partial_device(b200zk_ctx* ctx, const void* d_points, const void* d_scalars, size_t n,
                                 uint32_t flags, void* stream, void* d_partial256);
/* same, over resident (possibly precomputed) bases */
int b200zk_g1_msm_partial_resident_device(b200zk_ctx* ctx, uint64_t handle, const void* d_scalars, size_t n,
                                          uint32_t flags, void* stream, void* d_partial128);
int b200zk_g2_msm_partial_resident_device(b200zk_ctx* ctx, uint64_t handle, const void* d_scalars, size_t n,
                                          uint32_t flags, void* stream, void* d_partial256);
/* same, with the shard's scalars in (pinned) HOST memory: the upload is pipelined with the accumulation */
int b200zk_g1_msm_partial_resident(b200zk_ctx* ctx, uint64_t handle, const void* scalars, size_t n, uint32_t flags,
                                   void* stream, void* d_partial128);
int b200zk_g2_msm_partial_resident(b200zk_ctx* ctx, uint64_t handle, const void* scalars, size_t n, uint32_t flags,
                                   void* stream, void* d_partial256);
int b200zk_g1_fold_partials_device(b200zk_ctx* ctx, const void* d_partials, size_t count, uint32_t flags,
                                   void* stream, uint8_t out[64]);
int b200zk_g2_fold_partials_device(b200zk_ctx* ctx, const void* d_partials, size_t count, uint32_t flags,
                                   void* stream, uint8_t out[128]);

/* ---- device utilities (format conversion, synthetic workloads; used by tests and bench.py) ------------- */
/* canonical LE limbs <-> Montgomery limbs, in place in device memory; which = 0 for Fq, 1 for Fr */
int b200zk_field_to_mont_device(b200zk_ctx* ctx, void* d_data, size_t n, int which, void* stream);
int b200zk_field_from_mont_device(b200zk_ctx* ctx, void* d_data, size_t n, int which, void* stream);
/* out[i] = a[i] * b[i] (Montgomery product), the field core exposed for parity tests and microbenchmarks;
 * `repeat` > 1 chains out = out * b that many times (throughput measurement).
 * which: 0 = Fq, 1 = Fr; +2 = the dedicated squaring instead (out = a^2, chained: a^(2^repeat); d_b is read but unused) */
int b200zk_field_mul_device(b200zk_ctx* ctx, const void* d_a, const void* d_b, void* d_out, size_t n,
                            int which, uint32_t repeat, void* stream);
/* out[i] = (a[i]*b[i] - c[i]) * zinv over Fr (Montgomery data; zinv canonical LE): the pointwise step of the
 * Groth16 quotient H = (A*B - C)/Z_H evaluated on a coset, where Z_H is the constant h^n - 1 */
int b200zk_fr_quotient_device(b200zk_ctx* ctx, const void* d_a, const void* d_b, const void* d_c, void* d_out,
                              size_t n, const uint8_t zinv[32], void* stream);
/* counter-based splitmix64 scalars: element i = reduce_mod_r(4 outputs of state seed + 4*(start+i)*golden)
 * (SURVEY.md 8d); canonical limbs, or Montgomery with B200ZK_SCALARS_MONT */
int b200zk_fr_random_device(b200zk_ctx* ctx, void* d_out, size_t n, uint64_t seed, uint64_t start,
                            uint32_t flags, void* stream);
/* synthetic base chain P_i = (k + i*d) * G for i in [start, start+n) (native affine);
 * k, d: canonical LE 32-byte scalars */
int b200zk_g1_chain_device(b200zk_ctx* ctx, void* d_out, size_t start, size_t n, const uint8_t k[32],
                           const uint8_t d[32], void* stream);
int b200zk_g2_chain_device(b200zk_ctx* ctx, void* d_out, size_t start, size_t n, const uint8_t k[32],
                           const uint8_t d[32], void* stream);
/* on-curve check of n native affine points; *bad_index = first offending index or n */
int b200zk_g1_check_device(b200zk_ctx* ctx, const void* d_points, size_t n, void* stream, size_t* bad_index);
int b200zk_g2_check_device(b200zk_ctx* ctx, const void* d_points, size_t n, void* stream, size_t* bad_index);

/* tuning knobs (0 = automatic): window bits for the next MSM calls on this context */
int b200zk_set_msm_window(b200zk_ctx* ctx, uint32_t c);
/* chunks of the pipelined MSM schedule (sort of chunk k+1 overlaps the accumulation of chunk k); 0 = automatic
 * (4 from 2^22 points), 1 = one shot */
int b200zk_set_msm_chunks(b200zk_ctx* ctx, uint32_t chunks);
/* rounds of batched-affine pair summing run before the bucket accumulation (0..4; negative = automatic) */
int b200zk_set_msm_pair_rounds(b200zk_ctx* ctx, int rounds);
/* per-phase device time of the last *_device MSM call, in milliseconds:
 * [0] digit histogram, [1] scan, [2] scatter, [3] bucket accumulation, [4] bucket reduction, [5] final */
int b200zk_last_msm_phase_ms(b200zk_ctx* ctx, float out_ms[6]);
int b200zk_set_profiling(b200zk_ctx* ctx, int enabled);

/* Several resident-base MSMs over ONE scalar vector -- Groth16's [A]1, [B]1, [B]2 and [L]1 all multiply the witness
 * (the prove step behind crates/prover/src/backend/sp1.rs:97-134 / risc0.rs:71-82).  The scalar-dependent half of
 * the MSM (digit recoding, bucket histogram, scan, scatter: ~15 % of a 2^24 MSM) runs once and is shared; G1 and G2
 * handles may be mixed.  All handles must have been precomputed with the same window (or none of them) and, when
 * precomputed, hold the same number of points.  out: `count` slots of 128 bytes (a G1 result uses the first 64);
 * status[i] = 0, or 1 when result i is the identity.  Device scalars, like b200zk_g1_msm_resident_device. */
int b200zk_msm_multi_resident_device(b200zk_ctx* ctx, const uint64_t* handles, size_t count, const void* d_scalars, size_t n,
                                     uint32_t flags, void* stream, uint8_t* out /* count*128 */, int* status /* count */);

/* ---- the Groth16 prove arithmetic as one call --------------------------------------
 * What the SNARK wrap behind ProofFormat::Groth16 computes after witness generation
 * (crates/prover/src/backend/sp1.rs:97-134 -> gnark groth16.Prove; risc0.rs:24-29,71-82 -> risc0-groth16), over a
 * proving key that lives in HBM (b200zk_g{1,2}_bases_upload + b200zk_bases_precompute, once per process -- the
 * OnceLock setup of sp1.rs:30,93-95):
 *     quotient  3 iNTT + 3 coset NTT + (a*b - c)/Z_H + 1 coset iNTT    (coset generator 5, the context's NTT root)
 *     commit    [A]1, [B]1, [B]2 over the witness, [L]1 over its private part, [H]1 over the quotient's coefficients
 *     assemble  proof = A (64) | B2 (128, x_im|x_re|y_im|y_re) | C (64), C = [L]1 + [H]1; no blinding (r = s = 0)
 * Columns: 0 = A_g1, 1 = B_g1 (handle 0 = absent), 2 = B_g2, 3 = L_g1, 4 = H_g1.  count[k] = points of column k this
 * context multiplies; offset[k] = index of the scalar its first point multiplies -- into the witness for columns 0..3
 * (L_g1 of a whole key: offset = number of public inputs incl. the leading 1), into the quotient's coefficients for
 * column 4 (count <= 2^log_n - 1 for a whole key).  A single GPU holds whole columns (offsets 0,0,0,n_public,0); a
 * point-split rank holds a slice of each and passes the slice's start.  Columns that multiply the same scalar slice
 * under the same window plan share ONE digit sort. */
typedef struct b200zk_groth16_pk {
  uint32_t log_n;      /* quotient domain 2^log_n */
  uint32_t reserved;   /* 0 */
  uint64_t handle[5];
  uint64_t count[5];
  uint64_t offset[5];
} b200zk_groth16_pk;
/* witness: scalars (canonical LE unless B200ZK_SCALARS_BE / _MONT), at least max(offset+count) over columns 0..3;
 * a_evals, b_evals, c_evals: (A z), (B z), (C z) on the domain, 2^log_n Montgomery LE elements each.  HOST buffers by
 * default (copied in, left untouched); with B200ZK_G16_INPUTS_DEVICE device pointers, used in place (a_evals is
 * overwritten with the quotient's coefficients, b/c with their coset evaluations).  proof: 256 bytes out;
 * b_g1 (may be NULL): [B]1 as 64 bytes.  One stream, one synchronisation at the end. */
int b200zk_groth16_commit(b200zk_ctx* ctx, const b200zk_groth16_pk* pk, const void* witness, void* a_evals, void* b_evals,
                          void* c_evals, uint32_t flags, void* stream, uint8_t proof[256], uint8_t b_g1[64]);
/* multi-GPU halves: every rank leaves its five partial sums (768 B: A | B1 | B2 | L | H, XYZZ native limbs) in device
 * memory without synchronising; the host side all-gathers the blocks ONCE and every rank folds `count` of them */
int b200zk_groth16_commit_partial(b200zk_ctx* ctx, const b200zk_groth16_pk* pk, const void* witness, void* a_evals,
                                  void* b_evals, void* c_evals, uint32_t flags, void* stream, void* d_partials768);
int b200zk_groth16_fold(b200zk_ctx* ctx, const void* d_partials, size_t count, void* stream, uint8_t proof[256],
                        uint8_t b_g1[64]);

/* ---- BLS12-381 G1 / EIP-4844 blob commitments -------------------------------------
 * The L2 committer's "commit" step: /root/reference/crates/common/crypto/kzg.rs:259-272 (blob_to_kzg_commitment_and_proof ->
 * c_kzg blob_to_kzg_commitment), /root/reference/crates/common/types/blobs_bundle.rs:90-118, crates/l2/sequencer/
 * l1_committer.rs:1488-1521.  The same Pippenger kernels, instantiated over the 381-bit base field (12 x 32-bit limbs).
 * Points: 48-byte compressed (B200ZK_POINTS_COMPRESSED; bit 7 of byte 0 = compressed, bit 6 = infinity, bit 5 = the larger
 * y) -- the form the trusted setup ships in -- or 96-byte uncompressed big-endian x | y.  Statuses: 2 = a coordinate or a
 * scalar out of its field, 3 = not a curve point / malformed flag bits.  The subgroup check is the trusted setup's
 * business (c-kzg validates it when loading), not repeated here.  The handle works with b200zk_bases_precompute /
 * b200zk_bases_free like any other.  Scalars: 32-byte integers < the BLS12-381 group order r, little-endian limbs or
 * big-endian with B200ZK_SCALARS_BE (then checked against r); result: 48 bytes compressed. */
int b200zk_bls12_381_g1_bases_upload(b200zk_ctx* ctx, const void* points, size_t n, uint32_t flags, uint64_t* handle);
int b200zk_bls12_381_g1_msm_resident(b200zk_ctx* ctx, uint64_t handle, const void* scalars, size_t n, uint32_t flags,
                                     uint8_t out[48]);
/* blob_to_kzg_commitment for n_blobs blobs: blob = 4096 x 32-byte big-endian field elements (each < r, else status 2);
 * setup_handle = the 4096 Lagrange-form G1 points of the trusted setup (g1_lagrange_brp order, as c-kzg holds them);
 * commitments: n_blobs x 48 bytes */
int b200zk_kzg_blob_to_commitment(b200zk_ctx* ctx, uint64_t setup_handle, const uint8_t* blobs, size_t n_blobs,
                                  uint8_t* commitments);

/* ---- batched EIP-196 / EIP-197 precompile arithmetic ------------------------------
 * The three BN254 calls of the reference's `Crypto` trait, `count` independent items per call, HOST buffers:
 *   bn254_g1_add         crates/common/crypto/provider.rs:201-234   (levm ecadd,     crates/vm/levm/src/precompiles.rs:692-716)
 *   bn254_g1_mul         crates/common/crypto/provider.rs:239-272   (levm ecmul,     precompiles.rs:719-745)
 *   bn254_pairing_check  crates/common/crypto/provider.rs:277-330   (levm ecpairing, precompiles.rs:821-860)
 * Encodings: G1 = 64 B big-endian x|y, (0,0) = identity; G2 = 128 B x_im|x_re|y_im|y_re; scalars 32 B big-endian
 * (any 256-bit value: the group has prime order).  The function's return value reports infrastructure errors only;
 * input errors are PER ITEM in status[i], with the ZisK-style table (crates/guest-program/src/crypto/zisk.rs:144-172):
 *   0 ok, 1 ok and the result is the identity, 2 a coordinate >= p (levm: CoordinateExceedsFieldModulus,
 *   checked for every point of the item before any curve check), 3 a point is not on the curve or (G2) not in the
 *   order-r subgroup.  Outputs of failed items are zero.
 * The reference's own vectors for this surface (14 ecpairing cases + the out-of-range case,
 * test/tests/levm/precompile_tests.rs:17-151; 7*(1,2) of test/tests/l2/integration_tests.rs:572) run through these
 * entry points in tests/test_gpu_parity.py. */
int b200zk_bn254_g1_add_batch(b200zk_ctx* ctx, const uint8_t* a /* count*64 */, const uint8_t* b /* count*64 */, size_t count,
                              uint8_t* out /* count*64 */, uint8_t* status /* count */);
int b200zk_bn254_g1_mul_batch(b200zk_ctx* ctx, const uint8_t* points /* count*64 */, const uint8_t* scalars /* count*32 */, size_t count,
                              uint8_t* out /* count*64 */, uint8_t* status /* count */);
/* check i covers pairs [pair_offsets[i], pair_offsets[i+1]) of `pairs` (192 B each: G1 | G2); result[i] = 1 when
 * the product of its pairings is one (an empty check is 1), 0 otherwise or when status[i] != 0 */
int b200zk_bn254_pairing_check_batch(b200zk_ctx* ctx, const uint8_t* pairs, const uint32_t* pair_offsets /* count+1 */, size_t count,
                                     uint8_t* result /* count */, uint8_t* status /* count */);

#ifdef __cplusplus
}
#endif
#endif /* B200ZK_H */
