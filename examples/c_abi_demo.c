/* The drop-in boundary from plain C: no Python, no torch -- only include/b200zk.h and libb200zk.so.
 * This is what a cgo / Rust-FFI / JNI binding does.  Exit code 0 = every check passed.
 *
 *   gcc -O2 -std=c11 -I include examples/c_abi_demo.c -L ethrex_b200 -lb200zk -Wl,-rpath,'$ORIGIN/../../ethrex_b200' -o examples/build/c_abi_demo
 *
 * Checks (all against constants the reference holds):
 *   1. b200zk_bn254_g1_mul_batch: 7 * (1,2) = the point of /root/reference/test/tests/l2/integration_tests.rs:572
 *   2. b200zk_g1_msm (big-endian points and scalars): 3*G + 4*G = 7*G; 5*G + (r-5)*G = identity (status 1)
 *   3. b200zk_bn254_pairing_check_batch: e(G1, G2) * e(-G1, G2) = 1 (EIP-197 generator), and e(G1,G2)^2 != 1
 *   4. b200zk_fr_ntt: forward then inverse of 2^10 canonical big-endian values returns the input; NTT(delta_0) = 1...1
 *   5. error convention: a coordinate >= p gives status 2 (CoordinateExceedsFieldModulus), (1,3) status 3
 *   6. (with a fixture path as argv[1]) a REAL Groth16 instance through C only: the proving key is uploaded once
 *      (b200zk_g{1,2}_bases_upload + b200zk_bases_precompute), then ONE call of b200zk_groth16_commit -- 7 NTTs, the
 *      quotient, 4 G1 MSMs + 1 G2 MSM, C = L + H -- must return, byte for byte, the proof computed in the exponent
 *      (tests/groth16_toy.py writes the fixture; layout in write_c_fixture()). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "b200zk.h"

static int from_hex(const char* hex, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    unsigned v;
    if (sscanf(hex + 2 * i, "%2x", &v) != 1) return -1;
    out[i] = (uint8_t)v;
  }
  return 0;
}
#define CHECK(cond, what)                                              \
  do {                                                                 \
    if (!(cond)) { fprintf(stderr, "FAIL: %s (%s)\n", what, ctx ? b200zk_last_error(ctx) : ""); return 1; } \
    printf("ok   %s\n", what);                                         \
  } while (0)

static uint8_t* slurp(const char* path, size_t* len) {
  FILE* f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  uint8_t* buf = (uint8_t*)malloc((size_t)n);
  if (buf && fread(buf, 1, (size_t)n, f) != (size_t)n) { free(buf); buf = NULL; }
  fclose(f);
  *len = (size_t)n;
  return buf;
}

int main(int argc, char** argv) {
  b200zk_ctx* ctx = NULL;
  int rc = b200zk_init(0, &ctx);
  if (rc != B200ZK_OK) { fprintf(stderr, "b200zk_init: status %d (%s)\n", rc, b200zk_strerror(rc)); return 2; }

  uint8_t g[64] = {0}, seven_g[64], p_be[32], r_be[32];
  g[31] = 1; g[63] = 2;
  from_hex("17072b2ed3bb8d759a5325f477629386cb6fc6ecb801bd76983a6b86abffe078168ada6cd130dd52017bb54bfa19377aadfe3bf05d18f41b77809f7f60d4af9e", seven_g, 64);
  from_hex("30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47", p_be, 32);
  from_hex("30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001", r_be, 32);

  /* 1. ecMul */
  uint8_t k7[32] = {0}, out[64], st[4];
  k7[31] = 7;
  rc = b200zk_bn254_g1_mul_batch(ctx, g, k7, 1, out, st);
  CHECK(rc == B200ZK_OK && st[0] == 0 && !memcmp(out, seven_g, 64), "ecMul 7*(1,2) matches the reference's KAT");

  /* 2. MSM through the host entry point, EIP-196 bytes in and out */
  uint8_t pts[128], sc[64] = {0};
  memcpy(pts, g, 64); memcpy(pts + 64, g, 64);
  sc[31] = 3; sc[63] = 4;
  rc = b200zk_g1_msm(ctx, pts, sc, 2, B200ZK_POINTS_BE | B200ZK_SCALARS_BE, out);
  CHECK(rc == B200ZK_OK && !memcmp(out, seven_g, 64), "MSM 3*G + 4*G = 7*G");
  memset(sc, 0, 64); sc[31] = 5;
  { /* second scalar = r - 5, big-endian, with the borrow */
    int i = 63, borrow = 5; memcpy(sc + 32, r_be, 32);
    while (borrow && i >= 32) { int v = sc[i] - borrow; borrow = v < 0; sc[i] = (uint8_t)(v & 0xff); --i; }
  }
  rc = b200zk_g1_msm(ctx, pts, sc, 2, B200ZK_POINTS_BE | B200ZK_SCALARS_BE, out);
  uint8_t zero64[64] = {0};
  CHECK(rc == B200ZK_OK_INFINITY && !memcmp(out, zero64, 64), "MSM 5*G + (r-5)*G = identity, status 1");

  /* 3. pairing check with the EIP-197 G2 generator (x_im | x_re | y_im | y_re) */
  uint8_t g2[128], neg_g[64], pairs[2 * 192], res[2], pst[2];
  from_hex("198e9393920d483a7260bfb731fb5d25f1aa493335a9e71297e485b7aef312c2"
           "1800deef121f1e76426a00665e5c4479674322d4f75edadd46debd5cd992f6ed"
           "090689d0585ff075ec9e99ad690c3395bc4b313370b38ef355acdadcd122975b"
           "12c85ea5db8c6deb4aab71808dcb408fe3d1e7690c43d37b4ce6cc0166fa7daa", g2, 128);
  memcpy(neg_g, g, 64);
  memcpy(neg_g + 32, p_be, 32); neg_g[63] -= 2;  /* y = p - 2 (low byte of p is 0x47) */
  memcpy(pairs, g, 64); memcpy(pairs + 64, g2, 128);
  memcpy(pairs + 192, neg_g, 64); memcpy(pairs + 256, g2, 128);
  uint32_t offs[2] = {0, 2};
  rc = b200zk_bn254_pairing_check_batch(ctx, pairs, offs, 1, res, pst);
  CHECK(rc == B200ZK_OK && pst[0] == 0 && res[0] == 1, "e(G1,G2) * e(-G1,G2) == 1");
  memcpy(pairs + 192, g, 64);
  rc = b200zk_bn254_pairing_check_batch(ctx, pairs, offs, 1, res, pst);
  CHECK(rc == B200ZK_OK && pst[0] == 0 && res[0] == 0, "e(G1,G2)^2 != 1");

  /* 4. NTT round trip on canonical big-endian values */
  const uint32_t log_n = 10, n = 1u << log_n;
  uint8_t* a = (uint8_t*)calloc(n, 32);
  uint8_t* b = (uint8_t*)malloc((size_t)n * 32);
  for (uint32_t i = 0; i < n; ++i) { a[32 * i + 31] = (uint8_t)(i * 7 + 1); a[32 * i + 30] = (uint8_t)(i >> 3); a[32 * i + 5] = (uint8_t)(i * 13); }
  memcpy(b, a, (size_t)n * 32);
  rc = b200zk_fr_ntt(ctx, b, log_n, B200ZK_NTT_BE, NULL);
  int changed = memcmp(a, b, (size_t)n * 32) != 0;
  int rc2 = b200zk_fr_ntt(ctx, b, log_n, B200ZK_NTT_BE | B200ZK_NTT_INVERSE, NULL);
  CHECK(rc == B200ZK_OK && rc2 == B200ZK_OK && changed && !memcmp(a, b, (size_t)n * 32), "iNTT(NTT(a)) == a on 2^10 big-endian values");
  memset(b, 0, (size_t)n * 32); b[31] = 1;  /* delta_0 */
  rc = b200zk_fr_ntt(ctx, b, log_n, B200ZK_NTT_BE, NULL);
  int all_one = rc == B200ZK_OK;
  for (uint32_t i = 0; i < n && all_one; ++i) {
    for (int j = 0; j < 31; ++j) all_one = all_one && b[32 * i + j] == 0;
    all_one = all_one && b[32 * i + 31] == 1;
  }
  CHECK(all_one, "NTT(delta_0) = (1, ..., 1)");

  /* 5. error convention */
  uint8_t bad[128];
  memcpy(bad, p_be, 32); memset(bad + 32, 0, 32); bad[63] = 2;          /* x = p */
  memset(bad + 64, 0, 64); bad[64 + 31] = 1; bad[64 + 63] = 3;          /* (1, 3) */
  uint8_t ks[64] = {0}, outs[128], sts[2];
  ks[31] = 1; ks[63] = 1;
  rc = b200zk_bn254_g1_mul_batch(ctx, bad, ks, 2, outs, sts);
  CHECK(rc == B200ZK_OK && sts[0] == B200ZK_ERR_NOT_IN_FIELD && sts[1] == B200ZK_ERR_NOT_ON_CURVE, "status 2 for x >= p, status 3 for a point off the curve");
  rc = b200zk_g1_msm(ctx, bad, ks, 1, B200ZK_POINTS_BE | B200ZK_SCALARS_BE, out);
  CHECK(rc == B200ZK_ERR_NOT_IN_FIELD, "host MSM rejects x >= p with status 2");

  /* 6. Groth16 prove arithmetic in one call, against a proof computed in the exponent */
  if (argc > 1) {
    size_t len = 0;
    uint8_t* fx = slurp(argv[1], &len);
    CHECK(fx && len > 16 && !memcmp(fx, "G16F", 4), "fixture file loads");
    uint32_t hdr[3];
    memcpy(hdr, fx + 4, 12);
    const uint32_t k = hdr[0], m = hdr[1], npub = hdr[2];
    const size_t dn = (size_t)1 << k;
    const uint8_t* p = fx + 16;
    const uint8_t *a_g1 = p; p += (size_t)m * 64;
    const uint8_t *b_g1 = p; p += (size_t)m * 64;
    const uint8_t *b_g2 = p; p += (size_t)m * 128;
    const uint8_t *l_g1 = p; p += (size_t)(m - npub) * 64;
    const uint8_t *h_g1 = p; p += (dn - 1) * 64;
    const uint8_t *wit = p; p += (size_t)m * 32;
    uint8_t *ea = (uint8_t*)p; p += dn * 32;
    uint8_t *eb = (uint8_t*)p; p += dn * 32;
    uint8_t *ec = (uint8_t*)p; p += dn * 32;
    const uint8_t* want = p; p += 256;
    CHECK((size_t)(p - fx) == len, "fixture layout matches its header");
    b200zk_groth16_pk pk;
    memset(&pk, 0, sizeof pk);
    pk.log_n = k;
    int up = b200zk_g1_bases_upload(ctx, a_g1, m, B200ZK_POINTS_BE, &pk.handle[0]);
    up |= b200zk_g1_bases_upload(ctx, b_g1, m, B200ZK_POINTS_BE, &pk.handle[1]);
    up |= b200zk_g2_bases_upload(ctx, b_g2, m, B200ZK_POINTS_BE, &pk.handle[2]);
    up |= b200zk_g1_bases_upload(ctx, l_g1, m - npub, B200ZK_POINTS_BE, &pk.handle[3]);
    up |= b200zk_g1_bases_upload(ctx, h_g1, dn - 1, B200ZK_POINTS_BE, &pk.handle[4]);
    for (int c = 0; c < 5 && !up; ++c) up |= b200zk_bases_precompute(ctx, pk.handle[c], 0);
    CHECK(up == B200ZK_OK, "proving key uploaded and expanded into window tables (once per process)");
    pk.count[0] = pk.count[1] = pk.count[2] = m; pk.count[3] = m - npub; pk.count[4] = dn - 1;
    pk.offset[3] = npub;
    uint8_t proof[256], b1[64];
    rc = b200zk_groth16_commit(ctx, &pk, wit, ea, eb, ec, 0, NULL, proof, b1);
    CHECK(rc == B200ZK_OK && !memcmp(proof, want, 256), "b200zk_groth16_commit == the proof computed in the exponent (256 bytes)");
    rc = b200zk_groth16_commit(ctx, &pk, wit, ea, eb, ec, 0, NULL, proof, NULL);
    CHECK(rc == B200ZK_OK && !memcmp(proof, want, 256), "second proof over the resident key: same bytes");
    pk.count[4] = dn;  /* one more H point than the key holds */
    rc = b200zk_groth16_commit(ctx, &pk, wit, ea, eb, ec, 0, NULL, proof, NULL);
    CHECK(rc == B200ZK_ERR_INVALID_ARG, "a column count beyond the resident bases is rejected, not truncated");
    for (int c = 0; c < 5; ++c) b200zk_bases_free(ctx, pk.handle[c]);
    free(fx);
  }

  free(a); free(b);
  b200zk_destroy(ctx);
  printf("c_abi_demo: all checks passed\n");
  return 0;
}
