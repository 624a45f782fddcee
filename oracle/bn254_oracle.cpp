// oracle/bn254_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement (C++17, g++ -O3 -fopenmp, 4x64-bit Montgomery limbs on unsigned __int128)
// of the BN254 G1/G2 multi-scalar multiplication and Fr number-theoretic transform that
// ethrex's L2 prover reaches through ProofFormat::Groth16
// (/root/reference/crates/l2/sequencer/proof_coordinator.rs:252-256,
//  /root/reference/crates/prover/src/backend/sp1.rs:97-134, risc0.rs:24-29,71-82).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.  The product (ethrex_b200/, libb200zk.so) shares no code with it.
//
// PARITY STATUS: "parity unpinned" for MSM / NTT outputs.  The reference tree holds no MSM or
// NTT code, test or golden vector; the arithmetic lives in crates that
// are pinned in /root/reference/Cargo.lock but not vendored: ark-ec 0.5.0 (Cargo.lock:978),
// ark-poly 0.5.0 (Cargo.lock:1140), ark-ff 0.5.0, ark-bn254 0.5.0.  This file restates their
// *published* algorithms:
//   * ark-ec 0.5.0 src/scalar_mul/variable_base/mod.rs: msm_bigint_wnaf -- window
//     c = 3 if n < 32 else floor(log2(n)*69/100)+2, signed radix-2^c digits (make_digits),
//     2^(c-1) buckets per window, running-sum bucket reduction, Horner over windows;
//   * ark-poly 0.5.0 src/domain/radix2: out[k] = sum_j a[j] w^{jk}, natural order in/out,
//     w_n = g^(2^(28-log n)), g = 5^((r-1)/2^28); ifft scales by n^-1; coset_fft pre-multiplies
//     a[j] by h^j, coset_ifft post-multiplies by h^-j.
// It is pinned (tests/test_oracle.py) against: the independent pure-Python big-int
// implementation oracle/pyref.py; the reference's own KATs for the primitives -- 7*(1,2)
// (/root/reference/test/tests/l2/integration_tests.rs:572), the on-curve G1/G2 points of
// /root/reference/test/tests/levm/precompile_tests.rs:17-24, ALT_BN128_PRIME
// (/root/reference/crates/vm/levm/src/precompiles.rs:746-751); and the byte conventions of
// /root/reference/crates/common/crypto/provider.rs:201-330 (32-byte big-endian canonical
// coordinates, (0,0) = identity, G2 = x_im|x_re|y_im|y_re); and, through pyref's pairing, the 14 ecpairing
// KATs of /root/reference/test/tests/levm/precompile_tests.rs:17-140 (every point decodes here; this file's
// G1/G2 Pippenger results satisfy bilinearity under the pairing that replays those KATs).  Status codes follow the in-tree
// C-ABI precedent /root/reference/crates/guest-program/src/crypto/zisk.rs:144-172
// (0 ok, 1 ok-infinity, 2 not in field, 3 not on curve).
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t u64;
typedef unsigned __int128 u128;

// ------------------------------------------------------------------------------------------
// 256-bit helpers
static inline bool geq4(const u64* a, const u64* b) {
  for (int i = 3; i >= 0; --i) { if (a[i] != b[i]) return a[i] > b[i]; }
  return true;
}
static inline u64 add4(u64* r, const u64* a, const u64* b) {
  u128 c = 0;
  for (int i = 0; i < 4; ++i) { c += (u128)a[i] + b[i]; r[i] = (u64)c; c >>= 64; }
  return (u64)c;
}
static inline u64 sub4(u64* r, const u64* a, const u64* b) {
  u64 br = 0;
  for (int i = 0; i < 4; ++i) {
    u128 d = (u128)a[i] - b[i] - br; r[i] = (u64)d; br = (u64)(d >> 64) & 1;
  }
  return br;
}

// ------------------------------------------------------------------------------------------
// prime fields in Montgomery form (R = 2^256)
struct FqP {
  static constexpr u64 MOD[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
  static constexpr u64 INV = 0x87d20782e4866389ULL;  // -p^-1 mod 2^64
  static constexpr u64 R1[4] = {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL};
  static constexpr u64 R2[4] = {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL};
};
struct FrP {
  static constexpr u64 MOD[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
  static constexpr u64 INV = 0xc2e1f593efffffffULL;
  static constexpr u64 R1[4] = {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL};
  static constexpr u64 R2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL};
};

template <class Pm>
struct Fp {
  u64 v[4];
  static Fp zero() { Fp r; memset(r.v, 0, 32); return r; }
  static Fp one() { Fp r; memcpy(r.v, Pm::R1, 32); return r; }
  bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
  bool operator==(const Fp& o) const { return memcmp(v, o.v, 32) == 0; }
  bool operator!=(const Fp& o) const { return !(*this == o); }
  Fp operator+(const Fp& o) const {
    Fp r; u64 c = add4(r.v, v, o.v);
    if (c || geq4(r.v, Pm::MOD)) sub4(r.v, r.v, Pm::MOD);
    return r;
  }
  Fp operator-(const Fp& o) const {
    Fp r; if (sub4(r.v, v, o.v)) add4(r.v, r.v, Pm::MOD);
    return r;
  }
  Fp neg() const { return is_zero() ? *this : (Fp{{Pm::MOD[0], Pm::MOD[1], Pm::MOD[2], Pm::MOD[3]}} - *this); }
  Fp dbl() const { return *this + *this; }
  // CIOS Montgomery product
  Fp operator*(const Fp& o) const {
    u64 t[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
      u128 c = 0;
      for (int j = 0; j < 4; ++j) { c += (u128)v[j] * o.v[i] + t[j]; t[j] = (u64)c; c >>= 64; }
      c += t[4]; t[4] = (u64)c; u64 t5 = (u64)(c >> 64);
      u64 m = t[0] * Pm::INV;
      c = (u128)m * Pm::MOD[0] + t[0]; c >>= 64;
      for (int j = 1; j < 4; ++j) { c += (u128)m * Pm::MOD[j] + t[j]; t[j - 1] = (u64)c; c >>= 64; }
      c += t[4]; t[3] = (u64)c; t[4] = t5 + (u64)(c >> 64);
    }
    Fp r; memcpy(r.v, t, 32);
    if (t[4] || geq4(r.v, Pm::MOD)) sub4(r.v, r.v, Pm::MOD);
    return r;
  }
  Fp sqr() const { return *this * *this; }
  static Fp from_canonical(const u64* c) {  // c < modulus assumed (callers reduce first)
    Fp a; memcpy(a.v, c, 32); Fp r2; memcpy(r2.v, Pm::R2, 32); return a * r2;
  }
  void to_canonical(u64* out) const {
    Fp o; memset(o.v, 0, 32); o.v[0] = 1; Fp r = *this * o; memcpy(out, r.v, 32);
  }
  Fp pow(const u64* e) const {  // 256-bit exponent, little-endian limbs
    Fp acc = one();
    for (int i = 255; i >= 0; --i) {
      acc = acc.sqr();
      if ((e[i >> 6] >> (i & 63)) & 1) acc = acc * *this;
    }
    return acc;
  }
  Fp inv() const {  // Fermat; inv(0) = 0
    u64 e[4]; u64 two[4] = {2, 0, 0, 0}; sub4(e, Pm::MOD, two); return pow(e);
  }
};
typedef Fp<FqP> Fq;
typedef Fp<FrP> Fr;

// Fq2 = Fq[u]/(u^2+1)
struct Fq2 {
  Fq c0, c1;
  static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
  static Fq2 one() { return {Fq::one(), Fq::zero()}; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
  bool operator!=(const Fq2& o) const { return !(*this == o); }
  Fq2 operator+(const Fq2& o) const { return {c0 + o.c0, c1 + o.c1}; }
  Fq2 operator-(const Fq2& o) const { return {c0 - o.c0, c1 - o.c1}; }
  Fq2 neg() const { return {c0.neg(), c1.neg()}; }
  Fq2 dbl() const { return {c0.dbl(), c1.dbl()}; }
  Fq2 operator*(const Fq2& o) const {
    Fq a = c0 * o.c0, b = c1 * o.c1;
    Fq s = (c0 + c1) * (o.c0 + o.c1);
    return {a - b, s - a - b};
  }
  Fq2 sqr() const { return *this * *this; }
  Fq2 inv() const {
    Fq d = (c0.sqr() + c1.sqr()).inv();
    return {c0 * d, (c1 * d).neg()};
  }
};

// ------------------------------------------------------------------------------------------
// short-Weierstrass y^2 = x^3 + b, a = 0.  Affine (0,0) = identity.  Jacobian Z = 0 = identity.
template <class F> struct Aff { F x, y; bool inf() const { return x.is_zero() && y.is_zero(); } };
template <class F>
struct Jac {
  F X, Y, Z;
  static Jac identity() { return {F::one(), F::one(), F::zero()}; }
  bool inf() const { return Z.is_zero(); }
  Jac dbl() const {  // dbl-2009-l
    if (inf()) return *this;
    F A = X.sqr(), B = Y.sqr(), C = B.sqr();
    F D = ((X + B).sqr() - A - C).dbl();
    F E = A.dbl() + A, Fv = E.sqr();
    Jac r;
    r.X = Fv - D.dbl();
    r.Z = (Y * Z).dbl();
    r.Y = E * (D - r.X) - C.dbl().dbl().dbl();
    return r;
  }
  Jac add_mixed(const Aff<F>& q) const {  // madd-2007-bl with the exceptional cases
    if (q.inf()) return *this;
    if (inf()) return {q.x, q.y, F::one()};
    F Z1Z1 = Z.sqr(), U2 = q.x * Z1Z1, S2 = q.y * Z * Z1Z1;
    if (U2 == X) {
      if (S2 == Y) return dbl();
      return identity();
    }
    F H = U2 - X, HH = H.sqr(), I = HH.dbl().dbl(), J = H * I;
    F r = (S2 - Y).dbl(), V = X * I;
    Jac o;
    o.X = r.sqr() - J - V.dbl();
    o.Y = r * (V - o.X) - (Y * J).dbl();
    o.Z = (Z + H).sqr() - Z1Z1 - HH;
    return o;
  }
  Jac add(const Jac& q) const {  // add-2007-bl with the exceptional cases
    if (q.inf()) return *this;
    if (inf()) return q;
    F Z1Z1 = Z.sqr(), Z2Z2 = q.Z.sqr();
    F U1 = X * Z2Z2, U2 = q.X * Z1Z1, S1 = Y * q.Z * Z2Z2, S2 = q.Y * Z * Z1Z1;
    if (U1 == U2) {
      if (S1 == S2) return dbl();
      return identity();
    }
    F H = U2 - U1, I = H.dbl().sqr(), J = H * I, r = (S2 - S1).dbl(), V = U1 * I;
    Jac o;
    o.X = r.sqr() - J - V.dbl();
    o.Y = r * (V - o.X) - (S1 * J).dbl();
    o.Z = ((Z + q.Z).sqr() - Z1Z1 - Z2Z2) * H;
    return o;
  }
  Aff<F> to_affine() const {
    if (inf()) return {F::zero(), F::zero()};
    F zi = Z.inv(), zi2 = zi.sqr();
    return {X * zi2, Y * zi2 * zi};
  }
};
template <class F> static Aff<F> aff_neg(const Aff<F>& p) { return {p.x, p.y.neg()}; }

static Fq fq_from_u64(u64 x) { u64 c[4] = {x, 0, 0, 0}; return Fq::from_canonical(c); }
static Fq curve_b(const Fq*) { return fq_from_u64(3); }
static Fq2 curve_b(const Fq2*) {  // 3/(9+u)
  Fq2 nine_u = {fq_from_u64(9), fq_from_u64(1)};
  Fq2 three = {fq_from_u64(3), Fq::zero()};
  return three * nine_u.inv();
}
template <class F> static bool on_curve(const Aff<F>& p) {
  if (p.inf()) return true;
  return p.y.sqr() == p.x.sqr() * p.x + curve_b((const F*)nullptr);
}

// scalar (canonical 4xu64, any 256-bit value is reduced mod r by the callers) times point
template <class F> static Jac<F> scalar_mul(const u64* k, const Aff<F>& p) {
  Jac<F> acc = Jac<F>::identity();
  for (int i = 255; i >= 0; --i) {
    acc = acc.dbl();
    if ((k[i >> 6] >> (i & 63)) & 1) acc = acc.add_mixed(p);
  }
  return acc;
}

static void reduce_mod_r(u64* s) {  // 256-bit value -> canonical Fr (at most 5 subtractions: 2^256/r < 6)
  while (geq4(s, FrP::MOD)) sub4(s, s, FrP::MOD);
}

// ------------------------------------------------------------------------------------------
// Pippenger, ark-ec 0.5.0 msm_bigint_wnaf
static int ark_window(size_t n) {
  if (n < 32) return 3;
  int lg = 63 - __builtin_clzll((unsigned long long)n);
  // ark_std::log2(n) is ceil(log2 n); ln_without_floats(a) = log2(a) * 69 / 100
  if (((size_t)1 << lg) != n) lg += 1;
  return lg * 69 / 100 + 2;
}
static void make_digits(const u64* s, int w, int num_bits, int64_t* out) {
  const u64 radix = 1ULL << w, mask = radix - 1;
  u64 carry = 0;
  int digits = (num_bits + w - 1) / w;
  for (int i = 0; i < digits; ++i) {
    int bit_offset = i * w, idx = bit_offset / 64, bit = bit_offset % 64;
    u64 buf;
    if (bit < 64 - w || idx == 3) buf = s[idx] >> bit;
    else buf = (s[idx] >> bit) | (s[idx + 1] << (64 - bit));
    u64 coef = carry + (buf & mask);
    carry = (coef + radix / 2) >> w;
    int64_t d = (int64_t)coef - (int64_t)(carry << w);
    if (i == digits - 1) d += (int64_t)(carry << w);
    out[i] = d;
  }
}

template <class F>
static Jac<F> pippenger_serial_window(const Aff<F>* pts, const int64_t* digits, int ndig, int w, size_t n, int c) {
  std::vector<Jac<F>> buckets((size_t)1 << (c - 1), Jac<F>::identity());
  for (size_t i = 0; i < n; ++i) {
    int64_t d = digits[i * ndig + w];
    if (d > 0) buckets[d - 1] = buckets[d - 1].add_mixed(pts[i]);
    else if (d < 0) buckets[-d - 1] = buckets[-d - 1].add_mixed(aff_neg(pts[i]));
  }
  Jac<F> run = Jac<F>::identity(), res = Jac<F>::identity();
  for (size_t b = buckets.size(); b-- > 0;) { run = run.add(buckets[b]); res = res.add(run); }
  return res;
}

template <class F>
static Jac<F> msm_pippenger(const Aff<F>* pts, const u64* scalars, size_t n, int threads) {
  if (n == 0) return Jac<F>::identity();
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
  // gnark-style outer split over points when there are more threads than windows;
  // ark's `parallel` feature is the nchunks == 1 case (one task per window).
  const int num_bits = 254;
  size_t nchunks = 1;
  {
    int c0 = ark_window(n), w0 = (num_bits + c0 - 1) / c0;
    if (threads > w0 && n >= (size_t)1 << 14) nchunks = (threads + w0 - 1) / w0;
  }
  size_t chunk = (n + nchunks - 1) / nchunks;
  const int c = ark_window(chunk);
  const int ndig = (num_bits + c - 1) / c;
  std::vector<int64_t> digits(n * (size_t)ndig);
#pragma omp parallel for num_threads(threads) schedule(static)
  for (size_t i = 0; i < n; ++i) {
    u64 s[4]; memcpy(s, scalars + 4 * i, 32); reduce_mod_r(s);
    make_digits(s, c, num_bits, &digits[i * ndig]);
  }
  std::vector<Jac<F>> wsum(nchunks * ndig);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (size_t t = 0; t < nchunks * (size_t)ndig; ++t) {
    size_t ch = t / ndig; int w = (int)(t % ndig);
    size_t lo = ch * chunk, hi = std::min(n, lo + chunk);
    wsum[t] = lo < hi ? pippenger_serial_window(pts + lo, &digits[lo * ndig], ndig, w, hi - lo, c)
                      : Jac<F>::identity();
  }
  Jac<F> total = Jac<F>::identity();
  for (size_t ch = 0; ch < nchunks; ++ch) {
    const Jac<F>* ws = &wsum[ch * ndig];
    Jac<F> acc = Jac<F>::identity();
    for (int w = ndig - 1; w >= 1; --w) {
      acc = acc.add(ws[w]);
      for (int k = 0; k < c; ++k) acc = acc.dbl();
    }
    total = total.add(acc.add(ws[0]));
  }
  return total;
}

// ------------------------------------------------------------------------------------------
// NTT, ark-poly 0.5.0 Radix2EvaluationDomain semantics
static Fr fr_root_2_28() {  // 5^((r-1)/2^28), canonical value from SURVEY.md section 8c
  const u64 g[4] = {0x9bd61b6e725b19f0ULL, 0x402d111e41112ed4ULL, 0x00e0a7eb8ef62abcULL, 0x2a3c09f0a58a7e85ULL};
  return Fr::from_canonical(g);
}
static Fr fr_pow_u64(Fr b, u64 e) {
  Fr acc = Fr::one();
  while (e) { if (e & 1) acc = acc * b; b = b.sqr(); e >>= 1; }
  return acc;
}
static void ntt_core(Fr* a, unsigned log_n, const Fr& w, int threads) {
  const size_t n = (size_t)1 << log_n;
  // bit-reversal permutation, then iterative decimation-in-time
  for (size_t i = 0; i < n; ++i) {
    size_t j = 0;
    for (unsigned b = 0; b < log_n; ++b) j |= ((i >> b) & 1) << (log_n - 1 - b);
    if (i < j) std::swap(a[i], a[j]);
  }
  std::vector<Fr> tw(n / 2 ? n / 2 : 1);
  tw[0] = Fr::one();
  for (size_t i = 1; i < n / 2; ++i) tw[i] = tw[i - 1] * w;
  for (unsigned s = 1; s <= log_n; ++s) {
    const size_t m = (size_t)1 << s, half = m >> 1, step = n >> s;
#pragma omp parallel for num_threads(threads) schedule(static) if (n >= 4096)
    for (size_t bf = 0; bf < n / 2; ++bf) {
      size_t blk = bf / half, j = bf % half;
      Fr* lo = a + blk * m + j; Fr* hi = lo + half;
      Fr t = *hi * tw[j * step];
      Fr u = *lo;
      *lo = u + t; *hi = u - t;
    }
  }
}

// ------------------------------------------------------------------------------------------
// byte conventions
static void be32_to_limbs(const uint8_t* be, u64* out) {
  for (int i = 0; i < 4; ++i) {
    u64 v = 0;
    for (int j = 0; j < 8; ++j) v = (v << 8) | be[(3 - i) * 8 + j];
    out[i] = v;
  }
}
static void limbs_to_be32(const u64* in, uint8_t* be) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 8; ++j) be[(3 - i) * 8 + j] = (uint8_t)(in[i] >> (56 - 8 * j));
}
static void g1_to_be(const Aff<Fq>& p, uint8_t* out) {
  u64 c[4]; p.x.to_canonical(c); limbs_to_be32(c, out); p.y.to_canonical(c); limbs_to_be32(c, out + 32);
}
static void g2_to_be(const Aff<Fq2>& p, uint8_t* out) {  // x_im | x_re | y_im | y_re
  u64 c[4];
  p.x.c1.to_canonical(c); limbs_to_be32(c, out);
  p.x.c0.to_canonical(c); limbs_to_be32(c, out + 32);
  p.y.c1.to_canonical(c); limbs_to_be32(c, out + 64);
  p.y.c0.to_canonical(c); limbs_to_be32(c, out + 96);
}
// returns 0 ok, 1 identity, 2 coordinate >= p, 3 not on curve
static int g1_from_be(const uint8_t* in, Aff<Fq>* out) {
  u64 x[4], y[4]; be32_to_limbs(in, x); be32_to_limbs(in + 32, y);
  if (geq4(x, FqP::MOD) || geq4(y, FqP::MOD)) return 2;
  out->x = Fq::from_canonical(x); out->y = Fq::from_canonical(y);
  if (out->inf()) return 1;
  return on_curve(*out) ? 0 : 3;
}
static int g2_from_be(const uint8_t* in, Aff<Fq2>* out) {
  u64 c[4][4];
  for (int k = 0; k < 4; ++k) { be32_to_limbs(in + 32 * k, c[k]); if (geq4(c[k], FqP::MOD)) return 2; }
  out->x = {Fq::from_canonical(c[1]), Fq::from_canonical(c[0])};
  out->y = {Fq::from_canonical(c[3]), Fq::from_canonical(c[2])};
  if (out->inf()) return 1;
  return on_curve(*out) ? 0 : 3;
}

// splitmix64 counter generator, identical to oracle/pyref.py rand_fr
static inline u64 splitmix64(u64& st) {
  st += 0x9E3779B97F4A7C15ULL;
  u64 z = st;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static void rand_fr_canonical(u64 seed, u64 index, u64* out) {
  u64 st = seed + (4 * index) * 0x9E3779B97F4A7C15ULL;
  for (int i = 0; i < 4; ++i) out[i] = splitmix64(st);
  reduce_mod_r(out);
}

template <class F>
static void chain_generate(Aff<F>* out, size_t n, const u64* k, const u64* d, const Aff<F>& gen, int threads) {
  // P_i = (k + i*d) * G ; each thread starts from a scalar multiplication, then adds D repeatedly
  // and normalises its run with one batched inversion.
  if (n == 0) return;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
  Aff<F> D = scalar_mul(d, gen).to_affine();
  Fr kf = Fr::from_canonical(k), df = Fr::from_canonical(d);
  const size_t run = 1024;
  const size_t nruns = (n + run - 1) / run;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 4)
  for (size_t r = 0; r < nruns; ++r) {
    size_t lo = r * run, hi = std::min(n, lo + run);
    u64 ic[4] = {lo, 0, 0, 0};
    Fr s = kf + Fr::from_canonical(ic) * df;
    u64 sc[4]; s.to_canonical(sc);
    std::vector<Jac<F>> js(hi - lo);
    Jac<F> cur = scalar_mul(sc, gen);
    for (size_t i = lo; i < hi; ++i) { js[i - lo] = cur; cur = cur.add_mixed(D); }
    for (size_t i = lo; i < hi; ++i) out[i] = js[i - lo].to_affine();
  }
}

// ------------------------------------------------------------------------------------------
extern "C" {

int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// --- field conversions: canonical little-endian limbs <-> Montgomery limbs (in place, n elements)
void orc_fq_to_mont(u64* v, size_t n) { for (size_t i = 0; i < n; ++i) { Fq f = Fq::from_canonical(v + 4 * i); memcpy(v + 4 * i, f.v, 32); } }
void orc_fq_from_mont(u64* v, size_t n) { for (size_t i = 0; i < n; ++i) { Fq f; memcpy(f.v, v + 4 * i, 32); f.to_canonical(v + 4 * i); } }
void orc_fr_to_mont(u64* v, size_t n) {
#pragma omp parallel for schedule(static) if (n >= 4096)
  for (size_t i = 0; i < n; ++i) { Fr f = Fr::from_canonical(v + 4 * i); memcpy(v + 4 * i, f.v, 32); }
}
void orc_fr_from_mont(u64* v, size_t n) {
#pragma omp parallel for schedule(static) if (n >= 4096)
  for (size_t i = 0; i < n; ++i) { Fr f; memcpy(f.v, v + 4 * i, 32); f.to_canonical(v + 4 * i); }
}
// raw Montgomery products (pins the CUDA field core): out[i] = a[i]*b[i]
void orc_fq_mul(const u64* a, const u64* b, u64* out, size_t n) {
  for (size_t i = 0; i < n; ++i) { Fq x, y; memcpy(x.v, a + 4 * i, 32); memcpy(y.v, b + 4 * i, 32); Fq z = x * y; memcpy(out + 4 * i, z.v, 32); }
}
void orc_fr_mul(const u64* a, const u64* b, u64* out, size_t n) {
  for (size_t i = 0; i < n; ++i) { Fr x, y; memcpy(x.v, a + 4 * i, 32); memcpy(y.v, b + 4 * i, 32); Fr z = x * y; memcpy(out + 4 * i, z.v, 32); }
}

// --- synthetic inputs
// out[i] = (a[i]*b[i] - c[i]) * z : the pointwise step of the Groth16 quotient on the coset (all Montgomery limbs, z one element)
void orc_fr_quotient(const u64* a, const u64* b, const u64* c, const u64* z, u64* out, size_t n, int threads) {
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
  Fr zz; memcpy(zz.v, z, 32);
#pragma omp parallel for num_threads(threads) schedule(static)
  for (size_t i = 0; i < n; ++i) {
    Fr x, y, w; memcpy(x.v, a + 4 * i, 32); memcpy(y.v, b + 4 * i, 32); memcpy(w.v, c + 4 * i, 32);
    Fr r = (x * y - w) * zz;
    memcpy(out + 4 * i, r.v, 32);
  }
}
void orc_rand_fr(u64* out_canonical, u64 seed, u64 start, size_t n) {
#pragma omp parallel for schedule(static) if (n >= 4096)
  for (size_t i = 0; i < n; ++i) rand_fr_canonical(seed, start + i, out_canonical + 4 * i);
}
// native G1 affine: x|y Montgomery limbs (64 B), (0,0) = identity
void orc_g1_chain(u64* out_native, size_t n, const u64* k, const u64* d, int threads) {
  Aff<Fq> G = {fq_from_u64(1), fq_from_u64(2)};
  chain_generate((Aff<Fq>*)out_native, n, k, d, G, threads);
}
static Aff<Fq2> g2_generator() {
  static const uint8_t be[128] = {
    0x19,0x8e,0x93,0x93,0x92,0x0d,0x48,0x3a,0x72,0x60,0xbf,0xb7,0x31,0xfb,0x5d,0x25,0xf1,0xaa,0x49,0x33,0x35,0xa9,0xe7,0x12,0x97,0xe4,0x85,0xb7,0xae,0xf3,0x12,0xc2,
    0x18,0x00,0xde,0xef,0x12,0x1f,0x1e,0x76,0x42,0x6a,0x00,0x66,0x5e,0x5c,0x44,0x79,0x67,0x43,0x22,0xd4,0xf7,0x5e,0xda,0xdd,0x46,0xde,0xbd,0x5c,0xd9,0x92,0xf6,0xed,
    0x09,0x06,0x89,0xd0,0x58,0x5f,0xf0,0x75,0xec,0x9e,0x99,0xad,0x69,0x0c,0x33,0x95,0xbc,0x4b,0x31,0x33,0x70,0xb3,0x8e,0xf3,0x55,0xac,0xda,0xdc,0xd1,0x22,0x97,0x5b,
    0x12,0xc8,0x5e,0xa5,0xdb,0x8c,0x6d,0xeb,0x4a,0xab,0x71,0x80,0x8d,0xcb,0x40,0x8f,0xe3,0xd1,0xe7,0x69,0x0c,0x43,0xd3,0x7b,0x4c,0xe6,0xcc,0x01,0x66,0xfa,0x7d,0xaa};
  Aff<Fq2> g; g2_from_be(be, &g); return g;
}
// native G2 affine: x.c0|x.c1|y.c0|y.c1 Montgomery limbs (128 B)
void orc_g2_chain(u64* out_native, size_t n, const u64* k, const u64* d, int threads) {
  chain_generate((Aff<Fq2>*)out_native, n, k, d, g2_generator(), threads);
}
// sum_i s_i * (k + i*d) mod r  (closed-form scalar of an MSM over a chain), canonical in/out
void orc_chain_dot(const u64* scalars, size_t n, const u64* k, const u64* d, u64* out) {
  Fr kf = Fr::from_canonical(k), df = Fr::from_canonical(d);
  int T = orc_num_threads();
  std::vector<Fr> part(T, Fr::zero());
#pragma omp parallel num_threads(T)
  {
#ifdef _OPENMP
    int t = omp_get_thread_num();
#else
    int t = 0;
#endif
    size_t lo = n * t / T, hi = n * (t + 1) / T;
    u64 ic[4] = {lo, 0, 0, 0};
    Fr cur = kf + Fr::from_canonical(ic) * df, acc = Fr::zero();
    for (size_t i = lo; i < hi; ++i) {
      u64 s[4]; memcpy(s, scalars + 4 * i, 32); reduce_mod_r(s);
      acc = acc + Fr::from_canonical(s) * cur; cur = cur + df;
    }
    part[t] = acc;
  }
  Fr tot = Fr::zero();
  for (auto& p : part) tot = tot + p;
  tot.to_canonical(out);
}

// --- encodings
int orc_g1_be_to_native(const uint8_t* be, u64* out_native, size_t n) {
  for (size_t i = 0; i < n; ++i) { int rc = g1_from_be(be + 64 * i, (Aff<Fq>*)(out_native + 8 * i)); if (rc >= 2) return rc; }
  return 0;
}
void orc_g1_native_to_be(const u64* native, uint8_t* be, size_t n) { for (size_t i = 0; i < n; ++i) g1_to_be(*(const Aff<Fq>*)(native + 8 * i), be + 64 * i); }
int orc_g2_be_to_native(const uint8_t* be, u64* out_native, size_t n) {
  for (size_t i = 0; i < n; ++i) { int rc = g2_from_be(be + 128 * i, (Aff<Fq2>*)(out_native + 16 * i)); if (rc >= 2) return rc; }
  return 0;
}
void orc_g2_native_to_be(const u64* native, uint8_t* be, size_t n) { for (size_t i = 0; i < n; ++i) g2_to_be(*(const Aff<Fq2>*)(native + 16 * i), be + 128 * i); }

// --- single-point ops with the provider.rs / zisk.rs conventions (KAT pins)
int orc_g1_add_be(const uint8_t* p1, const uint8_t* p2, uint8_t* ret) {
  Aff<Fq> a, b; int r1 = g1_from_be(p1, &a); if (r1 >= 2) return r1; int r2 = g1_from_be(p2, &b); if (r2 >= 2) return r2;
  Jac<Fq> s = Jac<Fq>::identity().add_mixed(a).add_mixed(b);
  Aff<Fq> o = s.to_affine(); g1_to_be(o, ret); return o.inf() ? 1 : 0;
}
int orc_g1_mul_be(const uint8_t* point, const uint8_t* scalar_be, uint8_t* ret) {
  Aff<Fq> a; int r1 = g1_from_be(point, &a); if (r1 >= 2) return r1;
  u64 s[4]; be32_to_limbs(scalar_be, s); reduce_mod_r(s);
  Aff<Fq> o = scalar_mul(s, a).to_affine(); g1_to_be(o, ret); return o.inf() ? 1 : 0;
}
int orc_g2_mul_be(const uint8_t* point, const uint8_t* scalar_be, uint8_t* ret) {
  Aff<Fq2> a; int r1 = g2_from_be(point, &a); if (r1 >= 2) return r1;
  u64 s[4]; be32_to_limbs(scalar_be, s); reduce_mod_r(s);
  Aff<Fq2> o = scalar_mul(s, a).to_affine(); g2_to_be(o, ret); return o.inf() ? 1 : 0;
}

// --- MSM: native points (Montgomery), canonical little-endian scalars (reduced mod r here),
//     big-endian affine result.  method 0 = Pippenger (ark-ec rule), 1 = naive double-and-add.
int orc_g1_msm(const u64* points_native, const u64* scalars, size_t n, int method, int threads, uint8_t* out_be) {
  const Aff<Fq>* pts = (const Aff<Fq>*)points_native;
  Jac<Fq> acc = Jac<Fq>::identity();
  if (method == 1) {
    for (size_t i = 0; i < n; ++i) { u64 s[4]; memcpy(s, scalars + 4 * i, 32); reduce_mod_r(s); acc = acc.add(scalar_mul(s, pts[i])); }
  } else acc = msm_pippenger(pts, scalars, n, threads);
  Aff<Fq> o = acc.to_affine(); g1_to_be(o, out_be); return o.inf() ? 1 : 0;
}
int orc_g2_msm(const u64* points_native, const u64* scalars, size_t n, int method, int threads, uint8_t* out_be) {
  const Aff<Fq2>* pts = (const Aff<Fq2>*)points_native;
  Jac<Fq2> acc = Jac<Fq2>::identity();
  if (method == 1) {
    for (size_t i = 0; i < n; ++i) { u64 s[4]; memcpy(s, scalars + 4 * i, 32); reduce_mod_r(s); acc = acc.add(scalar_mul(s, pts[i])); }
  } else acc = msm_pippenger(pts, scalars, n, threads);
  Aff<Fq2> o = acc.to_affine(); g2_to_be(o, out_be); return o.inf() ? 1 : 0;
}
int orc_msm_window(size_t n) { return ark_window(n); }

// --- NTT over Fr, in place, Montgomery limbs, natural order in and out.
//     flags bit0 = inverse, bit1 = coset (coset_gen = canonical LE limbs; NULL -> 5).
//     root_2_28: canonical LE limbs of the 2^28-th root of unity, NULL -> ark/gnark 5^((r-1)/2^28).
enum { ORC_NTT_INVERSE = 1, ORC_NTT_COSET = 2 };
int orc_fr_ntt(u64* data, unsigned log_n, unsigned flags, const u64* coset_gen, const u64* root_2_28, int threads) {
  if (log_n > 28) return 4;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
  Fr* a = (Fr*)data;
  const size_t n = (size_t)1 << log_n;
  Fr g = root_2_28 ? Fr::from_canonical(root_2_28) : fr_root_2_28();
  Fr w = g;
  for (unsigned i = log_n; i < 28; ++i) w = w.sqr();
  Fr h = Fr::from_canonical((const u64[4]){5, 0, 0, 0});
  if (coset_gen) h = Fr::from_canonical(coset_gen);
  const bool inverse = flags & ORC_NTT_INVERSE, coset = flags & ORC_NTT_COSET;
  const size_t blk = 4096;
  if (!inverse && coset) {
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t b = 0; b < (n + blk - 1) / blk; ++b) {
      Fr x = fr_pow_u64(h, b * blk);
      for (size_t j = b * blk; j < std::min(n, (b + 1) * blk); ++j) { a[j] = a[j] * x; x = x * h; }
    }
  }
  ntt_core(a, log_n, inverse ? w.inv() : w, threads);
  if (inverse) {
    u64 nc[4] = {n, 0, 0, 0};
    Fr ninv = Fr::from_canonical(nc).inv();
    Fr hinv = coset ? h.inv() : Fr::one();
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t b = 0; b < (n + blk - 1) / blk; ++b) {
      Fr x = coset ? fr_pow_u64(hinv, b * blk) * ninv : ninv;
      for (size_t j = b * blk; j < std::min(n, (b + 1) * blk); ++j) { a[j] = a[j] * x; if (coset) x = x * hinv; }
    }
  }
  return 0;
}

// out = sum_j a[j] * x^j with x = w_n^k (Horner, O(n)): the definition of output k of the forward transform.
// Lets the tests spot-check transforms that are too large to run on the CPU in full.  Montgomery in/out.
int orc_fr_ntt_eval_output(const u64* data, unsigned log_n, u64 k, u64* out) {
  if (log_n > 28) return 4;
  const Fr* a = (const Fr*)data;
  const size_t n = (size_t)1 << log_n;
  Fr w = fr_root_2_28();
  for (unsigned i = log_n; i < 28; ++i) w = w.sqr();
  Fr x = fr_pow_u64(w, k % n);
  int T = 1;
#ifdef _OPENMP
  T = omp_get_max_threads();
  if (T > 16) T = 16;
#endif
  std::vector<Fr> part(T, Fr::zero());
  std::vector<size_t> lo(T + 1);
  for (int t = 0; t <= T; ++t) lo[t] = n * t / T;
#pragma omp parallel for num_threads(T) schedule(static, 1)
  for (int t = 0; t < T; ++t) {
    Fr acc = Fr::zero();
    for (size_t j = lo[t + 1]; j-- > lo[t];) acc = acc * x + a[j];  // sum_{j in block} a[j] x^(j - lo)
    part[t] = acc * fr_pow_u64(x, lo[t]);
  }
  Fr tot = Fr::zero();
  for (auto& p : part) tot = tot + p;
  memcpy(out, tot.v, 32);
  return 0;
}

}  // extern "C"
