// util.cu -- device-side format conversion, validation and synthetic-workload generators.
// Byte conventions: /root/reference/crates/common/crypto/provider.rs:201-330 (EIP-196/197 big-endian,
// (0,0) = identity, G2 = x_im|x_re|y_im|y_re) and /root/reference/crates/vm/levm/src/precompiles.rs:801-820
// (coordinates >= p are rejected).  Synthetic inputs: SURVEY.md section 8d.
#include "common.cuh"

namespace b200zk {

// ---- field helpers --------------------------------------------------------------------------------------------
template <class Fld> B2_D Fld load_be32(const uint8_t* in, bool* in_range) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(in);
  Fld v;
#pragma unroll
  for (int k = 0; k < 8; ++k) v.v[k] = __byte_perm(__ldg(w + 7 - k), 0, 0x0123);
  Fld m = Fld::modulus(), t;
  *in_range = detail::sub8(t.v, v.v, m.v) != 0;  // borrow <=> v < modulus
  return v;
}

template <class Fld>
__global__ void __launch_bounds__(256) field_convert(void* data, size_t n, int to_mont) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fld v = load_fe<Fld>(data, i);
    if (to_mont) {
      for (int k = 0; k < 5; ++k) { Fld m = Fld::modulus(), t; if (!detail::sub8(t.v, v.v, m.v)) v = t; }
      v = Fld::to_mont(v);
    } else v = Fld::from_mont(v);
    store_fe<Fld>(data, i, v);
  }
}

template <class Fld, bool SQUARE>
__global__ void __launch_bounds__(256) field_mul_kernel(const void* a, const void* b, void* out, size_t n, uint32_t repeat) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fld x = load_fe<Fld>(a, i), y = load_fe<Fld>(b, i);
    for (uint32_t r = 0; r < repeat; ++r) x = SQUARE ? Fld::sqr(x) : Fld::mul(x, y);
    store_fe<Fld>(out, i, x);
  }
}

// out[i] = (a[i]*b[i] - c[i]) * zinv : the Groth16 quotient numerator on the coset, divided by the (constant)
// vanishing polynomial value h^n - 1.  All Montgomery form; zinv given canonical.
__global__ void __launch_bounds__(256) fr_quotient(const void* a, const void* b, const void* c, void* out, size_t n, const uint32_t* zinv_canonical) {
  Fr z;
#pragma unroll
  for (int k = 0; k < 8; ++k) z.v[k] = zinv_canonical[k];
  z = Fr::to_mont(z);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fr t = Fr::sub(Fr::mul(load_fe<Fr>(a, i), load_fe<Fr>(b, i)), load_fe<Fr>(c, i));
    store_fe<Fr>(out, i, Fr::mul(t, z));
  }
}

// zinv = 1 / (5^(2^log_n) - 1), canonical limbs: the inverse of the vanishing polynomial of the size-2^log_n domain on the
// coset 5 * <w> (a constant there).  One thread, once per domain size.
__global__ void fr_coset_zinv(uint32_t log_n, uint32_t* out_canonical) {
  if (blockIdx.x || threadIdx.x) return;
  Fr g = Fr::zero(); g.v[0] = 5;
  g = Fr::to_mont(g);
  for (uint32_t i = 0; i < log_n; ++i) g = Fr::sqr(g);
  Fr z = Fr::from_mont(Fr::inv(Fr::sub(g, Fr::one())));
#pragma unroll
  for (int k = 0; k < 8; ++k) out_canonical[k] = z.v[k];
}

// ---- splitmix64 counter generator (identical to oracle/pyref.py rand_fr and the C++ oracle) -----------------------
B2_D uint64_t splitmix64(uint64_t& st) {
  st += 0x9E3779B97F4A7C15ull;
  uint64_t z = st;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256) fr_random(void* out, size_t n, uint64_t seed, uint64_t start, int mont) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint64_t st = seed + (4 * (start + i)) * 0x9E3779B97F4A7C15ull;
    Fr v;
#pragma unroll
    for (int k = 0; k < 4; ++k) { uint64_t z = splitmix64(st); v.v[2 * k] = (uint32_t)z; v.v[2 * k + 1] = (uint32_t)(z >> 32); }
    for (int k = 0; k < 5; ++k) { Fr m = Fr::modulus(), t; if (!detail::sub8(t.v, v.v, m.v)) v = t; }
    if (mont) v = Fr::to_mont(v);
    store_fe<Fr>(out, i, v);
  }
}

// ---- generators ---------------------------------------------------------------------------------------------------
template <class F> struct Generator;
template <> struct Generator<Fq> {
  static B2_D Affine<Fq> get() {
    Fq x = Fq::zero(), y = Fq::zero(); x.v[0] = 1; y.v[0] = 2;
    return {Fq::to_mont(x), Fq::to_mont(y)};
  }
};
template <> struct Generator<Fq2> {
  static B2_D Affine<Fq2> get() {  // EIP-197 generator (/root/reference/test/tests/levm/precompile_tests.rs:17-24, pair 1)
    const uint32_t xr[8] = {0xd992f6edu, 0x46debd5cu, 0xf75edaddu, 0x674322d4u, 0x5e5c4479u, 0x426a0066u, 0x121f1e76u, 0x1800deefu};
    const uint32_t xi[8] = {0xaef312c2u, 0x97e485b7u, 0x35a9e712u, 0xf1aa4933u, 0x31fb5d25u, 0x7260bfb7u, 0x920d483au, 0x198e9393u};
    const uint32_t yr[8] = {0x66fa7daau, 0x4ce6cc01u, 0x0c43d37bu, 0xe3d1e769u, 0x8dcb408fu, 0x4aab7180u, 0xdb8c6debu, 0x12c85ea5u};
    const uint32_t yi[8] = {0xd122975bu, 0x55acdadcu, 0x70b38ef3u, 0xbc4b3133u, 0x690c3395u, 0xec9e99adu, 0x585ff075u, 0x090689d0u};
    Fq a, b, c, d;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a.v[i] = xr[i]; b.v[i] = xi[i]; c.v[i] = yr[i]; d.v[i] = yi[i]; }
    return {{Fq::to_mont(a), Fq::to_mont(b)}, {Fq::to_mont(c), Fq::to_mont(d)}};
  }
};

// scratch[0] = D = d*G (affine)
template <class F> __global__ void chain_setup(const uint32_t* d_canonical, void* scratch) {
  if (blockIdx.x || threadIdx.x) return;
  uint32_t d[8];
  for (int i = 0; i < 8; ++i) d[i] = d_canonical[i];
  store_affine<F>(scratch, 0, xyzz_to_affine(xyzz_scalar_mul<F>(d, Generator<F>::get())));
}
// out[i - start] = (k + i*d) * G for i in [start, start + n): each thread owns a run of kRun consecutive i
static constexpr int kRun = 32;
template <class F>
__global__ void __launch_bounds__(128) chain_fill(void* out, size_t start, size_t n, const uint32_t* k_canonical, const uint32_t* d_canonical, const void* scratch) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t lo = t * kRun;
  if (lo >= n) return;
  size_t hi = lo + kRun < n ? lo + kRun : n;
  Fr kf, df, idx = Fr::zero();
  for (int i = 0; i < 8; ++i) { kf.v[i] = k_canonical[i]; df.v[i] = d_canonical[i]; }
  uint64_t i0 = start + lo;
  idx.v[0] = (uint32_t)i0; idx.v[1] = (uint32_t)(i0 >> 32);
  Fr s = Fr::from_mont(Fr::add(Fr::to_mont(kf), Fr::mul(Fr::to_mont(idx), Fr::to_mont(df))));
  Affine<F> D = load_affine_nc<F>(scratch, 0);
  XYZZ<F> cur = xyzz_scalar_mul<F>(s.v, Generator<F>::get());
  for (size_t i = lo; i < hi; ++i) {
    store_affine<F>(out, i, xyzz_to_affine(cur));
    xyzz_add_mixed(cur, D.x, D.y);
  }
}

// ---- validation / decoding -----------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128) points_check(const void* pts, size_t n, unsigned long long* bad) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Affine<F> p = load_affine_nc<F>(pts, i);
    if (!affine_on_curve(p)) atomicMin(bad, (unsigned long long)i);
  }
}
// status[0] = min index with coordinate >= p, status[1] = min index not on curve (init: ~0)
__global__ void __launch_bounds__(128) g1_decode_be(const uint8_t* be, void* native, size_t n, unsigned long long* status) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    bool okx, oky;
    Fq x = load_be32<Fq>(be + 64 * i, &okx), y = load_be32<Fq>(be + 64 * i + 32, &oky);
    if (!(okx && oky)) { atomicMin(status, (unsigned long long)i); continue; }
    Affine<Fq> p = {Fq::to_mont(x), Fq::to_mont(y)};
    if (!affine_on_curve(p)) atomicMin(status + 1, (unsigned long long)i);
    store_affine<Fq>(native, i, p);
  }
}
__global__ void __launch_bounds__(128) g2_decode_be(const uint8_t* be, void* native, size_t n, unsigned long long* status) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    bool ok[4];
    const uint8_t* b = be + 128 * i;
    Fq xi = load_be32<Fq>(b, &ok[0]), xr = load_be32<Fq>(b + 32, &ok[1]), yi = load_be32<Fq>(b + 64, &ok[2]), yr = load_be32<Fq>(b + 96, &ok[3]);
    if (!(ok[0] && ok[1] && ok[2] && ok[3])) { atomicMin(status, (unsigned long long)i); continue; }
    Affine<Fq2> p = {{Fq::to_mont(xr), Fq::to_mont(xi)}, {Fq::to_mont(yr), Fq::to_mont(yi)}};
    if (!affine_on_curve(p)) atomicMin(status + 1, (unsigned long long)i);
    store_affine<Fq2>(native, i, p);
  }
}

// ---- host wrappers ---------------------------------------------------------------------------------------------------
static unsigned egrid(b200zk_ctx* ctx, size_t n, unsigned block) {
  size_t g = (n + block - 1) / block, cap = (size_t)ctx->sm_count * 32;
  return (unsigned)(g < cap ? (g ? g : 1) : cap);
}

int fr_quotient_dev(b200zk_ctx* ctx, const void* d_a, const void* d_b, const void* d_c, void* d_out, size_t n, const uint32_t* d_zinv_canonical, cudaStream_t st) {
  if (!n) return B200ZK_OK;
  B2_LAUNCH(ctx, fr_quotient, egrid(ctx, n, 256), 256, 0, st, d_a, d_b, d_c, d_out, n, d_zinv_canonical);
  return B200ZK_OK;
}
int fr_coset_zinv_dev(b200zk_ctx* ctx, uint32_t log_n, cudaStream_t st, const uint32_t** d_zinv) {
  B2_TRY(ensure(ctx, ctx->ws_zinv, 64));
  if (ctx->zinv_log_n != log_n) {
    B2_LAUNCH(ctx, fr_coset_zinv, 1, 32, 0, st, log_n, (uint32_t*)ctx->ws_zinv.p);
    B2_CUDA(ctx, cudaStreamSynchronize(st));  // one-off per domain size: later calls on other streams may read it unordered
    ctx->zinv_log_n = log_n;
  }
  *d_zinv = (const uint32_t*)ctx->ws_zinv.p;
  return B200ZK_OK;
}
int points_be_to_native(b200zk_ctx* ctx, const void* d_be, void* d_native, size_t n, bool g2, cudaStream_t st) {
  B2_TRY(ensure(ctx, ctx->ws_out, 256));
  unsigned long long* status = (unsigned long long*)((uint8_t*)ctx->ws_out.p + 192);
  B2_CUDA(ctx, cudaMemsetAsync(status, 0xff, 16, st));
  if (n) {
    if (g2) B2_LAUNCH(ctx, g2_decode_be, egrid(ctx, n, 128), 128, 0, st, (const uint8_t*)d_be, d_native, n, status);
    else B2_LAUNCH(ctx, g1_decode_be, egrid(ctx, n, 128), 128, 0, st, (const uint8_t*)d_be, d_native, n, status);
  }
  unsigned long long* h = (unsigned long long*)(ctx->h_pinned + 1024);
  B2_CUDA(ctx, cudaMemcpyAsync(h, status, 16, cudaMemcpyDeviceToHost, st));
  B2_CUDA(ctx, cudaStreamSynchronize(st));
  if (h[0] != ~0ull && (h[1] == ~0ull || h[0] <= h[1])) return fail(ctx, B200ZK_ERR_NOT_IN_FIELD, "point coordinate >= field modulus");
  if (h[1] != ~0ull) return fail(ctx, B200ZK_ERR_NOT_ON_CURVE, "point not on curve");
  return B200ZK_OK;
}

}  // namespace b200zk

using namespace b200zk;

extern "C" {

int b200zk_field_to_mont_device(b200zk_ctx* ctx, void* d, size_t n, int which, void* stream) { b200zk::DeviceGuard guard(ctx);
  if (!ctx || (!d && n)) return fail(ctx, B200ZK_ERR_INVALID_ARG, "field_to_mont: null argument");
  cudaStream_t st = pick_stream(ctx, stream);
  if (!n) return B200ZK_OK;
  if (which == 0) B2_LAUNCH(ctx, field_convert<Fq>, egrid(ctx, n, 256), 256, 0, st, d, n, 1);
  else B2_LAUNCH(ctx, field_convert<Fr>, egrid(ctx, n, 256), 256, 0, st, d, n, 1);
  return B200ZK_OK;
}
int b200zk_field_from_mont_device(b200zk_ctx* ctx, void* d, size_t n, int which, void* stream) { b200zk::DeviceGuard guard(ctx);
  if (!ctx || (!d && n)) return fail(ctx, B200ZK_ERR_INVALID_ARG, "field_from_mont: null argument");
  cudaStream_t st = pick_stream(ctx, stream);
  if (!n) return B200ZK_OK;
  if (which == 0) B2_LAUNCH(ctx, field_convert<Fq>, egrid(ctx, n, 256), 256, 0, st, d, n, 0);
  else B2_LAUNCH(ctx, field_convert<Fr>, egrid(ctx, n, 256), 256, 0, st, d, n, 0);
  return B200ZK_OK;
}
int b200zk_field_mul_device(b200zk_ctx* ctx, const void* a, const void* b, void* out, size_t n, int which, uint32_t repeat, void* stream) { b200zk::DeviceGuard guard(ctx);
  if (!ctx || ((!a || !b || !out) && n)) return fail(ctx, B200ZK_ERR_INVALID_ARG, "field_mul: null argument");
  cudaStream_t st = pick_stream(ctx, stream);
  if (!n) return B200ZK_OK;
  if (!repeat) repeat = 1;
  // throughput runs want every SM saturated: no grid cap below n/256
  unsigned grid = (unsigned)((n + 255) / 256);
  // which: bit 0 = field (0 Fq, 1 Fr), bit 1 = squaring (out = a^2, repeated: a^(2^repeat); b is not used)
  if (which == 0) B2_LAUNCH(ctx, (field_mul_kernel<Fq, false>), grid, 256, 0, st, a, b, out, n, repeat);
  else if (which == 1) B2_LAUNCH(ctx, (field_mul_kernel<Fr, false>), grid, 256, 0, st, a, b, out, n, repeat);
  else if (which == 2) B2_LAUNCH(ctx, (field_mul_kernel<Fq, true>), grid, 256, 0, st, a, b, out, n, repeat);
  else if (which == 3) B2_LAUNCH(ctx, (field_mul_kernel<Fr, true>), grid, 256, 0, st, a, b, out, n, repeat);
  else return fail(ctx, B200ZK_ERR_INVALID_ARG, "field_mul: which must be 0..3");
  return B200ZK_OK;
}
int b200zk_fr_quotient_device(b200zk_ctx* ctx, const void* d_a, const void* d_b, const void* d_c, void* d_out, size_t n, const uint8_t zinv[32], void* stream) { b200zk::DeviceGuard guard(ctx);
  if (!ctx || !zinv || ((!d_a || !d_b || !d_c || !d_out) && n)) return fail(ctx, B200ZK_ERR_INVALID_ARG, "fr_quotient: null argument");
  cudaStream_t st = pick_stream(ctx, stream);
  if (!n) return B200ZK_OK;
  B2_TRY(ensure(ctx, ctx->ws_misc, 512));
  memcpy(ctx->h_pinned + 3072, zinv, 32);
  B2_CUDA(ctx, cudaMemcpyAsync((uint8_t*)ctx->ws_misc.p + 384, ctx->h_pinned + 3072, 32, cudaMemcpyHostToDevice, st));
  B2_LAUNCH(ctx, fr_quotient, egrid(ctx, n, 256), 256, 0, st, d_a, d_b, d_c, d_out, n, (const uint32_t*)((uint8_t*)ctx->ws_misc.p + 384));
  B2_CUDA(ctx, cudaStreamSynchronize(st));  // staging buffers are reused by the next call
  return B200ZK_OK;
}
int b200zk_fr_random_device(b200zk_ctx* ctx, void* d_out, size_t n, uint64_t seed, uint64_t start, uint32_t flags, void* stream) { b200zk::DeviceGuard guard(ctx);
  if (!ctx || (!d_out && n)) return fail(ctx, B200ZK_ERR_INVALID_ARG, "fr_random: null argument");
  cudaStream_t st = pick_stream(ctx, stream);
  if (!n) return B200ZK_OK;
  B2_LAUNCH(ctx, fr_random, egrid(ctx, n, 256), 256, 0, st, d_out, n, seed, start, (flags & B200ZK_SCALARS_MONT) ? 1 : 0);
  return B200ZK_OK;
}

}  // extern "C"
template <class F>
static int chain_device(b200zk_ctx* ctx, void* d_out, size_t start, size_t n, const uint8_t k[32], const uint8_t d[32], void* stream) {
  if (!ctx || !k || !d || (!d_out && n)) return fail(ctx, B200ZK_ERR_INVALID_ARG, "chain: null argument");
  cudaStream_t st = pick_stream(ctx, stream);
  if (!n) return B200ZK_OK;
  B2_TRY(ensure(ctx, ctx->ws_misc, 512));
  uint8_t* base = (uint8_t*)ctx->ws_misc.p;
  memcpy(ctx->h_pinned + 2048, k, 32);
  memcpy(ctx->h_pinned + 2048 + 32, d, 32);
  B2_CUDA(ctx, cudaMemcpyAsync(base, ctx->h_pinned + 2048, 64, cudaMemcpyHostToDevice, st));
  B2_LAUNCH(ctx, chain_setup<F>, 1, 32, 0, st, (const uint32_t*)(base + 32), base + 128);
  size_t runs = (n + kRun - 1) / kRun;
  B2_LAUNCH(ctx, chain_fill<F>, (unsigned)((runs + 127) / 128), 128, 0, st, d_out, start, n, (const uint32_t*)base, (const uint32_t*)(base + 32), base + 128);
  B2_CUDA(ctx, cudaStreamSynchronize(st));  // h_pinned staging is reused
  return B200ZK_OK;
}
extern "C" {
int b200zk_g1_chain_device(b200zk_ctx* ctx, void* d_out, size_t start, size_t n, const uint8_t k[32], const uint8_t d[32], void* stream) { b200zk::DeviceGuard guard(ctx);
  return chain_device<Fq>(ctx, d_out, start, n, k, d, stream);
}
int b200zk_g2_chain_device(b200zk_ctx* ctx, void* d_out, size_t start, size_t n, const uint8_t k[32], const uint8_t d[32], void* stream) { b200zk::DeviceGuard guard(ctx);
  return chain_device<Fq2>(ctx, d_out, start, n, k, d, stream);
}

}  // extern "C"
template <class F>
static int check_device(b200zk_ctx* ctx, const void* d_points, size_t n, void* stream, size_t* bad_index) {
  if (!ctx || !bad_index || (!d_points && n)) return fail(ctx, B200ZK_ERR_INVALID_ARG, "check: null argument");
  cudaStream_t st = pick_stream(ctx, stream);
  *bad_index = n;
  if (!n) return B200ZK_OK;
  B2_TRY(ensure(ctx, ctx->ws_out, 256));
  unsigned long long* flag = (unsigned long long*)((uint8_t*)ctx->ws_out.p + 192);
  B2_CUDA(ctx, cudaMemsetAsync(flag, 0xff, 8, st));
  B2_LAUNCH(ctx, points_check<F>, egrid(ctx, n, 128), 128, 0, st, d_points, n, flag);
  unsigned long long* h = (unsigned long long*)(ctx->h_pinned + 1024);
  B2_CUDA(ctx, cudaMemcpyAsync(h, flag, 8, cudaMemcpyDeviceToHost, st));
  B2_CUDA(ctx, cudaStreamSynchronize(st));
  if (*h != ~0ull) { *bad_index = (size_t)*h; return fail(ctx, B200ZK_ERR_NOT_ON_CURVE, "point not on curve"); }
  return B200ZK_OK;
}
extern "C" {
int b200zk_g1_check_device(b200zk_ctx* ctx, const void* p, size_t n, void* stream, size_t* bad) { b200zk::DeviceGuard guard(ctx); return check_device<Fq>(ctx, p, n, stream, bad); }
int b200zk_g2_check_device(b200zk_ctx* ctx, const void* p, size_t n, void* stream, size_t* bad) { b200zk::DeviceGuard guard(ctx); return check_device<Fq2>(ctx, p, n, stream, bad); }

}  // extern "C"
