"""oracle/pyref.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Independent pure-Python big-integer statement of the BN254 arithmetic on the
hot path.  It shares no code with the C++
oracle (oracle/bn254_oracle.cpp) nor with the CUDA library: it is the second,
slow, obviously-correct implementation that both are pinned against, and the
generator of the golden fixtures under tests/golden/.

PARITY STATUS: "parity unpinned" for MSM/NTT outputs -- the reference
(lambdaclass/ethrex @ 823d3abc) holds no MSM/NTT implementation, test or
golden vector.  What *is* pinned by the reference:
  * encodings: 32-byte big-endian canonical coordinates, (0,0) = identity,
    G2 = x_im | x_re | y_im | y_re
    (crates/common/crypto/provider.rs:201-330, crates/vm/levm/src/precompiles.rs:775-799);
  * field modulus ALT_BN128_PRIME (crates/vm/levm/src/precompiles.rs:746-751);
  * on-curve KAT points of test/tests/levm/precompile_tests.rs:17-24 and the
    ecmul KAT 7*(1,2) of test/tests/l2/integration_tests.rs:572;
  * all 14 `ecpairing` known-answer vectors of test/tests/levm/precompile_tests.rs:17-140
    (tests/golden/pairing_kats.json): the optimal-ate pairing at the end of this file, built on the same
    Fq/Fq2/G1/G2 code, reproduces every expected boolean (tests/test_oracle.py), and MSM results are tied to
    that pairing by bilinearity.
The MSM/NTT *semantics* restated here are those of the un-vendored third-party
crates pinned in the reference's Cargo.lock: ark-ec 0.5.0
`VariableBaseMSM::msm` (result = sum s_i*P_i, normalised to affine) and
ark-poly 0.5.0 `Radix2EvaluationDomain::{fft,ifft,coset_fft}` (Cargo.lock:978,1140).
"""
from __future__ import annotations

# --------------------------------------------------------------------------- constants
P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # base field Fq
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # scalar field Fr
MONT = 1 << 256
TWO_ADICITY = 28
FR_GENERATOR = 5  # ark-bn254 Fr::GENERATOR, gnark-crypto bn254/fr multiplicative generator
# ark-ff TWO_ADIC_ROOT_OF_UNITY = GENERATOR^((r-1)/2^28)
ROOT_2_28 = pow(FR_GENERATOR, (R - 1) >> TWO_ADICITY, R)
assert ROOT_2_28 == 19103219067921713944291392827692070036145651957329286315305642004821462161904
G1_GEN = (1, 2)
# EIP-197 G2 generator as ((x_re, x_im), (y_re, y_im)); precompile_tests.rs:17-24 pair 1.
G2_GEN = (
    (0x1800DEEF121F1E76426A00665E5C4479674322D4F75EDADD46DEBD5CD992F6ED,
     0x198E9393920D483A7260BFB731FB5D25F1AA493335A9E71297E485B7AEF312C2),
    (0x12C85EA5DB8C6DEB4AAB71808DCB408FE3D1E7690C43D37B4CE6CC0166FA7DAA,
     0x090689D0585FF075EC9E99AD690C3395BC4B313370B38EF355ACDADCD122975B),
)


# --------------------------------------------------------------------------- Fq2 = Fq[u]/(u^2+1)
def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_neg(a):
    return ((-a[0]) % P, (-a[1]) % P)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, P)
    return (a[0] * d % P, (-a[1] * d) % P)


F2_ZERO = (0, 0)
F2_ONE = (1, 0)
# twist coefficient b' = 3/(9+u)
B_G1 = 3
B_G2 = f2_mul((3, 0), f2_inv((9, 1)))
assert B_G2 == (0x2B149D40CEB8AAAE81BE18991BE06AC3B5B4C5E559DBEFA33267E6DC24A138E5,
                0x009713B03AF0FED4CD2CAFADEED8FDF4A74FA084E52D1852E4A2BD0685C315D2)


class _Fq:
    """Field-op bundle so the same affine group law serves G1 (Fq) and G2 (Fq2)."""
    zero = 0
    one = 1
    add = staticmethod(lambda a, b: (a + b) % P)
    sub = staticmethod(lambda a, b: (a - b) % P)
    mul = staticmethod(lambda a, b: a * b % P)
    neg = staticmethod(lambda a: (-a) % P)
    inv = staticmethod(lambda a: pow(a, -1, P))
    b = B_G1


class _Fq2:
    zero = F2_ZERO
    one = F2_ONE
    add = staticmethod(f2_add)
    sub = staticmethod(f2_sub)
    mul = staticmethod(f2_mul)
    neg = staticmethod(f2_neg)
    inv = staticmethod(f2_inv)
    b = B_G2


# --------------------------------------------------------------------------- group law (affine, None = identity)
def on_curve(F, pt):
    if pt is None:
        return True
    x, y = pt
    return F.mul(y, y) == F.add(F.mul(F.mul(x, x), x), F.b)


def pt_neg(F, a):
    return None if a is None else (a[0], F.neg(a[1]))


def pt_add(F, a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if y1 != y2 or y1 == F.zero:
            return None
        xx = F.mul(x1, x1)
        lam = F.mul(F.add(F.add(xx, xx), xx), F.inv(F.add(y1, y1)))
    else:
        lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
    y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
    return (x3, y3)


def pt_mul(F, k, a):
    """Double-and-add, MSB first.  k is reduced mod r (ark Fr semantics)."""
    k %= R
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = pt_add(F, acc, acc)
        if bit == "1":
            acc = pt_add(F, acc, a)
    return acc


def msm_naive(F, scalars, points):
    """sum s_i * P_i, one double-and-add per term (the definition, nothing clever)."""
    acc = None
    for s, pt in zip(scalars, points):
        acc = pt_add(F, acc, pt_mul(F, s, pt))
    return acc


def g1_add(a, b): return pt_add(_Fq, a, b)
def g1_mul(k, a): return pt_mul(_Fq, k, a)
def g1_msm(s, pts): return msm_naive(_Fq, s, pts)
def g2_add(a, b): return pt_add(_Fq2, a, b)
def g2_mul(k, a): return pt_mul(_Fq2, k, a)
def g2_msm(s, pts): return msm_naive(_Fq2, s, pts)
def g1_on_curve(a): return on_curve(_Fq, a)
def g2_on_curve(a): return on_curve(_Fq2, a)


assert g1_on_curve(G1_GEN) and g2_on_curve(G2_GEN)


# --------------------------------------------------------------------------- encodings (EIP-196/197, provider.rs:201-330)
def g1_to_be(pt) -> bytes:
    if pt is None:
        return b"\x00" * 64
    return pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")


def g1_from_be(b: bytes):
    x, y = int.from_bytes(b[:32], "big"), int.from_bytes(b[32:64], "big")
    return None if x == 0 and y == 0 else (x, y)


def g2_to_be(pt) -> bytes:
    if pt is None:
        return b"\x00" * 128
    (xr, xi), (yr, yi) = pt
    return b"".join(v.to_bytes(32, "big") for v in (xi, xr, yi, yr))


def g2_from_be(b: bytes):
    xi, xr, yi, yr = (int.from_bytes(b[i:i + 32], "big") for i in range(0, 128, 32))
    return None if xi == xr == yi == yr == 0 else ((xr, xi), (yr, yi))


def fr_to_be(v: int) -> bytes:
    return (v % R).to_bytes(32, "big")


# --------------------------------------------------------------------------- NTT (ark-poly Radix2EvaluationDomain semantics)
def root_of_unity(log_n: int, gen_2_28: int = ROOT_2_28) -> int:
    assert 0 <= log_n <= TWO_ADICITY
    return pow(gen_2_28, 1 << (TWO_ADICITY - log_n), R)


def ntt_direct(a, inverse=False, coset=None, gen_2_28: int = ROOT_2_28):
    """O(n^2) evaluation of the definition: out[k] = sum_j a[j] * (h*w^k)^j.
    inverse: out[j] = n^-1 * h^-j * sum_k a[k] w^{-jk}  (ark `coset_ifft`)."""
    n = len(a)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    w = root_of_unity(log_n, gen_2_28)
    if inverse:
        w = pow(w, -1, R)
    out = []
    for k in range(n):
        wk = pow(w, k, R)
        acc, x = 0, 1
        if not inverse and coset is not None:
            wk = wk * coset % R
        for j in range(n):
            acc = (acc + a[j] * x) % R
            x = x * wk % R
        out.append(acc)
    if inverse:
        ninv = pow(n, -1, R)
        out = [v * ninv % R for v in out]
        if coset is not None:
            hinv = pow(coset, -1, R)
            x = 1
            for j in range(n):
                out[j] = out[j] * x % R
                x = x * hinv % R
    return out


def ntt_fast(a, inverse=False, coset=None, gen_2_28: int = ROOT_2_28):
    """Recursive radix-2 (same function as ntt_direct, O(n log n)); used for 2^12 fixtures."""
    n = len(a)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    w = root_of_unity(log_n, gen_2_28)
    a = [v % R for v in a]
    if inverse:
        w = pow(w, -1, R)
    elif coset is not None:
        x = 1
        for j in range(n):
            a[j] = a[j] * x % R
            x = x * coset % R

    def rec(v, w):
        m = len(v)
        if m == 1:
            return v
        e = rec(v[0::2], w * w % R)
        o = rec(v[1::2], w * w % R)
        out = [0] * m
        x = 1
        for k in range(m // 2):
            t = x * o[k] % R
            out[k] = (e[k] + t) % R
            out[k + m // 2] = (e[k] - t) % R
            x = x * w % R
        return out

    out = rec(a, w)
    if inverse:
        ninv = pow(n, -1, R)
        out = [v * ninv % R for v in out]
        if coset is not None:
            hinv = pow(coset, -1, R)
            x = 1
            for j in range(n):
                out[j] = out[j] * x % R
                x = x * hinv % R
    return out


# --------------------------------------------------------------------------- deterministic synthetic inputs
_M64 = (1 << 64) - 1


def _splitmix64(state):
    state = (state + 0x9E3779B97F4A7C15) & _M64
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return state, z ^ (z >> 31)


def rand_fr(seed: int, index: int) -> int:
    """Counter-based generator shared verbatim with the C++ oracle and the CUDA library:
    element `index` = (4 splitmix64 outputs of state seed + 4*index*GOLDEN ... ) mod r.
    Counter-based (not a sequential stream) so any slice can be generated in parallel."""
    limbs = []
    st = (seed + (4 * index) * 0x9E3779B97F4A7C15) & _M64
    for _ in range(4):
        st, z = _splitmix64(st)
        limbs.append(z)
    v = limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | (limbs[3] << 192)
    return v % R


SEED_SCALARS = 0xB2000001
SEED_POINTS = 0xB2000002
SEED_NTT = 0xB2000003


def chain_scalar(seed: int):
    """(k, d): the synthetic base chain is P_i = (k + i*d) * G."""
    return rand_fr(seed, 0) | 1, rand_fr(seed, 1) | 1


def chain_points(F, gen, seed: int, n: int):
    k, d = chain_scalar(seed)
    p0, dd = pt_mul(F, k, gen), pt_mul(F, d, gen)
    out = []
    cur = p0
    for _ in range(n):
        out.append(cur)
        cur = pt_add(F, cur, dd)
    return out


def chain_msm_expected(F, gen, seed: int, scalars):
    """Closed form of MSM over the chain: (sum s_i*(k+i*d) mod r) * G."""
    k, d = chain_scalar(seed)
    acc = 0
    for i, s in enumerate(scalars):
        acc = (acc + s * (k + i * d)) % R
    return pt_mul(F, acc, gen)


# --------------------------------------------------------------------------- optimal ate pairing (slow, textbook)
# Only used to replay the reference's own BN254 known-answer tests -- the 14 `ecpairing` vectors of
# /root/reference/test/tests/levm/precompile_tests.rs:17-140 -- against this file's field/curve arithmetic, which
# is the arithmetic the MSM fixtures are generated with.  Fq12 = Fq[w]/(w^12 - 18 w^6 + 82) (i.e. w^6 = 9 + u),
# G2 points are untwisted into Fq12, the Miller loop uses affine line functions, the final exponentiation is a
# plain power.  ~1-2 s per pairing in CPython: fine for a handful of KATs.
ATE_LOOP_COUNT = 29793968203157093288
LOG_ATE_LOOP_COUNT = 63
_FQ12_MOD = [82, 0, 0, 0, 0, 0, -18, 0, 0, 0, 0, 0]  # w^12 = 18 w^6 - 82


def f12(coeffs):
    c = [int(x) % P for x in coeffs]
    return c + [0] * (12 - len(c))


F12_ONE = f12([1])
F12_ZERO = f12([0])


def f12_add(a, b): return [(x + y) % P for x, y in zip(a, b)]
def f12_sub(a, b): return [(x - y) % P for x, y in zip(a, b)]
def f12_neg(a): return [(-x) % P for x in a]
def f12_scalar(a, k): return [x * k % P for x in a]


def f12_mul(a, b):
    t = [0] * 23
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                t[i + j] += x * y
    for k in range(22, 11, -1):  # w^k = w^(k-12) * (18 w^6 - 82)
        v = t[k]
        if v:
            t[k - 6] += 18 * v
            t[k - 12] -= 82 * v
    return [x % P for x in t[:12]]


def _poly_deg(p):
    d = len(p) - 1
    while d and p[d] == 0:
        d -= 1
    return d


def f12_inv(a):
    """extended Euclid on polynomials over Fq (a != 0)"""
    lm, hm = [1] + [0] * 12, [0] * 13
    low, high = a + [0], [x % P for x in _FQ12_MOD] + [1]
    while _poly_deg(low):
        # r = high / low (polynomial rounded division)
        dega, degb = _poly_deg(high), _poly_deg(low)
        temp, o = high[:], [0] * 13
        inv_lead = pow(low[degb], -1, P)
        for i in range(dega - degb, -1, -1):
            o[i] = (o[i] + temp[degb + i] * inv_lead) % P
            for c in range(degb + 1):
                temp[c + i] = (temp[c + i] - o[i] * low[c]) % P
        r = o[: _poly_deg(o) + 1] + [0] * (13 - _poly_deg(o) - 1)
        nm, new = hm[:], high[:]
        for i in range(13):
            for j in range(13 - i):
                nm[i + j] = (nm[i + j] - lm[i] * r[j]) % P
                new[i + j] = (new[i + j] - low[i] * r[j]) % P
        lm, low, hm, high = nm, new, lm, low
    inv0 = pow(low[0], -1, P)
    return [x * inv0 % P for x in lm[:12]]


def f12_pow(a, e):
    r, b = F12_ONE, a
    while e:
        if e & 1:
            r = f12_mul(r, b)
        b = f12_mul(b, b)
        e >>= 1
    return r


class _Fq12:
    zero = F12_ZERO
    one = F12_ONE
    add = staticmethod(f12_add)
    sub = staticmethod(f12_sub)
    mul = staticmethod(f12_mul)
    neg = staticmethod(f12_neg)
    inv = staticmethod(f12_inv)
    b = f12([3])


def _twist(pt):
    """G2 point over Fq2 (c0 + c1 u) -> the curve y^2 = x^3 + 3 over Fq12 (u = w^6 - 9)."""
    if pt is None:
        return None
    (x0, x1), (y0, y1) = pt
    nx = f12([x0 - 9 * x1, 0, 0, 0, 0, 0, x1])
    ny = f12([y0 - 9 * y1, 0, 0, 0, 0, 0, y1])
    w2, w3 = f12([0, 0, 1]), f12([0, 0, 0, 1])
    return (f12_mul(nx, w2), f12_mul(ny, w3))


def _cast_g1(pt):
    return None if pt is None else (f12([pt[0]]), f12([pt[1]]))


def _linefunc(p1, p2, t):
    x1, y1 = p1
    x2, y2 = p2
    xt, yt = t
    if x1 != x2:
        m = f12_mul(f12_sub(y2, y1), f12_inv(f12_sub(x2, x1)))
        return f12_sub(f12_mul(m, f12_sub(xt, x1)), f12_sub(yt, y1))
    if y1 == y2:
        m = f12_mul(f12_scalar(f12_mul(x1, x1), 3), f12_inv(f12_scalar(y1, 2)))
        return f12_sub(f12_mul(m, f12_sub(xt, x1)), f12_sub(yt, y1))
    return f12_sub(xt, x1)


def miller_loop(q_g2, p_g1):
    """f_{6x+2,Q}(P) * line corrections, before the final exponentiation; 1 if either point is the identity."""
    if q_g2 is None or p_g1 is None:
        return F12_ONE
    Q, Pt = _twist(q_g2), _cast_g1(p_g1)
    R, f = Q, F12_ONE
    for i in range(LOG_ATE_LOOP_COUNT, -1, -1):
        f = f12_mul(f12_mul(f, f), _linefunc(R, R, Pt))
        R = pt_add(_Fq12, R, R)
        if ATE_LOOP_COUNT & (1 << i):
            f = f12_mul(f, _linefunc(R, Q, Pt))
            R = pt_add(_Fq12, R, Q)
    q1 = (f12_pow(Q[0], P), f12_pow(Q[1], P))
    nq2 = (f12_pow(q1[0], P), f12_neg(f12_pow(q1[1], P)))
    f = f12_mul(f, _linefunc(R, q1, Pt))
    R = pt_add(_Fq12, R, q1)
    f = f12_mul(f, _linefunc(R, nq2, Pt))
    return f


def final_exponentiate(f):
    return f12_pow(f, (P ** 12 - 1) // R)


def pairing_check(pairs) -> bool:
    """prod e(P_i, Q_i) == 1 -- the ECPAIRING precompile's answer (provider.rs:277-330)."""
    acc = F12_ONE
    for g1, g2 in pairs:
        acc = f12_mul(acc, miller_loop(g2, g1))
    return final_exponentiate(acc) == F12_ONE
