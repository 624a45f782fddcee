"""TEST INFRASTRUCTURE -- CPU restatement, never imported by the product.

Second, independent statement of the BN254 pairing check, written the way the CUDA kernel
(ethrex_b200/csrc/pairing.cu) is: the tower Fq2 = Fq[u]/(u^2+1), Fq6 = Fq2[v]/(v^3 - xi), Fq12 = Fq6[w]/(w^2 - v)
with xi = 9 + u, the PLAIN ate pairing f_{T,Q}(P) with T = t - 1 = 6x^2 (no Frobenius correction lines), affine
line functions on the twist, and the final exponentiation conj(f)/f followed by one generic power (p^6+1)/r.

pyref.pairing_check (py_ecc-style degree-12 polynomial field, optimal-ate loop 6x+2 with the two Frobenius
lines) is the other statement; both are pinned by the reference's 14 ecpairing vectors
(/root/reference/test/tests/levm/precompile_tests.rs:17-140, replayed from tests/golden/pairing_kats.json).
A pairing CHECK only asks whether the product is one, which every non-degenerate bilinear pairing on the
same groups answers identically -- the semantics of `Crypto::bn254_pairing_check`
(/root/reference/crates/common/crypto/provider.rs:277-330, ark `Bn254::multi_pairing(..) == one`).
"""
from pyref import P, R, f2_add, f2_sub, f2_mul, f2_inv, f2_neg

X = 4965661367192848881            # BN parameter
ATE_T = 6 * X * X                  # t - 1, 127 bits
FINAL_EXP = (P ** 6 + 1) // R      # after the easy part f^(p^6-1)
assert (P ** 6 + 1) % R == 0 and (P + 1 - (ATE_T + 1)) == R

XI = (9, 1)
F2_ZERO, F2_ONE = (0, 0), (1, 0)


def f2_mul_xi(a):
    return ((9 * a[0] - a[1]) % P, (9 * a[1] + a[0]) % P)


# ---- Fq6: (c0, c1, c2) = c0 + c1 v + c2 v^2 ---------------------------------------------------------------
F6_ZERO, F6_ONE = (F2_ZERO, F2_ZERO, F2_ZERO), (F2_ONE, F2_ZERO, F2_ZERO)


def f6_add(a, b): return tuple(f2_add(x, y) for x, y in zip(a, b))
def f6_sub(a, b): return tuple(f2_sub(x, y) for x, y in zip(a, b))
def f6_neg(a): return tuple(f2_neg(x) for x in a)


def f6_mul(a, b):
    t0, t1, t2 = f2_mul(a[0], b[0]), f2_mul(a[1], b[1]), f2_mul(a[2], b[2])
    c0 = f2_add(t0, f2_mul_xi(f2_sub(f2_sub(f2_mul(f2_add(a[1], a[2]), f2_add(b[1], b[2])), t1), t2)))
    c1 = f2_add(f2_sub(f2_sub(f2_mul(f2_add(a[0], a[1]), f2_add(b[0], b[1])), t0), t1), f2_mul_xi(t2))
    c2 = f2_add(f2_sub(f2_sub(f2_mul(f2_add(a[0], a[2]), f2_add(b[0], b[2])), t0), t2), t1)
    return (c0, c1, c2)


def f6_mul_v(a):
    return (f2_mul_xi(a[2]), a[0], a[1])


def f6_inv(a):
    A = f2_sub(f2_mul(a[0], a[0]), f2_mul_xi(f2_mul(a[1], a[2])))
    B = f2_sub(f2_mul_xi(f2_mul(a[2], a[2])), f2_mul(a[0], a[1]))
    C = f2_sub(f2_mul(a[1], a[1]), f2_mul(a[0], a[2]))
    F = f2_add(f2_mul(a[0], A), f2_mul_xi(f2_add(f2_mul(a[2], B), f2_mul(a[1], C))))
    Fi = f2_inv(F)
    return (f2_mul(A, Fi), f2_mul(B, Fi), f2_mul(C, Fi))


# ---- Fq12: (c0, c1) = c0 + c1 w ---------------------------------------------------------------------------
F12_ONE = (F6_ONE, F6_ZERO)


def f12_mul(a, b):
    t0, t1 = f6_mul(a[0], b[0]), f6_mul(a[1], b[1])
    c1 = f6_sub(f6_sub(f6_mul(f6_add(a[0], a[1]), f6_add(b[0], b[1])), t0), t1)
    return (f6_add(t0, f6_mul_v(t1)), c1)


def f12_conj(a): return (a[0], f6_neg(a[1]))


def f12_inv(a):
    t = f6_inv(f6_sub(f6_mul(a[0], a[0]), f6_mul_v(f6_mul(a[1], a[1]))))
    return (f6_mul(a[0], t), f6_neg(f6_mul(a[1], t)))


def f12_pow(a, e):
    acc = F12_ONE
    for i in reversed(range(e.bit_length())):
        acc = f12_mul(acc, acc)
        if (e >> i) & 1:
            acc = f12_mul(acc, a)
    return acc


# ---- lines on the twist (x', y') -> (x' w^2, y' w^3), w^6 = xi ---------------------------------------------
def _line(lam, xt, yt, p):
    """l(P) = yP - (lam xP) w + (lam xT - yT) w^3, w^3 = v w."""
    xp, yp = p
    c0 = ((yp % P, 0), F2_ZERO, F2_ZERO)
    c1 = (f2_neg((lam[0] * xp % P, lam[1] * xp % P)), f2_sub(f2_mul(lam, xt), yt), F2_ZERO)
    return (c0, c1)


def miller_ate(q, p):
    """f_{T,Q}(P); q = ((x_re, x_im), (y_re, y_im)) affine on the twist, p = (x, y) affine G1; neither at infinity."""
    f = F12_ONE
    rx, ry = q
    for i in reversed(range(ATE_T.bit_length() - 1)):
        lam = f2_mul(f2_mul((3, 0), f2_mul(rx, rx)), f2_inv(f2_add(ry, ry)))
        f = f12_mul(f12_mul(f, f), _line(lam, rx, ry, p))
        nx = f2_sub(f2_sub(f2_mul(lam, lam), rx), rx)
        ry = f2_sub(f2_mul(lam, f2_sub(rx, nx)), ry)
        rx = nx
        if (ATE_T >> i) & 1:
            lam = f2_mul(f2_sub(ry, q[1]), f2_inv(f2_sub(rx, q[0])))
            f = f12_mul(f, _line(lam, rx, ry, p))
            nx = f2_sub(f2_sub(f2_mul(lam, lam), rx), q[0])
            ry = f2_sub(f2_mul(lam, f2_sub(rx, nx)), ry)
            rx = nx
    return f


def final_exponentiate(f):
    return f12_pow(f12_mul(f12_conj(f), f12_inv(f)), FINAL_EXP)


def pairing_check(pairs) -> bool:
    """pairs: [(g1 affine or None, g2 affine or None)] with pyref's point conventions (None = identity)."""
    f = F12_ONE
    for g1, g2 in pairs:
        if g1 is None or g2 is None:
            continue
        f = f12_mul(f, miller_ate(g2, g1))
    return final_exponentiate(f) == F12_ONE
