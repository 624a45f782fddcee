"""TEST INFRASTRUCTURE: a real (small) Groth16 instance over BN254, built with the CPU oracle.

A random R1CS with n = 2^log_n constraints (constraint j defines a fresh variable as the product of two random linear
combinations of earlier ones), a trusted setup with KNOWN toxic waste, the proving-key query columns a prover needs,
and the verifier's equation.  Because the toxic waste is known, every proof element is also available "in the
exponent" as one scalar multiplication -- which gives the GPU pipeline (ethrex_b200.groth16.Groth16Prover: 7 NTTs,
the quotient, 5 MSMs) a bit-exact expected proof, on top of the pairing check that accepts it.

Groth16 (no blinding, r = s = 0):   A = alpha + sum z_i u_i      B = beta + sum z_i v_i
                                    C = (sum_{private} z_i (beta u_i + alpha v_i + w_i) + h(tau) Z(tau)) / delta
verifier:                           e(A, B) = e(alpha, beta) * e(sum_{public} z_i IC_i, gamma) * e(C, delta)
with u_i = A_i(tau), v_i = B_i(tau), w_i = C_i(tau) the QAP polynomials of variable i at tau.
"""
import random

import cpu_oracle as orc
import pyref as o

R = o.R
N_PUBLIC = 2  # z_0 = 1 and one public input


def _g1(k: int) -> bytes:
    return orc.g1_mul_be(o.g1_to_be(o.G1_GEN), (k % R).to_bytes(32, "big"))[1]


def _g2(k: int) -> bytes:
    return orc.g2_mul_be(o.g2_to_be(o.G2_GEN), (k % R).to_bytes(32, "big"))[1]


class ToyGroth16:
    def __init__(self, log_n: int, seed: int = 0xB2001616):
        rng = random.Random(seed)
        self.log_n, self.n = log_n, 1 << log_n
        n = self.n
        self.m = N_PUBLIC + n
        self.A, self.B = [], []
        for j in range(n):
            avail = N_PUBLIC + j
            self.A.append({rng.randrange(avail): rng.randrange(1, R) for _ in range(3)})
            self.B.append({rng.randrange(avail): rng.randrange(1, R) for _ in range(3)})
        # C row j = the fresh variable N_PUBLIC + j
        self.tau, self.alpha, self.beta, self.gamma, self.delta = (rng.randrange(2, R) for _ in range(5))
        tau = self.tau
        omega = o.root_of_unity(log_n)
        self.Zt = (pow(tau, n, R) - 1) % R
        ninv = pow(n, -1, R)
        lag = []
        wj = 1
        for j in range(n):  # L_j(tau) = Z(tau)/n * w^j / (tau - w^j)
            lag.append(self.Zt * ninv % R * wj % R * pow((tau - wj) % R, -1, R) % R)
            wj = wj * omega % R
        self.u, self.v, self.w = [0] * self.m, [0] * self.m, [0] * self.m
        for j in range(n):
            for i, cf in self.A[j].items():
                self.u[i] = (self.u[i] + cf * lag[j]) % R
            for i, cf in self.B[j].items():
                self.v[i] = (self.v[i] + cf * lag[j]) % R
            self.w[N_PUBLIC + j] = lag[j]
        dinv, ginv = pow(self.delta, -1, R), pow(self.gamma, -1, R)
        self.k = [(self.beta * self.u[i] + self.alpha * self.v[i] + self.w[i]) % R for i in range(self.m)]
        # ---- proving key columns (alpha / beta ride on variable 0, whose value is 1) ----
        self.a_g1 = b"".join(_g1(self.u[i] + (self.alpha if i == 0 else 0)) for i in range(self.m))
        self.b_g1 = b"".join(_g1(self.v[i] + (self.beta if i == 0 else 0)) for i in range(self.m))
        self.b_g2 = b"".join(_g2(self.v[i] + (self.beta if i == 0 else 0)) for i in range(self.m))
        self.l_g1 = b"".join(_g1(self.k[i] * dinv) for i in range(N_PUBLIC, self.m))
        self.h_g1 = b"".join(_g1(pow(tau, k, R) * self.Zt % R * dinv) for k in range(n - 1))
        # ---- verifying key ----
        self.vk_alpha_g1, self.vk_beta_g2 = _g1(self.alpha), _g2(self.beta)
        self.vk_gamma_g2, self.vk_delta_g2 = _g2(self.gamma), _g2(self.delta)
        self.ic = [self.k[i] * ginv % R for i in range(N_PUBLIC)]  # kept as scalars: IC(pub) is one scalar multiplication

    # ---- witness ----
    def assign(self, x: int):
        z = [1, x % R]
        for j in range(self.n):
            left = sum(cf * z[i] for i, cf in self.A[j].items()) % R
            right = sum(cf * z[i] for i, cf in self.B[j].items()) % R
            z.append(left * right % R)
        return z

    def evaluations(self, z):
        a = [sum(cf * z[i] for i, cf in self.A[j].items()) % R for j in range(self.n)]
        b = [sum(cf * z[i] for i, cf in self.B[j].items()) % R for j in range(self.n)]
        c = [z[N_PUBLIC + j] for j in range(self.n)]
        return a, b, c

    # ---- the proof "in the exponent" (what the MSM/NTT pipeline must reproduce bit for bit) ----
    def expected_proof(self, z) -> bytes:
        n = self.n
        a_s = (self.alpha + sum(zi * ui for zi, ui in zip(z, self.u))) % R
        b_s = (self.beta + sum(zi * vi for zi, vi in zip(z, self.v))) % R
        at = sum(zi * ui for zi, ui in zip(z, self.u)) % R
        bt = sum(zi * vi for zi, vi in zip(z, self.v)) % R
        ct = sum(zi * wi for zi, wi in zip(z, self.w)) % R
        hz = (at * bt - ct) % R  # = h(tau) Z(tau): the R1CS holds on the whole domain
        c_s = (sum(z[i] * self.k[i] for i in range(N_PUBLIC, self.m)) + hz) % R * pow(self.delta, -1, R) % R
        assert (a_s * b_s - self.alpha * self.beta - self.gamma * sum(z[i] * self.ic[i] for i in range(N_PUBLIC)) - self.delta * c_s) % R == 0
        del n
        return _g1(a_s) + _g2(b_s) + _g1(c_s)

    # ---- verifier: the ecpairing calldata whose check must be 1 ----
    def verifier_calldata(self, proof: bytes, public_x: int) -> bytes:
        A, B, C = proof[:64], proof[64:192], proof[192:256]
        ax, ay = int.from_bytes(A[:32], "big"), int.from_bytes(A[32:], "big")
        neg_a = A[:32] + ((o.P - ay) % o.P).to_bytes(32, "big") if (ax or ay) else A
        ic = _g1(self.ic[0] + self.ic[1] * (public_x % R))
        return neg_a + B + self.vk_alpha_g1 + self.vk_beta_g2 + ic + self.vk_gamma_g2 + C + self.vk_delta_g2


def write_c_fixture(path: str, log_n: int = 4, public_x: int = 0x1234567):
    """Binary fixture for examples/c_abi_demo.c section 6 (little-endian host):
    "G16F" | u32 log_n | u32 m | u32 n_public | A_g1 (m x 64, EIP-196) | B_g1 (m x 64) | B_g2 (m x 128, EIP-197) |
    L_g1 ((m - n_public) x 64) | H_g1 ((2^log_n - 1) x 64) | witness (m x 32, canonical LE) |
    (A z), (B z), (C z) on the domain (2^log_n x 32 each, Montgomery LE) | expected proof (256)."""
    import struct
    toy = ToyGroth16(log_n)
    z = toy.assign(public_x)
    a, b, c = toy.evaluations(z)
    r_mont = (1 << 256) % R
    le = lambda vals, mont: b"".join(int(v % R * (r_mont if mont else 1) % R).to_bytes(32, "little") for v in vals)  # noqa: E731
    blob = b"G16F" + struct.pack("<III", log_n, toy.m, N_PUBLIC) + toy.a_g1 + toy.b_g1 + toy.b_g2 + toy.l_g1 + toy.h_g1
    blob += le(z, False) + le(a, True) + le(b, True) + le(c, True) + toy.expected_proof(z)
    with open(path, "wb") as f:
        f.write(blob)
    return toy
