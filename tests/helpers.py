"""Shared test helpers (tests may use the oracle; the product may not)."""
import numpy as np

import cpu_oracle as orc
import pyref


def to_dev(a: np.ndarray):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64).copy()).cuda()


def to_host(t) -> np.ndarray:
    return t.cpu().numpy().view(np.uint64)


def dev_empty(n_u64: int):
    import torch
    return torch.empty(n_u64, dtype=torch.int64, device="cuda")


def scalars_special(n: int, seed: int = pyref.SEED_SCALARS) -> np.ndarray:
    """random scalars with the edge cases of SURVEY.md section 8c mixed in."""
    s = orc.rand_fr(seed, 0, n)
    special = [0, 1, 2, pyref.R - 1, pyref.R - 2, (1 << 253), (1 << 254) - 1 - pyref.R * 0, 0xFFFF, 1 << 16, (1 << 128) - 1]
    for i, v in enumerate(special):
        if i < n:
            s[(i * 7919) % n] = orc.int_to_limbs(v % pyref.R)
    return s


def chain_kd(seed: int = pyref.SEED_POINTS):
    return pyref.chain_scalar(seed)


def expected_chain_msm_g1(scalars: np.ndarray, k: int, d: int, start: int = 0) -> bytes:
    """closed form: (sum s_i (k + (start+i) d)) * G, evaluated by the CPU oracle."""
    dot = orc.chain_dot(scalars, (k + start * d) % pyref.R, d)
    rc, out = orc.g1_mul_be(pyref.g1_to_be(pyref.G1_GEN), dot.to_bytes(32, "big"))
    return out


def expected_chain_msm_g2(scalars: np.ndarray, k: int, d: int, start: int = 0) -> bytes:
    dot = orc.chain_dot(scalars, (k + start * d) % pyref.R, d)
    rc, out = orc.g2_mul_be(pyref.g2_to_be(pyref.G2_GEN), dot.to_bytes(32, "big"))
    return out
