"""Latency of one EIP-4844 blob commitment through the C ABI (b200zk_kzg_blob_to_commitment: 128 KiB host blob in, 48-byte
compressed commitment out; a 4096-point BLS12-381 G1 MSM over the resident Lagrange setup), with and without the setup's window
table.  The setup is synthetic (L_i(tau) * G for a known tau, as in tests/test_gpu_bls.py).  JSON lines."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402

import bls_ref as bls  # noqa: E402
import ethrex_b200 as eb  # noqa: E402

TAU = 0x5A3C91E7B2D4F60819ACBD3E57F1024689BDF0135792468ACE0FDB9753102468 % bls.R


def main():
    ctx = eb.Context(0)
    lag = bls.lagrange_setup_scalars(TAU)
    setup = b"".join(bls.compress(p) for p in bls.generator_multiples(lag))
    rng = np.random.default_rng(4844)
    vals = [int.from_bytes(rng.bytes(32), "big") % bls.R for _ in range(4096)]
    blob = b"".join(v.to_bytes(32, "big") for v in vals)
    expect = bls.compress(bls.mul(sum(v * l for v, l in zip(vals, lag)) % bls.R, bls.G1))
    for table in (False, True):
        h = ctx.bls12_381_g1_bases_upload(setup, 4096)
        if table:
            ctx.bases_precompute(h, 0)
        assert ctx.kzg_blob_to_commitment(h, blob) == [expect]
        for _ in range(5):
            ctx.kzg_blob_to_commitment(h, blob)
        reps = 50
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.kzg_blob_to_commitment(h, blob)
        ms = (time.perf_counter() - t0) / reps * 1e3
        six = blob * 6  # a full block's worth of blobs in one call
        assert ctx.kzg_blob_to_commitment(h, six) == [expect] * 6
        t0 = time.perf_counter()
        for _ in range(10):
            ctx.kzg_blob_to_commitment(h, six)
        ms6 = (time.perf_counter() - t0) / 10 * 1e3
        print(json.dumps({"probe": "kzg_blob_to_commitment", "window_table": table, "ms_per_blob": ms, "ms_per_call_of_6_blobs": ms6, "points": 4096, "verified_vs_closed_form": True}), flush=True)
        ctx.bases_free(h)
    ctx.close()


if __name__ == "__main__":
    main()
