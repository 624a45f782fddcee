"""The workload `ncu` is pointed at to produce profiles/r2_ncu_kernels.json: exactly ONE 2^24-point G1 MSM, ONE 2^24-point
G2 MSM (both over resident window tables, the bench's configuration) and ONE forward 2^24 NTT, after one warm-up of each
at the same size (workspaces, twiddles).  Never a bench number: ncu serialises and replays every launch.

    ncu --set full --clock-control none -k regex:"msm_accumulate|ntt_pass|msm_hist|msm_scatter|msm_sort" --csv --page raw \
        --log-file gpurun_out/r2_ncu_full.csv python tools/profile_workload.py
    python tools/ncu_to_profile_json.py gpurun_out/r2_ncu_full.csv      # here, no GPU needed
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import torch  # noqa: E402

import ethrex_b200 as eb  # noqa: E402
import pyref  # noqa: E402


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    which = sys.argv[2] if len(sys.argv) > 2 else "g1,g2,ntt"
    n = 1 << log_n
    torch.cuda.set_device(0)
    ctx = eb.Context(0)
    k, d = pyref.chain_scalar(0xB2000002)
    sc = torch.empty(4 * n, dtype=torch.int64, device="cuda")
    ctx.fr_random_device(sc, n, 0xB2000001, 0)
    for g2 in (False, True):
        if ("g2" if g2 else "g1") not in which:
            continue
        pts = torch.empty((16 if g2 else 8) * n, dtype=torch.int64, device="cuda")
        (ctx.g2_chain_device if g2 else ctx.g1_chain_device)(pts, 0, n, k, d)
        h = (ctx.g2_bases_from_device if g2 else ctx.g1_bases_from_device)(pts, n)
        del pts
        torch.cuda.empty_cache()
        ctx.bases_precompute(h, int(os.environ.get("B200ZK_PROFILE_WINDOW", "0")))
        fn = ctx.g2_msm_resident_device if g2 else ctx.g1_msm_resident_device
        fn(h, sc, n)  # warm-up (profiled too: tools/ncu_to_profile_json.py keeps the LAST launch of each kernel)
        fn(h, sc, n)
        ctx.bases_free(h)
        torch.cuda.empty_cache()
    if "ntt" in which:
        a = torch.empty(4 * n, dtype=torch.int64, device="cuda")
        ctx.fr_random_device(a, n, 0xB2000003, 0, eb.SCALARS_MONT)
        ctx.fr_ntt_device(a, log_n, 0)
        ctx.fr_ntt_device(a, log_n, 0)
    ctx.synchronize()
    ctx.close()
    print("profile workload done")


if __name__ == "__main__":
    main()
