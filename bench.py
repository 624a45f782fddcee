#!/usr/bin/env python
"""bench.py -- BN254 G1 MSM points/s (+ Fr NTT elements/s) at 2^24 on B200, one process per GPU.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU arm: oracle port timed on the host cores

A "step" is one pass of the hot path over one batch of synthetic input: one 2^24-point BN254 G1 MSM
(BASELINE.json metric, fits one GPU).  `value` = points/s with bases and scalars resident in HBM;
`e2e` = the same MSM through the reference-facing C-ABI call with HOST scalars (pinned) and resident
bases (the proving key is fixed across proofs), host<->device copies inside the timed region.  The Fr NTT
half of the metric (2^24 forward + inverse) is timed in its own K-step loop and reported under "ntt".

Multi-GPU (weak scaling): every rank owns its own 2^24-point shard of one (N * 2^24)-point MSM, reduces
it to one XYZZ partial sum, NCCL all-gathers the 128-byte partials and folds them -- the all-gather + local
fold is the "allreduce of partial sums" (NCCL has no elliptic-curve reduction).  Timing: CUDA events on
the launching stream, barrier + synchronize on both sides, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, torchrun notices) also write to fd 1,
# so keep a private handle on the real stdout for the result line and point fd 1 at stderr for everything else.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)
sys.stdout = sys.stderr


def emit_result(line: dict):
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()


SEED_SCALARS, SEED_POINTS, SEED_NTT = 0xB2000001, 0xB2000002, 0xB2000003
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_msm_sample(log_sample: int, threads: int = 0):
    """(points/s, seconds, cores) of the CPU oracle's Pippenger on a 2^log_sample slice of the same workload."""
    import cpu_oracle as orc
    import pyref
    n = 1 << log_sample
    k, d = pyref.chain_scalar(SEED_POINTS)
    pts = orc.g1_chain(n, k, d)
    s = orc.rand_fr(SEED_SCALARS, 0, n)
    t0 = time.perf_counter()
    out = orc.g1_msm(pts, s, 0, threads)
    dt = time.perf_counter() - t0
    return n / dt, dt, (threads or orc.num_threads()), out, (pts, s, k, d)


def cpu_ntt_sample(log_sample: int, threads: int = 0):
    import cpu_oracle as orc
    n = 1 << log_sample
    a = orc.fr_to_mont(orc.rand_fr(SEED_NTT, 0, n))
    t0 = time.perf_counter()
    orc.fr_ntt(a, log_sample, 0, threads=threads)
    dt = time.perf_counter() - t0
    return n / dt, dt


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the box's host cores.  The reference's own
    (third-party, Rust/Go) MSM cannot be built here (no cargo/go, sources not vendored: SURVEY.md 8c), so this
    is the oracle port (ark-ec 0.5.0 Pippenger rule, OpenMP over windows/chunks)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import cpu_oracle as orc
    log_sample = args.cpu_log_n
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_msm_sample(min(log_sample, 16))
    times, rate = [], 0.0
    for _ in range(args.steps):
        rate, dt, cores, _, _ = cpu_msm_sample(log_sample)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = (1 << log_sample) / (ms / 1e3)
    ntt_rate, ntt_dt = cpu_ntt_sample(min(22, args.log_n))
    line = {
        "impl": "reference", "metric": "bn254_g1_msm_points_per_sec", "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32x8 Montgomery (254-bit modular integer)", "data": "synthetic",
        "config": {"workload": f"2^{args.log_n}-point BN254 G1 MSM per GPU (chain bases P_i=(k+i*d)G, uniform Fr scalars), bases+scalars resident in HBM",
                   "points_per_gpu": 1 << args.log_n, "total_points": args.gpus << args.log_n,
                   "reference_arm": f"rank 0 times a 2^{log_sample}-point slice of that workload per step on the host cores (points/s does not depend on the slice length beyond 2^20)"},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": orc.num_threads(), "kind": "port",
                         "sample": f"2^{log_sample}-point slice of the 2^{args.log_n} workload, Pippenger c={orc.lib().orc_msm_window(1 << log_sample)} (ark-ec rule), all host threads"},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ntt": {"metric": "fr_ntt_elems_per_sec", "value": ntt_rate, "unit": "elements/s", "sample": f"2^{min(22, args.log_n)} forward, {ntt_dt:.3f} s"},
        "gpu_launches": 0,
    }
    emit_result(line)


# ---------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import ethrex_b200 as eb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = eb.Context(local)
    n, log_n = 1 << args.log_n, args.log_n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic workload, generated on the device (deterministic: SURVEY.md section 8d)
    import pyref
    k, d = pyref.chain_scalar(SEED_POINTS)
    start = rank * n
    d_points = torch.empty(8 * n, dtype=torch.int64, device="cuda")
    d_scalars = torch.empty(4 * n, dtype=torch.int64, device="cuda")
    ctx.g1_chain_device(d_points, start, n, k, d)
    ctx.fr_random_device(d_scalars, n, SEED_SCALARS, start)
    d_partial = torch.zeros(16, dtype=torch.int64, device="cuda")
    # the proving key is loaded once: resident bases, expanded to their window multiples (b200zk_bases_precompute)
    handle = ctx.g1_bases_from_device(d_points, n)
    if not args.no_precompute:
        ctx.bases_precompute(handle, args.window)
    elif args.window:
        ctx.set_msm_window(args.window)
    result = {}

    from ethrex_b200.dist import msm_sharded

    def msm_step():
        if world == 1:
            result["out"] = ctx.g1_msm_resident_device(handle, d_scalars, n)
        else:
            result["out"] = msm_sharded(ctx, d_points, d_scalars, n, handle=handle)

    def timed_loop(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        # start the clock sampler BEFORE the barrier: forking nvidia-smi from a process that maps gigabytes of pinned
        # memory takes ~100 ms, and with it after the barrier every other rank's timed region absorbed that wait at the
        # first all_gather (measured at N=8: 61.7 ms "per step" max-over-ranks against 39.9 ms on every rank)
        sampler = ClockSampler(local) if rank == 0 else None
        barrier()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        l0 = ctx.launch_count
        evs[0].record()
        for i in range(steps):
            fn()
            evs[i + 1].record()  # per-step marks inside the one timed region (min / median / max below)
        barrier()
        ms = evs[0].elapsed_time(evs[steps]) / steps
        per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(steps))
        step_stats.update({"min": per[0], "median": per[len(per) // 2], "max": per[-1]})
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), ctx.launch_count - l0, (sampler.stop() if sampler else None)

    step_stats = {}
    ms_step, launches, clocks = timed_loop(msm_step, args.steps, args.warmup)
    msm_step_stats = dict(step_stats)
    value = world * n / (ms_step / 1e3)

    # ---- correctness of what was timed: closed form of the chain MSM (rank 0, outside the timed region)
    verified = None
    if rank == 0 and not args.no_verify:
        import cpu_oracle as orc
        tot = 0
        for r in range(world):
            s = torch.empty(4 * n, dtype=torch.int64, device="cuda")
            ctx.fr_random_device(s, n, SEED_SCALARS, r * n)
            tot = (tot + orc.chain_dot(s.cpu().numpy().view(np.uint64).reshape(n, 4), (k + r * n * d) % R_MOD, d)) % R_MOD
            del s
        _, exp = orc.g1_mul_be(pyref.g1_to_be(pyref.G1_GEN), tot.to_bytes(32, "big"))
        verified = bool(exp == result["out"])
        if not verified:
            raise SystemExit("bench.py: GPU MSM result differs from the oracle's closed form -- refusing to report a number")

    # ---- per-kernel time of the dominant kernel (bucket accumulation), live, CUDA events on the launch stream
    ctx.set_profiling(True)
    acc_ms, phases = [], None
    for _ in range(max(2, min(args.steps, 5))):
        ctx.g1_msm_partial_resident_device(handle, d_scalars, n, d_partial)
        phases = ctx.last_msm_phase_ms()
        acc_ms.append(phases["accumulate"])
    ctx.set_profiling(False)
    acc = sum(acc_ms) / len(acc_ms)
    peak, peak_src = _peaks()
    algo_bytes = n * 96 + 64  # n x (32 B scalar + 64 B affine base) read + 64 B written
    achieved = algo_bytes / (acc / 1e3) / 1e9
    # DRAM traffic of one launch from the committed `ncu --set full` capture of this exact configuration
    # (profiles/r1c_prof_msm_r1b_summary.txt: dram__bytes_read.sum + dram__bytes_write.sum); null for other configs
    traffic = 29.340339e9 + 0.198330e9 if (log_n == 24 and not args.no_precompute and not args.window) else None
    MODMUL_PEAK = 67.7e9  # measured 254-bit Montgomery products/s of the chip (profiles/r1_probe_field_mul.md)
    adds = n * 13 if (not args.no_precompute and not args.window and log_n >= 20) else None
    # multiply instructions of one XYZZ mixed addition, in units of one Montgomery product (136 IMAD-type
    # instructions): 6 products + one two-product/one-reduction mul2 (200) + 2 dedicated squarings (108 each)
    PE = (6 * 136 + 200 + 2 * 108) / 136.0
    roofline = {"bound": "hbm", "kernel": "msm_accumulate<Fq>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel_ms": acc, "phases_ms": phases,
                "binding_roofline": {"bound": "fmaheavy pipe (IMAD.WIDE): 254-bit modular products", "peak_products_per_s": MODMUL_PEAK,
                                     "achieved_products_per_s": (adds * PE / (acc / 1e3)) if adds else None,
                                     "frac": (adds * PE / (acc / 1e3) / MODMUL_PEAK) if adds else None,
                                     "ncu_sm__pipe_fmaheavy_cycles_active_pct": 91.1, "ncu_source": "profiles/r1h_prof_summary.md"},
                "note": "integer-compute-bound kernel (n*13 XYZZ mixed additions of 9.06 product-equivalents: 6 products, one mul2, 2 squarings): the HBM fraction is small by "
                        "construction, see DESIGN.md section 4; kernel_ms is measured in the one-shot schedule (phases are not separable "
                        "in the chunk-pipelined one that `value` runs)"}

    # ---- e2e: C-ABI call with HOST scalars (pinned), resident bases, result read back -- rank-local shard
    e2e = None
    if not args.no_e2e:
        h_scalars = torch.empty(4 * n, dtype=torch.int64).pin_memory()
        h_scalars.copy_(d_scalars)
        d_stage = torch.empty(4 * n, dtype=torch.int64, device="cuda") if world > 1 else None

        def e2e_step():
            if world == 1:
                result["e2e"] = ctx.g1_msm_resident(handle, h_scalars, n)
            else:  # every rank ships its own shard of scalars (pipelined upload), then partial -> all_gather -> fold
                result["e2e"] = msm_sharded(ctx, None, h_scalars, n, handle=handle)
        for _ in range(max(1, args.warmup // 2)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        wall_t = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(wall_t, op=dist.ReduceOp.MAX)
        wall = float(wall_t.item())
        assert result["e2e"] == result["out"]
        e2e = {"value": world * n / wall, "unit": "points/s", "h2d_bytes_per_step": world * n * 32, "d2h_bytes_per_step": world * 64,
               "ms_per_step": wall * 1e3,
               "api": "b200zk_g1_msm_resident (pinned host scalars -> result bytes; bases resident in HBM)" if world == 1 else
                      "pinned host scalars -> ethrex_b200.dist.msm_sharded (b200zk_g1_msm_partial_resident, NCCL all_gather, fold) -> result bytes"}
        del d_stage
        del h_scalars

    # ---- NTT half of the metric: forward + inverse at 2^log_n, resident, K steps each
    ntt = None
    if not args.no_ntt:
        d_ntt = torch.empty(4 * n, dtype=torch.int64, device="cuda")
        ctx.fr_random_device(d_ntt, n, SEED_NTT, start, eb.SCALARS_MONT)
        ref = d_ntt.clone()
        fwd_ms, fl, _ = timed_loop(lambda: ctx.fr_ntt_device(d_ntt, log_n, 0), args.steps, args.warmup)
        d_ntt.copy_(ref)
        ctx.fr_ntt_device(d_ntt, log_n, 0)
        inv_ms, _, _ = timed_loop(lambda: ctx.fr_ntt_device(d_ntt, log_n, eb.NTT_INVERSE), args.steps, args.warmup)
        # round trip check on fresh data
        d_ntt.copy_(ref)
        ctx.fr_ntt_device(d_ntt, log_n, 0)
        ctx.fr_ntt_device(d_ntt, log_n, eb.NTT_INVERSE)
        ok = bool(torch.equal(d_ntt, ref))
        if not ok:
            raise SystemExit("bench.py: iNTT(NTT(a)) != a")
        ntt_bytes = 64 * n
        ntt = {"metric": "fr_ntt_elems_per_sec", "value": world * n / (fwd_ms / 1e3), "unit": "elements/s", "forward_ms": fwd_ms, "inverse_ms": inv_ms,
               "inverse_value": world * n / (inv_ms / 1e3), "roundtrip_ok": ok, "launches_per_transform": fl // max(1, args.steps),
               "roofline": {"bound": "hbm", "achieved": ntt_bytes / (fwd_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": ntt_bytes / (fwd_ms / 1e3) / 1e9 / peak, "traffic": 3.061e9 if log_n == 24 else None,
                            "binding_roofline": {"bound": "fmaheavy pipe: ~12 modular products per element", "frac": 12 * n / (fwd_ms / 1e3) / 67.7e9},
                            "note": "whole transform (3 passes at 2^24); algorithmic bytes = 64*n; traffic = sum of the three passes' "
                                    "dram bytes from profiles/r1c_prof_ntt_r1b_summary.txt"}}
        # end to end through the host-buffer C-ABI call: pinned host buffer in, transformed in place, copies included
        if not args.no_e2e and world == 1:
            h_ntt = torch.empty(4 * n, dtype=torch.int64).pin_memory()
            h_ntt.copy_(ref)
            ctx.fr_ntt(h_ntt, log_n, 0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                ctx.fr_ntt(h_ntt, log_n, 0)
            wall = (time.perf_counter() - t0) / args.steps
            ntt["e2e"] = {"value": n / wall, "unit": "elements/s", "ms_per_step": wall * 1e3, "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 32 * n,
                          "api": "b200zk_fr_ntt (pinned host buffer, in place): PCIe-bound, 2 x 512 MiB per transform"}
            del h_ntt
        del d_ntt, ref

    # ---- config #5: Groth16-shaped wrap (7 NTT + quotient + 4 G1 MSM + 1 G2 MSM) through B200Backend.prove, N=1
    proof = None
    if not args.no_proof:
        from ethrex_b200.backend import B200Backend, ProofFormat
        from ethrex_b200.groth16 import SyntheticWrapCircuit
        torch.cuda.empty_cache()
        t0 = time.perf_counter()
        circuit = SyntheticWrapCircuit(ctx, args.proof_log_n, precompute=True, rank=rank, world=world)
        ctx.synchronize()
        setup_s = time.perf_counter() - t0
        backend = B200Backend(ctx, circuit)
        backend.prove({"batch": 0})  # warm-up (workspaces, twiddles)
        times = []
        digests = []
        for i in range(max(2, min(args.steps, 5))):
            barrier()
            pr, dt = backend.prove_timed({"batch": i + 1}, ProofFormat.GROTH16)
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            times.append(float(tt.item()))
            digests.append(pr.proof.hex()[:16])
        circuit.close()
        proof = {"metric": "groth16_wrap_prove_wall_ms", "value": 1e3 * sorted(times)[len(times) // 2], "unit": "ms", "higher_is_better": False,
                 "domain_log2": args.proof_log_n, "proving_key_setup_s": setup_s, "n_gpus": world, "proof_prefix": digests[-1],
                 "multi_gpu": "proving-key columns point-split across ranks, 5 x (partial MSM, NCCL all_gather, fold); NTTs replicated" if world > 1 else "single GPU",
                 "work": "3 iNTT + 3 coset NTT + quotient + 1 coset iNTT, 4 G1 MSM + 1 G2 MSM (synthetic R1CS, chain proving key, no blinding; STARK stage excluded)"}

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only), bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import cpu_oracle as orc
        rate, dt, cores, _, _ = cpu_msm_sample(args.cpu_log_n)
        ntt_rate, ntt_dt = cpu_ntt_sample(min(22, log_n))
        cpu = {"value": rate, "unit": "points/s", "cores": cores, "kind": "port",
               "sample": f"2^{args.cpu_log_n}-point slice of the same workload, {dt:.2f} s, Pippenger c={orc.lib().orc_msm_window(1 << args.cpu_log_n)} (ark-ec 0.5.0 rule)",
               "ntt": {"value": ntt_rate, "unit": "elements/s", "sample": f"2^{min(22, log_n)} forward NTT, {ntt_dt:.2f} s"}}

    if rank == 0:
        line = {
            "metric": "bn254_g1_msm_points_per_sec", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "step_ms": msm_step_stats, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x8 Montgomery (254-bit modular integer)", "data": "synthetic",
            "config": {"workload": f"2^{log_n}-point BN254 G1 MSM per GPU (chain bases P_i=(k+i*d)G, uniform Fr scalars), bases+scalars resident in HBM",
                       "points_per_gpu": n, "total_points": world * n, "l2": "inputs (1.6 GB/GPU) larger than L2; no flush needed",
                       "multi_gpu": "point-split, NCCL all_gather of 128-B XYZZ partials + local fold" if world > 1 else "single GPU"},
            "verified_vs_oracle": verified, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "e2e": e2e, "ntt": ntt, "proof": proof, "cpu_baseline": cpu,
        }
        emit_result(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-n", type=int, default=24)
    ap.add_argument("--cpu-log-n", type=int, default=22, help="size of the CPU-arm sample (2^k points)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ntt", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-proof", action="store_true")
    ap.add_argument("--proof-log-n", type=int, default=22, help="domain size of the synthetic Groth16 wrap (config #5)")
    ap.add_argument("--no-precompute", action="store_true", help="plain resident bases (no 2^(cw) P_i table)")
    ap.add_argument("--window", type=int, default=0, help="force the MSM window bits (0 = automatic)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
