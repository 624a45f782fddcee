"""profiles/r2_scaling.md from the committed bench lines (N = 1, 2, 4, 8): absolute throughput per N, the efficiency the
driver would compute, the strong-scaling (config 4) and proof (config 5) blocks.   python tools/scaling_table.py > profiles/r2_scaling.md"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
files = {1: "r2_bench.json", 2: "r2_bench_n2.json", 4: "r2_bench_n4.json", 8: "r2_bench_n8.json"}
rows = {n: json.load(open(os.path.join(ROOT, "profiles", f))) for n, f in files.items() if os.path.exists(os.path.join(ROOT, "profiles", f))}
base = rows[1]
print("# 1 → 8 GPUs on the final tree (`bench.py`, one rank per GPU under torchrun, max over ranks, CUDA events)\n")
print("Weak scaling (the headline: every rank runs its own 2²⁴-point G1 MSM; no data-path collective), strong scaling of ONE 2²⁴-point")
print("MSM point-split N ways (config 4: one NCCL all_gather of XYZZ partials + local fold) and the Groth16-shaped prove at domain 2²⁴")
print("(config 5: dealt NTTs, three broadcasts, one all_gather of 768-byte blocks). One-GPU times of the strong block are measured in the same run.\n")
print("| N | G1 MSM points/s (weak) | ms per step | vs N × one GPU | e2e points/s (host scalars) | NTT elements/s (weak) | accumulate: fraction of the live product ceiling |")
print("|---|---|---|---|---|---|---|")
for n, d in sorted(rows.items()):
    br = d["roofline"]["binding_roofline"]
    print(f"| {n} | {d['value']:.3e} | {d['ms_per_step']:.2f} | {100 * d['value'] / (n * base['value']):.1f} % | {d['e2e']['value']:.3e} | {d['ntt']['value']:.3e} | {100 * br['frac']:.1f} % |")
print("\n| N | ONE 2²⁴ G1 MSM (ms) | speed-up | ONE 2²⁴ G2 MSM (ms) | speed-up | both equal the one-GPU bytes and the closed form | prove at 2²⁴ (ms) | speed-up | proof bytes (prefix) |")
print("|---|---|---|---|---|---|---|---|---|")
p1 = base["proof"]["value"]
print(f"| 1 | {base['ms_per_step']:.2f} | 1 | {base['g2']['ms_per_step']:.2f} | 1 | closed form: yes | {p1:.1f} | 1 | `{base['proof']['proof_prefix']}` |")
for n, d in sorted(rows.items()):
    s = d.get("strong")
    if not s:
        continue
    ok = all(s[k] for k in ("g1_equals_one_gpu_result", "g1_verified_vs_oracle", "g2_equals_one_gpu_result", "g2_verified_vs_oracle"))
    print(f"| {n} | {s['g1_ms']:.2f} | ×{s['g1_speedup_vs_one_gpu']:.2f} | {s['g2_ms']:.2f} | ×{s['g2_speedup_vs_one_gpu']:.2f} | {'yes' if ok else 'NO'} | {d['proof']['value']:.1f} | ×{p1 / d['proof']['value']:.2f} | `{d['proof']['proof_prefix']}` |")
print("\nAt 2²¹ points per rank (N = 8) the per-MSM fixed costs — the sort's floor, the 1 ms bucket reduction over 2¹⁹ buckets, ~25 launches, the")
print("gather and the fold — are a quarter of the step; the prove also pays its three NTT broadcasts. The HBM-roofline fraction of the dominant")
print("kernel is in each line's `roofline` (≈0.8 %: the path is bound by the multiplier pipe, DESIGN.md §4).")
