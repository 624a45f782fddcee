"""How does the CPU oracle scale on this box?  (Picks the thread count for bench.py's cpu_baseline.)"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cpu_oracle as orc, pyref
k, d = pyref.chain_scalar(pyref.SEED_POINTS)
print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "omp", orc.num_threads())
try:
    print("cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e:
    print("cpu.max n/a", e)
for lg in (20, 22):
    n = 1 << lg
    pts = orc.g1_chain(n, k, d)
    s = orc.rand_fr(pyref.SEED_SCALARS, 0, n)
    for thr in (1, 8, 16, 32, 64, 128):
        if thr == 1 and lg > 20:
            continue
        t0 = time.perf_counter(); orc.g1_msm(pts, s, 0, thr); dt = time.perf_counter() - t0
        print(json.dumps({"msm_log_n": lg, "threads": thr, "s": round(dt, 3), "mpts_per_s": round(n / dt / 1e6, 3)}), flush=True)
a = orc.fr_to_mont(orc.rand_fr(pyref.SEED_NTT, 0, 1 << 22))
for thr in (1, 8, 32, 64, 128):
    t0 = time.perf_counter(); orc.fr_ntt(a, 22, 0, threads=thr); dt = time.perf_counter() - t0
    print(json.dumps({"ntt_log_n": 22, "threads": thr, "s": round(dt, 3), "melem_per_s": round((1 << 22) / dt / 1e6, 2)}), flush=True)
