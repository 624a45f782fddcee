"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the keys the driver reads,
uses every host CPU even when torchrun's OMP_NUM_THREADS=1 is in the environment, and non-zero ranks stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-log-n", "14"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_line_and_thread_count():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_oracle as orc
    out = _run({"OMP_NUM_THREADS": "1"})
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "bn254_g1_msm_points_per_sec" and d["unit"] == "points/s"
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["cores"] == orc.effective_cpus()  # not 1: OMP_NUM_THREADS=1 came from the launcher, not the user
    assert d["config"]["workload"].startswith("2^24-point BN254 G1 MSM per GPU")


def test_reference_arm_is_silent_on_other_ranks():
    assert _run({"RANK": "1", "WORLD_SIZE": "2"}).strip() == ""


def test_importing_bench_leaves_stdout_alone_and_counts_ntt_products(capsys):
    """Tools and tests import bench for source_hash() / the product count: the import must not redirect the importer's
    stdout (only main() claims it), and the NTT product model must follow csrc/ntt.cu's schedule: three 8-stage passes at
    2^24 = 3 x 3.0 stage products + one inter-pass product in passes 2 and 3 with the last pass's direct table, two there
    without it; 2^26 exceeds one lookup in its second pass."""
    sys.path.insert(0, ROOT)
    import bench
    print("still here")
    assert capsys.readouterr().out == "still here\n"
    assert abs(bench.ntt_products_per_element(24) - 11.012) < 1e-2
    assert abs(bench.ntt_products_per_element(24, 0) - 12.012) < 1e-2
    assert abs(bench.ntt_products_per_element(16) - 7.008) < 1e-2           # (8, 8): one lookup, no table needed
    assert bench.ntt_products_per_element(26) > bench.ntt_products_per_element(24) + 1.9
    assert len(bench.source_hash()) == 16
