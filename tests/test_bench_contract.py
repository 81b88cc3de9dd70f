"""CPU checks of bench.py's contract: workloads, the reference arm's JSON line, rank behaviour,
and that the product arm fails loudly (no CPU fallback) when no GPU is visible."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_workloads_match_baseline_configs():
    import bench
    assert bench.workload("weak16384", 1)[:5] == (16384, 16384, 1, 1, "weak")
    assert bench.workload("weak16384", 8)[:5] == (65536, 32768, 4, 2, "weak")       # 2-D decomposition
    assert bench.workload("strong32768", 8)[:5] == (32768, 32768, 8, 1, "strong")
    assert bench.workload("cavity4096", 1)[:4] == (4096, 4096, 1, 1)
    with pytest.raises(SystemExit):
        bench.workload("weak16384", 3)
    # Re = 1000: omega = 2 Re / (6 L u0 + Re)  (slidingLid.py:28)
    assert abs(bench.omega_for_re(4096) - 0.5784359093012493) < 1e-15
    assert bench.BYTES_PER_CELL == 144


def run_bench(args, env_extra=None):
    env = dict(os.environ, LBM_REF_BLOCK="96")
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = run_bench(["--impl", "reference", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MLUPS" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["dtype"] == "f64" and d["vs_baseline"] is None and "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = run_bench(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], {"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_fails_loudly_without_gpu():
    from latticeboltzmann_b200 import _lib
    if _lib.load().lb_device_count() > 0:
        pytest.skip("a GPU is visible")
    r = run_bench(["--steps", "1", "--warmup", "3", "--no-e2e", "--no-cpu-baseline"])
    assert r.returncode != 0
    assert r.stdout.strip() == ""          # no fabricated line


def test_roofline_traffic_is_reproducible_from_the_committed_ncu_logs(tmp_path):
    """`roofline.traffic` in bench.py's line comes from profiles/r02_kernel_dram.json; that file must be exactly what
    tools/ncu_dram_summary.py derives from the committed ncu launch lists of the bench command."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    committed = json.load(open(os.path.join(root, "profiles", "r02_kernel_dram.json")))
    out = tmp_path / "dram.json"
    subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_dram_summary.py"), str(out),
                    "exact+t2:268435456:profiles/r02_launches_weak16384.csv",
                    "exact:268435456:profiles/r02_launches_weak16384_t1.csv"], cwd=root, check=True, stdout=subprocess.DEVNULL)
    derived = json.load(open(out))
    assert derived == committed
    t2 = [c for c in committed["captures"] if c["arith"] == "exact+t2"][0]
    assert 1.0 <= t2["dram_bytes_per_launch"] / t2["algorithmic_bytes_per_launch"] < 1.05      # redundant halo reads only
    import bench
    got = bench.ncu_traffic_per_launch("exact+t2", 268435456)
    assert got and got["dram_bytes_per_launch"] == t2["dram_bytes_per_launch"]
