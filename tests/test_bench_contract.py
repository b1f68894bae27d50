"""bench.py's reference arm (the CPU oracle on host cores) prints the contract's JSON line; the GPU arm's helpers (algorithmic
bytes / flops of SURVEY.md 8d) are pure functions checked here without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "16"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "LMPC solves/sec (batched)" and line["unit"] == "solves/s"
    assert line["higher_is_better"] is True and line["dtype"] == "f64" and line["gpu_launches"] == 0
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "sample" in cb and cb["value"] == line["value"]
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_counts():
    sys.path.insert(0, ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    # 3 376 B per solve with a shared model at ph = 20, 6 672 B with a per-controller model
    assert b.algorithmic_bytes(20, True) == 3376
    assert b.algorithmic_bytes(20, False) == 6672
    f75 = b.algorithmic_flops(20, 75, 2, 1)
    f100 = b.algorithmic_flops(20, 100, 2, 1)
    assert 7.0e6 < f75 < 8.5e6 and f100 > f75
    # the structured NLMPC count is far below the dense O(nz^3) one
    nz = 30 * 3 + 30 * 2 + 1
    assert b.nlmpc_flops(3, 2, 30, 30, 2, 1, 0) < 2.0 * nz ** 3 / 3.0
