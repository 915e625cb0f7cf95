"""CPU: the reference arm of bench.py (`--impl reference`: the reference's own CPU code from oracle/_ref on the
host cores) prints ONE JSON line with the keys the driver reads, on the same `config` / `metric` / `unit` as the
GPU arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", *extra], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run()
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert abs(d["value"] - d["details"]["frames_per_step"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert cb["cores"] <= len(os.sched_getaffinity(0))          # the cores it says are cores it may use
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # same workload description as the GPU arm prints for the same flags
    sys.path.insert(0, ROOT)
    import bench
    bench.CFG = "D"
    assert d["config"] == bench._config() and d["metric"] == bench._metric()
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert base["metric"].split()[0].lower() in d["metric"].lower()


def test_reference_arm_other_configs_and_ranks():
    d = _run("--config", "B")
    assert "config B" in d["config"]["workload"] and d["value"] > 0
    # under torchrun only rank 0 works and prints; the other ranks exit 0 silently
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT,
                         env=env)
    assert out.returncode == 0 and not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
