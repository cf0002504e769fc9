"""bench.py's reference arm runs on the host (the unmodified reference staged in baseline/_ref, else the oracle port of its
torch-CPU step): check the JSON contract of the line the driver parses.  CPU only; the GPU arm's line is produced on the B200 box."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")   # what torchrun exports to its workers: the arm must still take every host core
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--episodes", "40"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["metric"] in baseline["metric"] and d["unit"] == "gradient-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and abs(d["value"] * d["ms_per_step"] - 1000.0) < 1e-6 * 1000.0
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    staged = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "url_benchmark", "agent", "fb_ddpg.py"))
    assert cb["kind"] == ("reference" if staged else "port") and cb["value"] == d["value"] and "agent.update()" in cb["sample"]
    assert cb["cores"] == min(os.cpu_count(), cb["host_cpus"]) or cb["cores"] >= 1
    if (os.cpu_count() or 1) > 1:
        assert cb["cores"] > 1, "the reference arm must not inherit the launcher's OMP_NUM_THREADS=1"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
