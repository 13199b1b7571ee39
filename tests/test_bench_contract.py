"""The bench.py contract, checked on the CPU through the reference arm (no GPU needed): exactly one JSON line
on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "grid_point_laplacian_steps_per_sec" and d["unit"] == "pt-steps/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["vs_baseline"] is None and "workload" in d["config"]
    # "reference" where the unmodified reference package is importable (build container: /root/reference; GPU box:
    # baseline/_ref), else the oracle port
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["sample"] and d["steps"] == 1 and d["warmup"] == 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_gpu_arm_does_not_import_the_oracle():
    """Only the cpu_baseline / reference legs of bench.py may touch oracle/ (the synthetic inputs come from
    bench_inputs.py)."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    gpu_arm = src[src.index("# ------------------------------------------------------------------------------------------ GPU arm"):src.index("# ------------------------------------------------------------------------------------------ banded arm")]
    banded = src[src.index("def banded_arm("):src.index("# ------------------------------------------------------------------------------------------ reference arm")]
    assert "oracle" not in gpu_arm.replace("cpu_baseline", "") and "oracle" not in banded
    pkg = os.path.join(ROOT, "gcm_filters_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, name)).read(), name
