"""bench.py without a GPU: the product arm refuses to run (there is no CPU path to fall back to), the reference arm falls
back to the scalar port on the host cores and still prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
ENV = dict(os.environ, CUDA_VISIBLE_DEVICES="")


def test_reference_arm_without_a_gpu_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                        "--workload", "720p", "--cpu-budget-px", "2e5"], capture_output=True, text=True, env=ENV, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Gpix/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Gpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["config"]["width"] == 1280


def test_product_arm_without_a_gpu_fails_loudly():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--workload", "720p"],
                       capture_output=True, text=True, env=ENV, timeout=300)
    assert r.returncode != 0
    assert "no CPU implementation" in r.stderr
    assert not r.stdout.strip()
