"""The reference arm of bench.py (the oracle port timed on the host cores) runs without a GPU: check its JSON
contract here; the GPU arm is exercised on the B200 box."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                       timeout=600, env=e, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout.strip().splitlines()


def test_reference_arm_json_contract():
    lines = _run("--impl", "reference", "--workload", "sparse_attention", "--steps", "2", "--warmup", "1")
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "sparse_attention" and d["vs_baseline"] is None and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 times the CPU path; the other ranks print nothing and exit 0."""
    assert _run("--impl", "reference", "--workload", "sparse_attention", "--steps", "1", "--warmup", "1", "--gpus", "2",
                env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
