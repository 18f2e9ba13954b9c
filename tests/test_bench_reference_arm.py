"""`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) at a reduced size: it must finish on the
host cores without a GPU and print ONE JSON line with the contract's keys.  Only the oracle port runs here."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "10", "--chi", "8", "--nk", "4",
                        "--steps", "1", "--warmup", "0"] + extra, capture_output=True, text=True, timeout=300, env=e, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_contract_line():
    lines = run_bench([])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["unit"] == "sweeps/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("QTT linsolve sweeps/s")
    for key in ("n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs and prints; the other ranks exit 0 without work."""
    assert run_bench([], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
