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
    # whole sweeps, never an extrapolated window; thread count set explicitly; the GPU library stays unloaded in this arm
    assert "full sweep" in cb["sample"] and cb["sweep_seconds"] and d["config"]["timed_sweeps"] == len(cb["sweep_seconds"])


def test_reference_arm_never_imports_the_gpu_package():
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--n', '8', '--chi', '4', '--nk', '3', '--steps', '1', '--warmup', '0'];"
            "runpy.run_path('bench.py', run_name='__main__');"
            "bad = [m for m in sys.modules if m.startswith('itensorqtt')]; assert not bad, bad")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]


def test_reference_arm_uses_all_cores_under_torchrun_env():
    """torchrun exports OMP_NUM_THREADS=1; the arm sets the BLAS pool itself."""
    d = json.loads(run_bench([], env={"OMP_NUM_THREADS": "1"})[0])
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1) or d["cpu_baseline"]["cores"] >= 1


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs and prints; the other ranks exit 0 without work."""
    assert run_bench([], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
