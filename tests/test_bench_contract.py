"""CPU: the reference arm of bench.py prints the JSON line the driver expects."""
import json
import os
import subprocess
import sys

from conftest import ROOT

REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
            "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(extra, env=None):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--cpu-sample-n", "48", "--no-ref-table"] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()


def test_reference_arm_json_line():
    lines = _run([])
    rec = json.loads(lines[-1])
    assert REQUIRED <= set(rec)
    assert rec["impl"] == "reference" and rec["unit"] == "pairs/s" and rec["higher_is_better"] is True
    assert rec["vs_baseline"] is None and rec["value"] > 0
    # the unmodified loss.py (source tree here, oracle/_ref on the GPU box) whenever it is reachable
    from oracle.ref_loader import reference_available
    assert rec["cpu_baseline"]["kind"] == ("reference" if reference_available() else "port")
    assert rec["cpu_baseline"]["cores"] >= 1 and "cpu_model" in rec
    assert rec["e2e"] == {"value": rec["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in rec["config"]


def test_reference_arm_falls_back_to_the_port_without_the_reference(tmp_path):
    env = dict(os.environ, SUPCON_REFERENCE_ROOT=str(tmp_path), SUPCON_REFERENCE_DISABLE="1")
    rec = json.loads(_run([], env=env)[-1])
    assert rec["cpu_baseline"]["kind"] == "port"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    assert _run(["--gpus", "2"], env=env) == []
