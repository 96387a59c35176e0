"""CPU-only: the reference arm of bench.py (CPU port on the host cores) runs and prints one well-formed JSON line."""

from __future__ import annotations

import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_one_json_line():
    proc = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "S", "--steps", "1", "--warmup", "3", "--cpu-budget", "0.3"],
                          capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "exposure-path traversals/sec" and d["unit"] == "traversals/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "traversals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_ignores_torchrun_thread_cap_and_never_loads_the_engine():
    """Under torchrun the children inherit OMP_NUM_THREADS=1; the reference arm must still use every host core it may run on,
    and it must not touch libabb200.so (ABB_LIB points at a file that does not exist: loading the engine would raise)."""
    import os

    env = dict(os.environ, OMP_NUM_THREADS="1", ABB_LIB="does-not-exist.so", RANK="0", WORLD_SIZE="1")
    proc = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "S", "--steps", "1", "--warmup", "3", "--cpu-budget", "0.3", "--gpus", "2"],
                          capture_output=True, text=True, timeout=600, env=env)
    assert proc.returncode == 0, proc.stderr[-2000:]
    d = json.loads([ln for ln in proc.stdout.splitlines() if ln.strip()][0])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["n_gpus"] == 2 and d["impl"] == "reference"
    ref = d["cpu_baseline"]["reference_python"]
    assert ref is None or ref["impact_of"][0]["findings_per_s_single_process"] > 0


def test_reference_arm_other_ranks_exit_silently():
    import os

    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    proc = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "S", "--gpus", "2"], capture_output=True, text=True, timeout=120, env=env)
    assert proc.returncode == 0 and proc.stdout.strip() == ""


def test_only_the_json_line_reaches_stdout():
    """Library chatter written to file descriptor 1 (NCCL's version banner at N > 1) or printed must not pollute the one-line contract."""
    code = "import os, bench; bench.claim_stdout(); os.write(1, b'NCCL version 2.28.9+cuda12.9\\n'); print('noise', flush=True); bench.emit({'ok': 1})"
    proc = subprocess.run([sys.executable, "-c", code], cwd=str(ROOT), capture_output=True, text=True, timeout=120)
    assert proc.returncode == 0, proc.stderr[-2000:]
    assert proc.stdout.strip() == '{"ok": 1}'
    assert "NCCL version" in proc.stderr and "noise" in proc.stderr
