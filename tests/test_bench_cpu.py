"""bench.py contract pieces that do not need a GPU: the reference arm's JSON line."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3",
                        "--warmup", "3", "--cpu-n", "192"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    if "unavailable" in line:
        pytest.skip(line["unavailable"])
    assert line["metric"] == "Gcell-updates/s" and line["unit"] == "Gcell-updates/s" and line["higher_is_better"]
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_nonzero_rank_of_reference_arm_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_stdout_carries_the_json_line_only():
    """quiet_stdout() points fd 1 at stderr for the run (NCCL prints its banner on stdout at
    N > 1) and emit() writes the line on the original stdout."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.quiet_stdout(); "
            "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n'); print('library chatter'); "
            "bench.emit('{\"metric\": \"x\"}')" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout == '{"metric": "x"}\n'
    assert "NCCL version" in p.stderr and "library chatter" in p.stderr
