"""bench.py's CPU arm (`--impl reference`) on a tiny workload: one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cells", "600", "--donors", "30",
                          "--contexts", "4", "--hk-rank", "3", "--snps", "16", "--cpu-sample-snps", "3", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, check=True)
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in rec
    assert rec["impl"] == "reference" and rec["unit"] == "tests/s" and rec["value"] > 0
    assert rec["cpu_baseline"]["kind"] == "port" and rec["e2e"]["h2d_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--cells", "200", "--snps", "4"],
                         capture_output=True, text=True, timeout=120, check=True, env=env)
    assert out.stdout.strip() == ""
