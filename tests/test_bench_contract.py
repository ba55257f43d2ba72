"""bench.py contract on CPU: the reference arm prints exactly ONE JSON line on stdout with the agreed keys."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(extra, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "32", "--steps", "1",
                          "--warmup", "0"] + extra, capture_output=True, text=True, timeout=600, cwd=ROOT,
                         env=dict(os.environ, **(env or {})))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    line = _run([])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "cascade_volumes_per_sec_128cubed" and line["unit"] == "volumes/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    # the reference's own modules when /root/reference or oracle/_ref is there (kind "reference"), else the oracle port
    from oracle import ref_loader
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port")
    assert "workload" in line["config"] and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_falls_back_to_the_oracle_port():
    line = _run([], env={"DP_BENCH_PORT": "1"})
    assert line["cpu_baseline"]["kind"] == "port" and line["value"] > 0


def test_reference_arm_of_the_training_workload():
    line = _run(["--workload", "train"])
    assert line["impl"] == "reference" and line["metric"] == "dose_pyfer_train_samples_per_sec_128cubed"
    assert line["value"] > 0 and line["unit"] == "samples/s"


def test_non_zero_ranks_of_the_reference_arm_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "32", "--steps", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
