"""CPU: the `bench.py --impl reference` arm (the CPU oracle port timed on the host cores) prints one JSON line with the
keys the driver reads, finishes quickly on a bounded sample, and never touches a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, extra_args=("--ref_kind", "port")):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--cpu_sample_batch", "1", *extra_args], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()


def test_reference_arm_prints_the_contract_line():
    lines = _run()
    line = json.loads(lines[-1])
    assert line["impl"] == "reference" and line["metric"] == "video-text samples/sec" and line["unit"] == "samples/s"
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["warmup"] == 1
    assert line["value"] > 0 and abs(line["ms_per_step"] * 1e-3 * line["value"] - 1.0) < 1e-6  # batch 1 per step
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and 1 <= cb["cores"] <= 32 and cb["value"] == line["value"] and "oracle" in cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert "FT-Align" in line["config"]["workload"]


def test_reference_arm_runs_on_rank_zero_only():
    lines = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert lines == [] or all(not l.startswith("{") for l in lines)


def test_reference_arm_runs_the_staged_unmodified_reference_at_the_per_rank_batch():
    """kind = "reference": oracle/_ref (oracle/build_ref.py), same per-rank batch as the GPU arm (here 2 to stay quick)"""
    sys.path.insert(0, ROOT)
    from oracle import build_ref
    if build_ref.ref_root() is None:
        import pytest
        pytest.skip("reference not staged")
    lines = _run(extra_args=("--ref_kind", "reference", "--batch", "2", "--mode", "ft_align"))
    line = json.loads(lines[-1])
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and "unmodified reference" in cb["sample"] and cb["cores"] >= 1
    assert line["steps"] == 2 and abs(line["ms_per_step"] * 1e-3 * line["value"] - 2.0) < 1e-6
    assert "per-GPU batch 2" in line["config"]["workload"]
