"""bench.py's host-side arithmetic (no GPU): SURVEY.md 8(d)'s algorithmic bytes and the profiler-derived block that
the bench line quotes from profiles/*_metrics.json instead of typed-in constants."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_formulas():
    W, A, Q, n = 781, 456, 4, 1 << 19
    ntt, merkle = bench.algorithmic_bytes(W, A, Q, n, 1, 4)
    assert ntt == 8 * W * n * 4 + 8 * A * n * 4 + 8 * Q * n * 3
    assert merkle == sum(8 * c * n * 2 + 32 * (2 * n * 2 - 16) for c in (W, A, Q))
    # the dominant launch group of the bench line: trace tree of a config-2 proof
    assert 8 * W * (2 * n) + 32 * (2 * 2 * n - 16) == 6618611200
    assert bench.rows_for(1024) == 1 << 19 and bench.rows_for(1) == 1 << 16 and bench.rows_for(8192) == 1 << 22


def test_profile_metrics_are_read_from_the_committed_capture():
    m = bench.profile_metrics()
    assert m is not None and m["_file"].startswith("profiles/") and m.get("commit") and m.get("date")
    with open(os.path.join(ROOT, "profiles", "current_metrics.json")) as f:
        assert json.load(f)["file"] == os.path.basename(m["_file"])
    p = bench.leaf_hash_profile(m, 98 * (1 << 20))
    assert p is not None and "leaf_hash" in p["source"]
    # DRAM traffic of the captured launch is within 2 % of the algorithmic bytes (no re-reads), and the instruction
    # count per permutation is what DESIGN.md 4.1 quotes
    assert abs(p["dram_bytes_per_launch"] / 6618611200 - 1) < 0.02
    assert 15e3 < p["thread_instructions_per_permutation"] < 30e3
    assert 0 < p["issue_active"] <= 1 and 0 < p["alu_pipe_busy"] <= 1


def test_workload_config_names_the_baseline_workload():
    class A:
        instances, gpus = 1024, 1
    c = bench.workload_config(A)
    assert c["trace_rows"] == 1 << 19 and c["trace_columns"] == 781 and "1024 scalar-muls" in c["workload"]
    assert "model" not in c
