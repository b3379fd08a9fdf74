"""bench.py on a box without a GPU: the workload table against SURVEY section 8, and the reference arm
(`--impl reference`: the compiled reference, or the oracle port where it is not built) printing exactly one JSON
line with the contract's keys."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_reference_format_sizes_match_the_survey_table():
    # nnz lower (default export) = ndet * (off-diagonal per row / 2 + 1)
    for name, ndet, lower in (("cfg1", 8281, 3138499), ("cfg2", 42504, 2061444), ("cfg3", 1002001, 1113223111),
                              ("cfg4", 3312400, 5289902800)):
        spec = bench.workload_spec(name)
        assert round(bench.ref_size_per_row(spec, ndet) * ndet) == lower
    assert bench.ref_size_per_row(bench.workload_spec("cfg5"), 10) is None  # selected space: sampled instead
    with pytest.raises(SystemExit):
        bench.workload_spec("nope")


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sparse_op_build_nnz_per_s" and d["unit"] == "nnz/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"]
    assert 1 <= d["cpu_baseline"]["cores"] == d["config"]["processes"] <= 32 and d["cpu_baseline"]["single_core_value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"] == {"value": d["value"], "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "cfg1" in d["config"]["workload"]


def test_reference_arm_on_other_ranks_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload",
                          "cfg1", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_parity_gate_row_sample_covers_the_boundaries():
    """bench.gate_rows: first / last rows of the shard, both sides of alpha-string boundaries and of uniform CTA-range
    boundaries, random rest; every row inside the shard, sorted, unique."""
    import numpy as np
    rng = np.random.default_rng(0)
    row0, nloc, nb = 414050, 414050, 1820  # rank 1 of 8 at config 4
    rows = bench.gate_rows(row0, nloc, nb, rng, 1000)
    assert len(rows) >= 1000 and np.all(np.diff(rows) > 0) and rows[0] == row0 and rows[-1] == row0 + nloc - 1
    bounds = np.arange(-(-row0 // nb) * nb, row0 + nloc, nb)
    hit = [b for b in bounds if b in rows and b - 1 in rows]
    assert len(hit) >= 20  # alpha-string boundaries, both sides
    per = -(-nloc // 148)
    assert any((row0 + per * k) in rows for k in range(1, 148))
    assert len(bench.gate_rows(5, 0, 7, rng, 100)) == 0
    small = bench.gate_rows(0, 3, 0, rng, 50)
    assert small.tolist() == [0, 1, 2]


def test_golden_e0_lookup():
    for key, n in (("cfg3", 14), ("syn10", 10)):
        spec = bench.workload_spec(key)
        spec["key"] = key
        e0, tol, src = bench.golden_e0(spec)
        assert e0 is not None and tol == 1e-10 and "e0_syn.json" in src
    spec = bench.workload_spec("cfg1")
    spec["key"] = "cfg1"
    assert bench.golden_e0(spec)[0] == -14.617409507  # pyci/test/test_routines.py:44
    spec = bench.workload_spec("cfg4")   # the matrix does not fit a CPU box: string-driven direct CI (make_golden_e0_direct.py)
    spec["key"] = "cfg4"
    e0, tol, src = bench.golden_e0(spec)
    assert abs(e0 - (-34.525629697)) < 1e-8 and tol == 1e-10 and "direct CI" in src
    spec = bench.workload_spec("cfg5")
    spec["key"] = "cfg5"
    assert bench.golden_e0(spec) == (None, None, None)


def test_leg_watchdog_prints_the_line_when_an_extra_leg_hangs():
    """bench.leg_watchdog: the headline line survives an extra leg that never returns; a cancelled timer does nothing."""
    code = ("import sys, time; sys.path.insert(0, %r); import bench; bench.claim_stdout(); "
            "line = {'metric': 'm', 'parity': {'ok': True}}; "
            "bench.leg_watchdog(0.3, int(sys.argv[1]), line, 'cfg5', 'label'); time.sleep(30)") % ROOT
    out = subprocess.run([sys.executable, "-c", code, "0"], capture_output=True, text=True, timeout=25)
    assert out.returncode == 0
    d = json.loads(out.stdout.strip())
    assert d["metric"] == "m" and "watchdog" in d["cfg5"]["error"] and d["cfg5"]["workload"] == "label"
    other = subprocess.run([sys.executable, "-c", code, "1"], capture_output=True, text=True, timeout=25)
    assert other.returncode == 0 and other.stdout.strip() == ""
    t = bench.leg_watchdog(0.2, 0, {}, "x", "y")
    t.cancel()
    import time
    time.sleep(0.4)  # still here: a cancelled timer does not fire
