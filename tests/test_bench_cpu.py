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
