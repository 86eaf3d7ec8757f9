"""The bench line contract, checked on the committed lines in profiles/ (the lines themselves are produced on a B200 by
`python bench.py`; nothing here runs a kernel): every key the driver reads is present and self-consistent."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "roofline", "clocks", "e2e", "gpu_launches"]


def _line(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


@pytest.mark.parametrize("name", ["r2_bench_n1.json", "r2_bench_n2.json", "r2_bench_n4.json", "r2_bench_n8.json",
                                  "r2_bench_cyl3d_n1.json", "r2_bench_c1.json", "r2_bench_c3.json", "r2_bench_c4.json"])
def test_committed_bench_lines_follow_the_contract(name):
    d = _line(os.path.join(ROOT, "profiles", name))
    for k in REQUIRED:
        assert k in d, k
    assert d["metric"] == "rhs_dof_evals_per_s" and d["unit"] == "DOF-evals/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"], "the host-buffer figure cannot exceed the device-resident one"
    assert d["gpu_launches"] > 0 and d["value"] > 0
    # value = DOFs of the whole job per evaluation / time per evaluation
    if "dofs_per_gpu" in d and d["scaling"] == "weak":
        assert abs(d["value"] / (d["dofs_per_gpu"] * d["n_gpus"] / (d["ms_per_step"] * 1e-3)) - 1) < 1e-6
    if d["n_gpus"] == 1:
        c = d["cpu_baseline"]
        assert c["value"] > 0 and c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["sample"]
    elif name.startswith("r2_bench_n"):
        assert d["multirank_parity"]["max_rel_diff"] < 1e-12


def test_weak_scaling_lines_are_consistent():
    v = {n: _line(os.path.join(ROOT, "profiles", f"r2_bench_n{n}.json"))["value"] for n in (1, 2, 4, 8)}
    for n in (2, 4, 8):
        assert 0.85 * n * v[1] < v[n] < 1.05 * n * v[1], (n, v)


def test_bench_cli_parses():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "--workload" in out.stdout and "--impl" in out.stdout


def test_every_profile_named_in_the_readme_exists():
    import re
    txt = open(os.path.join(ROOT, "profiles", "README.md")).read()
    have = set(os.path.basename(p) for p in glob.glob(os.path.join(ROOT, "profiles", "*")))
    for name in re.findall(r"`(r2_[A-Za-z0-9_]+\.(?:json|txt|csv|log))`", txt):
        assert name in have, name
