"""CPU check of the driver contract of bench.py (reference arm, which needs no GPU): ONE JSON line on stdout carrying the
keys the driver reads, the BASELINE.json metric, and the tier-specific `cpu_baseline` / `e2e` objects; ranks other than 0
print nothing and exit 0."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(extra_env, *args):
    env = dict(os.environ, CAPDEC_CPU_SAMPLE="2", **extra_env)
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", *args],
                          capture_output=True, text=True, env=env, timeout=900, cwd=str(ROOT))


def test_reference_arm_prints_one_contract_line():
    r = _run({})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    j = json.loads(lines[0])
    base = json.loads((ROOT / "BASELINE.json").read_text())
    assert j["impl"] == "reference" and j["metric"] == base["metric"] and j["unit"] == "captions/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in j, k
    assert j["value"] > 0 and j["higher_is_better"] is True and j["vs_baseline"] is None and j["data"] == "synthetic"
    assert "workload" in j["config"] and "model" not in j["config"]
    cb = j["cpu_baseline"]
    # the reference's own classes whenever tools/install_ref.sh has placed them under baseline/_ref, else the oracle port
    have_ref = (ROOT / "baseline" / "_ref" / "train.py").exists()
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] >= 1 and cb["value"] == j["value"]
    assert "sample" in cb and ("baseline/_ref/train.py" in cb["sample"]) == have_ref
    # same `config` object as the GPU arm prints (the driver's same_config check): identical keys and strings
    sys.path.insert(0, str(ROOT))
    import bench
    assert j["config"] == bench.bench_config("c2", 1)
    assert j["e2e"] == {"value": j["value"], "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_falls_back_to_the_oracle_port_without_baseline_ref(tmp_path):
    """A box that never received baseline/_ref still gets a CPU line (kind "port")."""
    import shutil
    work = tmp_path / "repo"
    work.mkdir()
    for name in ("bench.py", "oracle", "BASELINE.json"):
        src = ROOT / name
        (shutil.copytree if src.is_dir() else shutil.copy)(src, work / name)
    env = dict(os.environ, CAPDEC_CPU_SAMPLE="2")
    r = subprocess.run([sys.executable, str(work / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=900, cwd=str(work))
    assert r.returncode == 0, r.stderr[-2000:]
    j = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][-1])
    assert j["cpu_baseline"]["kind"] == "port" and j["value"] > 0


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""
