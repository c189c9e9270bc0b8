"""bench.py's reference arm runs the unmodified reference on the CPU (from /root/reference or its staged copy oracle/_ref;
the oracle port only where neither exists): one bounded step, JSON contract of the line."""
import json
import os
import subprocess
import sys

from helpers import ROOT


import pytest

from oracle import ref_import as R


@pytest.mark.parametrize("tree", ["default", "staged"])
def test_reference_arm_json_line(tree):
    env = dict(os.environ, OMP_NUM_THREADS="1")   # what torchrun exports: the arm pins its thread count itself
    if tree == "staged":
        if not R.stage_reference():
            pytest.skip("no staged reference tree (oracle/_ref) and no /root/reference to stage it from")
        env["CNEUS_REFERENCE_ROOT"] = R.STAGED_ROOT
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-rays", "128"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["metric"].startswith("rays/sec") and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == ("reference" if (tree == "staged" or R.reference_available()) else "port")
    assert cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"]
    assert d["config"]["rays_per_step"] == 128   # the arm states the true size of its step
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
