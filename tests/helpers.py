"""Shared test helpers: golden fixtures, error metrics."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "golden"))

from oracle import neus_oracle as O  # noqa: E402
import make_golden as MG  # noqa: E402  (only its CASES table / digest helpers; reference import is lazy)

CASE_NAMES = MG.available_cases()


def load_case(name):
    cfg, trained, n_rays = MG.case_cfg(name)
    G = dict(np.load(os.path.join(HERE, "golden", name + ".npz")))
    P = MG.case_params(name)
    assert str(G["params_sha256"]) == MG.params_digest(P), "synthetic parameter generator drifted from the fixtures"
    return cfg, P, G


def T(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a)).to(dtype)


def rel_err(a, b):
    """max |a-b| / max |b|  (the metric BASELINE.md section 2 states for per-ray outputs)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


# measured errors of the GPU parity tests: collected during the session, written to gpurun_out/parity_errors.json by
# conftest.pytest_sessionfinish (copied to profiles/ per round; the tolerances in the tests are set from these numbers)
MEASURED = {}


def record(test, case, key, value):
    MEASURED.setdefault(test, {}).setdefault(case, {})[key] = float(value)
    return float(value)
