import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def repo_root():
    return ROOT


def pytest_sessionfinish(session, exitstatus):
    """Dump the errors the GPU parity tests measured (helpers.record) next to the other GPU-run artefacts."""
    try:
        import json
        import helpers
        if helpers.MEASURED:
            out = os.path.join(ROOT, "gpurun_out")
            os.makedirs(out, exist_ok=True)
            worst = {t: {k: max(c.get(k, 0.0) for c in cases.values()) for k in sorted({k for c in cases.values() for k in c})}
                     for t, cases in helpers.MEASURED.items()}
            with open(os.environ.get("CNEUS_PARITY_OUT", os.path.join(out, "parity_errors.json")), "w") as fh:
                json.dump({"metric": "max|err| / max|ref| unless the key says otherwise", "worst_over_cases": worst,
                           "per_case": helpers.MEASURED}, fh, indent=1, sort_keys=True)
    except Exception as e:   # never turn a green run red because of the report
        print("parity error report not written:", e)
