import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def load_golden(tag):
    """Fixture produced by the reference itself (oracle/make_golden.py). Returns a dict plus
    the derived inputs the reference's prepare() would have written:
      X        normalised HVG counts (f64), cnmf.py:533-542
      tpm      TPM over all genes, cnmf.py:245-251;  tpm_std = std(ddof=0), cnmf.py:440
    """
    z = dict(np.load(os.path.join(GOLDEN_DIR, tag + ".npz"), allow_pickle=True))
    counts = z["counts"].astype(np.float64)
    hvg = z["hvg_idx"]
    X = counts[:, hvg].copy()
    X /= X.std(axis=0, ddof=1)
    tpm = counts / counts.sum(axis=1, keepdims=True) * 1e6
    z["X"] = X
    z["tpm"] = tpm
    z["tpm_std"] = tpm.std(axis=0, ddof=0)
    z["solver"] = str(z["solver"])
    return z


@pytest.fixture(scope="session", params=["sim_mu", "sim_cd"])
def golden(request):
    return load_golden(request.param)
