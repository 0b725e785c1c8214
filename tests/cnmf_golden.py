"""Loader for the fixtures produced by the reference itself (oracle/make_golden.py)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(tag):
    """Returns the fixture dict plus the derived inputs the reference's prepare() would have written:
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
    z["init"] = str(z["init"]) if "init" in z else "random"
    # what was passed to the reference's prepare(beta_loss=...) and the beta the solver ran with (cnmf.py:629-631)
    bl = str(z["beta_loss"]) if "beta_loss" in z else ("2.0" if z["solver"] == "mu" else "frobenius")
    z["beta_loss_arg"] = {"2.0": 2.0, "frobenius": "frobenius"}.get(bl, bl)
    z["beta"] = {"kullback-leibler": 1, "itakura-saito": 0}.get(bl, 2)
    return z
