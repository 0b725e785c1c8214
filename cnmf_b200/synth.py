"""Deterministic synthetic counts for benchmarks and tests.

Mirrors what the reference pipeline would hold after ``prepare`` (cnmf.py:487-556):
a non-negative cells x genes matrix whose columns are scaled to unit variance without
centring (cnmf.py:542), plus the matching TPM matrix (cnmf.py:245-251) and per-gene
TPM mean / std(ddof=0) (cnmf.py:436-445).

Not an oracle and not a kernel: plain numpy on the host.
"""
import itertools

import numpy as np


def make_counts(n_cells, n_genes, k_true=12, seed=0, libsize=1500.0):
    """Poisson counts from Dirichlet usages x Gamma spectra. Returns int32 (cells x genes)
    with all-zero rows / columns removed."""
    rng = np.random.RandomState(seed)
    U = rng.dirichlet(np.full(k_true, 0.3), size=n_cells)
    S = rng.gamma(0.3, 1.0, size=(k_true, n_genes))
    S /= S.sum(axis=1, keepdims=True)
    lib = rng.lognormal(np.log(libsize), 0.4, size=n_cells)
    lam = (U * lib[:, None]) @ S
    counts = rng.poisson(lam).astype(np.int32)
    counts = counts[:, counts.sum(axis=0) > 0]
    counts = counts[counts.sum(axis=1) > 0]
    return counts


def normalise(counts, dtype=np.float32):
    """X = counts / std(counts, ddof=1) per gene (no centring), cnmf.py:534,542.
    Genes with zero variance are dropped first (they would give inf)."""
    c = counts.astype(np.float64)
    sd = c.std(axis=0, ddof=1)
    keep = sd > 0
    X = c[:, keep] / sd[keep]
    return np.ascontiguousarray(X.astype(dtype)), keep


def tpm_of(counts):
    """TPM over the same genes (cnmf.py:245-251) + mean / std(ddof=0) (cnmf.py:439-440)."""
    c = counts.astype(np.float64)
    tpm = c / c.sum(axis=1, keepdims=True) * 1e6
    return tpm, tpm.mean(axis=0), tpm.std(axis=0, ddof=0)


def restart_table(ks, n_iter, seed=14):
    """(k, iter, nmf_seed) rows in the reference's order (cnmf.py:597-610)."""
    ks = list(ks)
    k_list = sorted(set(ks))
    n_runs = len(ks) * n_iter
    np.random.seed(seed)
    seeds = np.random.randint(low=1, high=(2 ** 31) - 1, size=n_runs)
    rows = []
    for i, (k, r) in enumerate(itertools.product(k_list, range(n_iter))):
        rows.append((int(k), int(r), int(seeds[i])))
    return rows
