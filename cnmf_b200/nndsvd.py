"""NNDSVD initialisation (`--init nndsvd`, cnmf.py:1252; scikit-learn's `_initialize_nmf`, SK/decomposition/_nmf.py:247-371)
on the host, for the batched GPU solve to start from (`Dataset.factorize(W0=, H0=)` -> `cnmf_factorize_init`).

The reference hands `init` and `random_state=seed` to `non_negative_factorization` (cnmf.py:672, 738-739), which
  1. takes a randomized SVD of X with `n_components` triplets (`randomized_svd`, SK/utils/extmath.py: range finder with
     10 oversamples, 7 (k < 0.1 min(shape)) or 4 LU-normalised power iterations, QR, SVD of the thin projection, sign
     flip) -- the seed only enters through the Gaussian test matrix,
  2. keeps the leading triplet's absolute values and, for every further triplet, the dominant of its positive / negative
     parts (Boutsidis & Gallopoulos), scaled by sqrt(S_j * sigma),
  3. zeroes entries below eps = 1e-6 and, for the 'a' / 'ar' variants, fills the zeros with the mean of X / with small
     random values.
This module restates those steps with numpy / scipy.linalg (same LAPACK entry points, same order of operations) so
that the starting point is the one the reference's call would use; tests/test_host_logic.py holds it against
scikit-learn's own function.  It is an SVD on the host, not a CUDA kernel: the random initialisation is the path's
default (and the only one the device generator covers); NNDSVD runs once per restart before the batched solve.
"""
import numpy as np
from scipy import linalg, sparse

INITS = ("random", "nndsvd", "nndsvda", "nndsvdar")


def resolve_init(init, n_components, n_samples, n_features):
    """scikit-learn's validation of `init` (SK/decomposition/_nmf.py:283-300): None -> 'nndsvda' when it is possible."""
    if init is not None and init not in INITS:
        raise ValueError("Invalid init parameter: got %r instead of one of %r" % (init, (None,) + INITS))
    if init is not None and init != "random" and n_components > min(n_samples, n_features):
        raise ValueError("init = '{}' can only be used when n_components <= min(n_samples, n_features)".format(init))
    if init is None:
        init = "nndsvda" if n_components <= min(n_samples, n_features) else "random"
    return init


def _svd_flip_v(u, v):
    """svd_flip(u, v, u_based_decision=False): sign of the largest |entry| of every row of v made positive."""
    idx = np.argmax(np.abs(v), axis=1)
    signs = np.sign(v[np.arange(v.shape[0]), idx])
    u *= signs[np.newaxis, :]
    v *= signs[:, np.newaxis]
    return u, v


def _svd_flip_u(u, v):
    idx = np.argmax(np.abs(u.T), axis=1)
    signs = np.sign(u.T[np.arange(u.shape[1]), idx])
    u *= signs[np.newaxis, :]
    v *= signs[:, np.newaxis]
    return u, v


def randomized_svd(M, n_components, seed):
    """`randomized_svd(M, n_components, random_state=seed)` with scikit-learn's defaults (n_oversamples=10, n_iter='auto',
    power_iteration_normalizer='auto' -> LU, transpose='auto', flip_sign=True, gesdd)."""
    rng = np.random.RandomState(seed)
    n_random = n_components + 10
    n_samples, n_features = M.shape
    n_iter = 7 if n_components < 0.1 * min(M.shape) else 4
    transpose = n_samples < n_features
    if transpose:
        M = M.T
    Q = rng.normal(size=(M.shape[1], n_random))
    if M.dtype == np.float32:
        Q = Q.astype(np.float32, copy=False)
    for _ in range(n_iter):
        Q, _ = linalg.lu(M @ Q, permute_l=True, check_finite=False)
        Q, _ = linalg.lu(M.T @ Q, permute_l=True, check_finite=False)
    Q, _ = linalg.qr(M @ Q, mode="economic", check_finite=False)
    B = Q.T @ M
    if sparse.issparse(B):
        B = B.toarray()
    Uhat, s, Vt = linalg.svd(B, full_matrices=False, lapack_driver="gesdd")
    U = Q @ Uhat
    if not transpose:
        U, Vt = _svd_flip_u(U, Vt)
        return U[:, :n_components], s[:n_components], Vt[:n_components, :]
    U, Vt = _svd_flip_v(U, Vt)
    return Vt[:n_components, :].T, s[:n_components], U[:, :n_components].T


def nndsvd_init(X, n_components, seed, init="nndsvd", eps=1e-6):
    """(W n x k, H k x g) as `_initialize_nmf(X, n_components, init, eps, random_state=seed)` returns them."""
    if init not in ("nndsvd", "nndsvda", "nndsvdar"):
        raise ValueError("nndsvd_init: init must be one of nndsvd / nndsvda / nndsvdar (got %r)" % (init,))
    U, S, V = randomized_svd(X, n_components, seed)
    W = np.zeros_like(U)
    H = np.zeros_like(V)
    W[:, 0] = np.sqrt(S[0]) * np.abs(U[:, 0])
    H[0, :] = np.sqrt(S[0]) * np.abs(V[0, :])
    for j in range(1, n_components):
        x, y = U[:, j], V[j, :]
        x_p, y_p = np.maximum(x, 0), np.maximum(y, 0)
        x_n, y_n = np.abs(np.minimum(x, 0)), np.abs(np.minimum(y, 0))
        x_p_nrm, y_p_nrm = np.sqrt(np.dot(x_p, x_p)), np.sqrt(np.dot(y_p, y_p))      # SK/utils/extmath.py norm(): sqrt(squared_norm)
        x_n_nrm, y_n_nrm = np.sqrt(np.dot(x_n, x_n)), np.sqrt(np.dot(y_n, y_n))
        m_p, m_n = x_p_nrm * y_p_nrm, x_n_nrm * y_n_nrm
        if m_p > m_n:
            u, v, sigma = x_p / x_p_nrm, y_p / y_p_nrm, m_p
        else:
            u, v, sigma = x_n / x_n_nrm, y_n / y_n_nrm, m_n
        lbd = np.sqrt(S[j] * sigma)
        W[:, j] = lbd * u
        H[j, :] = lbd * v
    W[W < eps] = 0
    H[H < eps] = 0
    if init == "nndsvda":
        avg = X.mean()
        W[W == 0] = avg
        H[H == 0] = avg
    elif init == "nndsvdar":
        rng = np.random.RandomState(seed)
        avg = X.mean()
        W[W == 0] = abs(avg * rng.standard_normal(size=len(W[W == 0])) / 100)
        H[H == 0] = abs(avg * rng.standard_normal(size=len(H[H == 0])) / 100)
    return W, H
