"""ORACLE (test infrastructure) -- numpy restatement of the NMF solvers the reference
reaches through ``cNMF._nmf`` (cnmf.py:661-674 -> sklearn.decomposition.non_negative_factorization).

The arithmetic lives in scikit-learn 1.9.0 (third-party, NOT under /root/reference;
``SK/`` below = site-packages/sklearn).  Each function cites the lines it restates.
Pinned by tests/test_oracle_golden.py against fixtures produced by the reference itself
(oracle/make_golden.py) and against live sklearn calls.
"""
import numpy as np

EPSILON = np.finfo(np.float32).eps  # SK/decomposition/_nmf.py:32


def init_random(X_mean, n_samples, n_features, k, seed, dtype=np.float64):
    """SK/decomposition/_nmf.py:296-307: H is drawn FIRST, then W, from the legacy
    RandomState(seed) (SK/utils/validation.py check_random_state); cast, then abs."""
    avg = np.sqrt(X_mean / k)
    rng = np.random.RandomState(seed)
    H = avg * rng.standard_normal(size=(k, n_features)).astype(dtype, copy=False)
    W = avg * rng.standard_normal(size=(n_samples, k)).astype(dtype, copy=False)
    np.abs(H, out=H)
    np.abs(W, out=W)
    return W, H


def frobenius_error(X, W, H):
    """SK/decomposition/_nmf.py:113-127 (dense branch) with square_root=True."""
    return np.sqrt(((X - W @ H) ** 2).sum())


def frobenius_error_trace(X, W, H, norm_X_sq=None):
    """SK/decomposition/_nmf.py:116-120 (sparse branch; the reference's default X is CSR):
    ||X||^2 + <W^T W H, H> - 2 <X H^T, W>.  This is the identity the CUDA path uses."""
    if norm_X_sq is None:
        norm_X_sq = float((X.astype(np.float64) ** 2).sum())
    norm_WH = np.sum((W.T @ W) * (H @ H.T))
    cross = np.sum((X @ H.T) * W)
    return np.sqrt(max(norm_X_sq + norm_WH - 2.0 * cross, 0.0))


def mu_frobenius(X, W, H, tol=1e-4, max_iter=1000, l1_reg_W=0.0, l2_reg_W=0.0,
                 l1_reg_H=0.0, l2_reg_H=0.0, update_H=True, error_fn=frobenius_error):
    """Multiplicative-update NMF, beta=2.
    SK/decomposition/_nmf.py:726-888 (loop), :535-549,610-624 (W), :633-635,696-721 (H)."""
    W = W.copy()
    H = H.copy()
    err0 = prev = error_fn(X, W, H)
    XHt = HHt = None
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        if XHt is None:
            XHt = X @ H.T
        num = XHt if update_H else XHt.copy()
        if HHt is None:
            HHt = H @ H.T
        den = W @ HHt
        if l1_reg_W > 0:
            den += l1_reg_W
        if l2_reg_W > 0:
            den = den + l2_reg_W * W
        den[den == 0] = EPSILON
        num /= den
        W *= num
        if update_H:
            num = W.T @ X
            den = np.linalg.multi_dot([W.T, W, H])
            if l1_reg_H > 0:
                den += l1_reg_H
            if l2_reg_H > 0:
                den = den + l2_reg_H * H
            den[den == 0] = EPSILON
            num /= den
            H *= num
            XHt = HHt = None
        if tol > 0 and n_iter % 10 == 0:
            err = error_fn(X, W, H)
            if (prev - err) / err0 < tol:
                break
            prev = err
    return W, H, n_iter


def _cd_sweep(A, Gram, B):
    """SK/decomposition/_cdnmf_fast.pyx:8-37 restated row-outer (rows are independent,
    verified identical to the t-outer Cython order).
    Vectorised over rows; sequential over the K coordinates."""
    K = A.shape[1]
    viol = 0.0
    for t in range(K):
        g = -B[:, t] + A @ Gram[t, :]
        pg = np.where(A[:, t] == 0, np.minimum(0.0, g), g)
        viol += np.abs(pg).sum()
        if Gram[t, t] != 0:
            A[:, t] = np.maximum(A[:, t] - g / Gram[t, t], 0.0)
    return viol


def cd_frobenius(X, W, H, tol=1e-4, max_iter=1000, l1_reg_W=0.0, l2_reg_W=0.0,
                 l1_reg_H=0.0, l2_reg_H=0.0, update_H=True):
    """Coordinate-descent NMF (the reference's DEFAULT for beta_loss='frobenius',
    cnmf.py:629-631).  SK/decomposition/_nmf.py:399-518 (loop, shuffle=False),
    :369-396 (one half-step)."""
    W = W.copy()
    Ht = np.ascontiguousarray(H.T.copy())
    viol0 = 1.0
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        HHt = Ht.T @ Ht
        XHt = X @ Ht
        if l2_reg_W != 0:
            HHt = HHt.copy()
            HHt.flat[:: HHt.shape[0] + 1] += l2_reg_W
        if l1_reg_W != 0:
            XHt = XHt - l1_reg_W
        viol = _cd_sweep(W, HHt, XHt)
        if update_H:
            WtW = W.T @ W
            XtW = X.T @ W
            if l2_reg_H != 0:
                WtW = WtW.copy()
                WtW.flat[:: WtW.shape[0] + 1] += l2_reg_H
            if l1_reg_H != 0:
                XtW = XtW - l1_reg_H
            viol += _cd_sweep(Ht, WtW, XtW)
        if n_iter == 1:
            viol0 = viol
        if viol0 == 0:
            break
        if viol / viol0 <= tol:
            break
    return W, Ht.T.copy(), n_iter


def nmf(X, k, seed, solver="mu", tol=1e-4, max_iter=1000, dtype=np.float64,
        alpha_W=0.0, alpha_H=0.0, l1_ratio=0.0):
    """One restart as cNMF.factorize issues it (cnmf.py:738-741): random init + solver."""
    X = np.asarray(X, dtype=dtype)
    n, g = X.shape
    W, H = init_random(X.mean(), n, g, k, seed, dtype=dtype)
    l1W, l2W, l1H, l2H = reg_terms(n, g, alpha_W, alpha_H, l1_ratio)
    if solver == "mu":
        return mu_frobenius(X, W, H, tol, max_iter, l1W, l2W, l1H, l2H)
    return cd_frobenius(X, W, H, tol, max_iter, l1W, l2W, l1H, l2H)


def reg_terms(n_samples, n_features, alpha_W, alpha_H, l1_ratio):
    """SK/decomposition/_nmf.py:1249-1260 (alpha_H='same' never reaches here: cNMF always
    passes a float, cnmf.py:619-620)."""
    l1W = n_features * alpha_W * l1_ratio
    l1H = n_samples * alpha_H * l1_ratio
    l2W = n_features * alpha_W * (1.0 - l1_ratio)
    l2H = n_samples * alpha_H * (1.0 - l1_ratio)
    return l1W, l2W, l1H, l2H


def refit(X, H, solver="mu", tol=1e-4, max_iter=1000, dtype=np.float64):
    """cNMF.refit_usage (cnmf.py:776-802): NMF with H fixed (update_H=False).
    W0: SK/decomposition/_nmf.py:1223-1228 -- 'mu': constant sqrt(X.mean()/k); 'cd': zeros."""
    X = np.asarray(X, dtype=dtype)
    H = np.asarray(H, dtype=dtype)
    n = X.shape[0]
    k = H.shape[0]
    if solver == "mu":
        W0 = np.full((n, k), np.sqrt(X.mean() / k), dtype=dtype)
        W, _, it = mu_frobenius(X, W0, H, tol, max_iter, update_H=False)
    else:
        W0 = np.zeros((n, k), dtype=dtype)
        W, _, it = cd_frobenius(X, W0, H, tol, max_iter, update_H=False)
    return W, it
