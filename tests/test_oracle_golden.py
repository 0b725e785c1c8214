"""Pin the oracle (oracle/nmf_ref.py, oracle/consensus_ref.py) against fixtures produced by
the UNMODIFIED reference (oracle/make_golden.py -> tests/golden/*.npz), and against live
scikit-learn calls with the kwargs the reference passes (cnmf.py:618-631,738-741).
CPU only."""
import warnings

import numpy as np
import pytest

from oracle import consensus_ref, nmf_ref
from cnmf_b200.synth import restart_table


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_seed_rule_matches_reference(golden):
    # cnmf.py:597-610 -- seeds drawn with np.random.seed(seed); randint(1, 2**31-1, n_runs)
    rows = restart_table(list(golden["ks"]), int(golden["n_iter"]), int(golden["seed"]))
    assert np.array_equal(np.array(rows, dtype=np.int64), golden["table"])


def test_factorize_restatement_matches_reference(golden):
    X = golden["X"]
    solver = golden["solver"]
    for k in golden["ks"]:
        merged = golden["merged_k%d" % k]
        rows = [r for r in golden["table"] if r[0] == k]
        for (kk, it, seed) in rows:
            W, H, n_it = nmf_ref.nmf(X, int(kk), int(seed), solver=solver, beta=golden["beta"], init=golden["init"])
            ref = merged[it * k:(it + 1) * k]
            assert rel_l2(H, ref) < 1e-10, (solver, k, it)


def test_trace_form_error_equals_dense_form(golden):
    # the CUDA path evaluates ||X-WH|| through the trace identity (SK/_nmf.py:116-120)
    X = golden["X"]
    W, H = nmf_ref.init_random(X.mean(), X.shape[0], X.shape[1], 5, 7)
    a = nmf_ref.frobenius_error(X, W, H)
    b = nmf_ref.frobenius_error_trace(X, W, H)
    assert abs(a - b) / a < 1e-12
    W2, H2, it2 = nmf_ref.mu_frobenius(X, W, H, error_fn=nmf_ref.frobenius_error_trace)
    W1, H1, it1 = nmf_ref.mu_frobenius(X, W, H)
    assert it1 == it2 and rel_l2(H2, H1) < 1e-12


def test_consensus_restatement_matches_reference(golden):
    solver = golden["solver"]
    for k in golden["ks"]:
        out = consensus_ref.consensus(golden["merged_k%d" % k], golden["X"], golden["tpm"],
                                      golden["tpm_std"], golden["hvg_idx"], int(k),
                                      density_threshold=float(golden["dt"]), solver=solver, beta=golden["beta"])
        # ||x||^2+||y||^2-2x.y cancels catastrophically for near-identical unit rows (d ~ 1e-4 here):
        # fp64 summation-order noise of 1e-16 in d^2 is 1e-8 relative -- hence rtol 1e-6, not 1e-12
        # (init='nndsvd': the restarts differ by ~1e-8 -- every density is that cancellation noise itself, compared as "zero")
        atol = 1e-12 if golden["init"] == "random" else 1e-7
        assert np.allclose(out["local_density"], golden["density_k%d" % k], rtol=1e-6, atol=atol)
        # the reference test's own criterion: sum of squared differences < 1e-4
        # (tests/test_reproducibility.py:111-112), plus a much tighter relative bound
        for name, key in (("consensus_spectra", "cspectra"), ("consensus_usages", "cusages"),
                          ("gene_spectra_tpm", "tpmspec"), ("gene_spectra_score", "score")):
            ref = golden["%s_k%d" % (key, k)]
            assert rel_l2(out[name], ref) < 1e-8, (name, k, rel_l2(out[name], ref))


def test_kmeans_restatement_matches_sklearn():
    from sklearn.cluster import KMeans
    rng = np.random.RandomState(3)
    centres = rng.rand(6, 40)
    X = np.vstack([c + 0.05 * rng.randn(30, 40) for c in centres])
    X = np.abs(X)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        km = KMeans(n_clusters=6, n_init=10, random_state=1).fit(X)
    labels, inertia, centers = consensus_ref.kmeans(X, 6)
    assert np.array_equal(labels, km.labels_)
    assert abs(inertia - km.inertia_) / km.inertia_ < 1e-10
    assert np.allclose(centers, km.cluster_centers_, atol=1e-12)


def test_distance_and_density_match_sklearn():
    from sklearn.metrics.pairwise import euclidean_distances
    rng = np.random.RandomState(0)
    S = consensus_ref.l2_normalize_rows(np.abs(rng.randn(50, 30)))
    D = euclidean_distances(S)
    assert np.allclose(consensus_ref.euclidean_distances(S), D, atol=1e-13)
    n = 7
    part = np.argpartition(D, n + 1)[:, :n + 1]
    dens = D[np.arange(50)[:, None], part].sum(1) / n      # cnmf.py:893-896
    assert np.allclose(consensus_ref.local_density(D, n), dens, atol=1e-13)


def test_silhouette_matches_sklearn():
    from sklearn.metrics import silhouette_score
    rng = np.random.RandomState(1)
    X = rng.rand(60, 10)
    labels = rng.randint(0, 4, 60)
    assert abs(consensus_ref.silhouette(X, labels) - silhouette_score(X, labels)) < 1e-12


def test_stats_branch_matches_reference(golden):
    # cnmf.py:922-936 (k_selection statistics): no density filter; silhouette + ||X - W H||^2
    solver = golden["solver"]
    for k in golden["ks"]:
        merged = golden["merged_k%d" % k]
        l2 = consensus_ref.l2_normalize_rows(merged)
        labels, _, _ = consensus_ref.kmeans(l2, int(k))
        med = consensus_ref.cluster_medians(l2, labels, int(k))
        rf, _ = nmf_ref.refit(golden["X"], med, solver, beta=golden["beta"])
        err = ((golden["X"] - rf @ med) ** 2).sum()
        stats = golden["stats_k%d" % k]
        # (init='nndsvd': intra-cluster distances are ~1e-8 cancellation noise, so the silhouette is 1 - noise)
        assert abs(consensus_ref.silhouette(l2, labels) - stats[2]) < (1e-9 if golden["init"] == "random" else 1e-7)
        assert abs(err - stats[3]) / stats[3] < 1e-9
