#!/usr/bin/env python
"""GPU probe: where the consensus stage spends its time (per phase), at c3 (K=6, 9, 13) and at c4 size (R=4000, K=20)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from cnmf_b200 import consensus as cs
from cnmf_b200.engine import Engine
from cnmf_b200.synth import restart_table

def run(name, ks, n_iter, eng, reps=2, n_cells=None):
    if n_cells:
        bench.WORKLOADS[name] = dict(bench.WORKLOADS["c3"], n_cells=n_cells)
    X, tpm, tpm_std = bench.make_data(name, want_tpm=True)
    ds = eng.dataset(X)
    tds = eng.dataset(tpm)
    for k in ks:
        rows = restart_table([k], n_iter, seed=14)
        t0 = time.perf_counter()
        sp, _, it, _ = ds.factorize([r[0] for r in rows], [r[2] for r in rows], bench.NMF_KW)
        tf = time.perf_counter() - t0
        merged = np.vstack(sp)
        for rep in range(reps):
            cs.STATS.clear()
            t0 = time.perf_counter()
            res = cs.consensus_numerics(eng, merged, k, ds, bench.NMF_KW, tpm_ds=tds, hvg_idx=np.arange(X.shape[1]), tpm_std_hvg=tpm_std)
            torch.cuda.synchronize()
            tot = 1e3 * (time.perf_counter() - t0)
        # phase detail: KMeans (seeding / Lloyd / final) and the three refits (setup / solve), library-side timers
        eng.profile(True)
        S = cs.SpectraMatrix(eng, merged).l2_normalize()
        t0 = time.perf_counter(); cs.kmeans(S, k); km_ms = 1e3 * (time.perf_counter() - t0)
        km = eng.last_timing()
        detail = {"kmeans_ms": round(km_ms, 2), "km_seed_ms": round(km["rng_ms"], 2), "km_lloyd_ms": round(km["solve_ms"], 2),
                  "km_lloyd_iters_max": km["h2d_ms"], "km_final_ms": round(km["d2h_ms"], 2)}
        t0 = time.perf_counter(); W, it_a, _ = ds.refit(res["median_spectra"], bench.NMF_KW); ra = 1e3 * (time.perf_counter() - t0)
        ta = eng.last_timing()
        U = res["norm_usages"]
        t0 = time.perf_counter(); Ht, it_b, _ = tds.refit(np.ascontiguousarray(U.T), bench.NMF_KW, transposed=True); rb = 1e3 * (time.perf_counter() - t0)
        tb = eng.last_timing()
        detail.update(refit_a_ms=round(ra, 2), refit_a_setup=round(ta["h2d_ms"], 2), refit_a_solve=round(ta["solve_ms"], 2), refit_a_it=it_a,
                      refit_b_ms=round(rb, 2), refit_b_setup=round(tb["h2d_ms"], 2), refit_b_solve=round(tb["solve_ms"], 2), refit_b_it=it_b)
        eng.profile(False)
        print(json.dumps({"case": name, "detail": detail, "N": X.shape[0], "k": k, "R": merged.shape[0], "factorize_s": round(tf, 3), "consensus_ms": round(tot, 1),
                          "phases_ms": {a: round(b, 1) for a, b in cs.STATS["phases_ms"].items()},
                          "lloyd_iters": cs.STATS.get("lloyd_iters"), "refits": [list(map(int, r)) for r in cs.STATS.get("refits", [])]}), flush=True)
    ds.close(); tds.close()

eng = Engine(0)
run("c3", [6, 9, 13], 100, eng)
run("c4", [20], 200, eng, n_cells=68000)
