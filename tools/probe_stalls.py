#!/usr/bin/env python
"""GPU probe: which library call do the sporadic ~70 ms stalls of the consensus stage sit in?  Wraps every C-ABI entry point
with a wall-clock timer and lists the calls that took longer than 10 ms during the bench's consensus arm."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from cnmf_b200 import _lib, consensus as cs
from cnmf_b200.engine import Engine
from cnmf_b200.synth import restart_table

lib = _lib.load()
slow = []

class Timed:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn
    def __call__(self, *a):
        t0 = time.perf_counter()
        r = self.fn(*a)
        dt = 1e3 * (time.perf_counter() - t0)
        if dt > 10.0:
            slow.append((self.name, round(dt, 1)))
        return r

class Proxy:
    def __init__(self, lib):
        self._lib = lib
        self._cache = {}
    def __getattr__(self, name):
        if name not in self._cache:
            self._cache[name] = Timed(name, getattr(self._lib, name))
        return self._cache[name]

X, tpm, tpm_std = bench.make_data("c3", want_tpm=True)
eng = Engine(0)
eng.lib = Proxy(lib)
ds = eng.dataset(X); tds = eng.dataset(tpm)
import gc
if os.environ.get("PROBE_GC_FREEZE", "1") == "1":
    gc.collect(); gc.freeze()
for k in (7, 9, 11, 13):
    rows = restart_table([k], 100, seed=14)
    sp, _, _, _ = ds.factorize([r[0] for r in rows], [r[2] for r in rows], bench.NMF_KW)
    merged = np.vstack(sp)
    for rep in range(2):
        slow.clear()
        t0 = time.perf_counter()
        tt = []
        res = cs.consensus_numerics(eng, merged, k, ds, bench.NMF_KW, tpm_ds=tds, hvg_idx=np.arange(X.shape[1]), tpm_std_hvg=tpm_std)
        torch.cuda.synchronize()
        print(json.dumps({"k": k, "rep": rep, "ms": round(1e3 * (time.perf_counter() - t0), 1), "slow_calls": slow,
                          "phases": {a: round(b, 1) for a, b in cs.STATS["phases_ms"].items()}}), flush=True)
        cs.STATS.clear()
