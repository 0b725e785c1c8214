#!/usr/bin/env python
"""Multi-GPU check (run under torchrun on a box with >= 2 GPUs; not a pytest file):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tests/gpu_dist_check.py

Every rank factorizes its share of the fixture's restarts (idx % world == rank), the spectra are
all-gathered over NCCL, and every rank checks the merged result against the reference's merged spectra."""
import os
import sys
import tempfile
import warnings

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cnmf_golden import load_golden  # noqa: E402
from cnmf_b200 import cNMF, save_df_to_npz  # noqa: E402
from cnmf_b200.parallel import dist_info, factorize_distributed, init_process_group  # noqa: E402


def main():
    rank, world, local = dist_info()
    dist = init_process_group("nccl")
    g = load_golden("sim_mu")
    base = os.path.join(tempfile.gettempdir(), "cnmf_dist_check")
    if rank == 0:
        os.makedirs(base, exist_ok=True)
        counts = g["counts"].astype(np.float64)
        df = pd.DataFrame(counts, index=["c%d" % i for i in range(counts.shape[0])],
                          columns=["g%d" % i for i in range(counts.shape[1])])
        save_df_to_npz(df, os.path.join(base, "counts.df.npz"))
        obj = cNMF(output_dir=base, name="run")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            obj.prepare(os.path.join(base, "counts.df.npz"), components=list(g["ks"]), n_iter=int(g["n_iter"]),
                        seed=int(g["seed"]), beta_loss=2.0, num_highvar_genes=len(g["hvg_idx"]), densify=True)
    if dist is not None:
        dist.barrier()
    obj = cNMF(output_dir=base, name="run", device=local)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        merged = factorize_distributed(obj)
    worst = 0.0
    for k in g["ks"]:
        ref = g["merged_k%d" % k]
        got = merged[int(k)].values
        for it in range(ref.shape[0] // k):
            e = np.linalg.norm(got[it * k:(it + 1) * k] - ref[it * k:(it + 1) * k]) / np.linalg.norm(ref[it * k:(it + 1) * k])
            lim = 1e-3 if (int(k), it) == (4, 0) else 1e-4     # the ill-conditioned restart named in test_gpu_parity.py
            assert e < lim, (k, it, e)
            worst = max(worst, e)
    print("rank %d/%d: merged spectra match the reference fixture (worst rel-L2 %.2e)" % (rank, world, worst), flush=True)
    # consensus sweep sharded K -> GPU, fed from the device slab the all-gather left behind (no files, no H2D)
    import yaml
    from cnmf_b200 import consensus as cs
    from cnmf_b200 import io as cio
    from cnmf_b200.parallel import consensus_ks_of_rank
    sharded = obj.last_sharded_spectra
    table = g["table"]
    kw = yaml.load(open(obj.paths["nmf_run_parameters"]), Loader=yaml.FullLoader)
    norm = cio.read_matrix(obj.paths["normalized_counts"])
    norm_ds = obj._dataset(norm.X)
    for k in consensus_ks_of_rank(sorted(int(x) for x in g["ks"]), rank, world):
        jobs = [j for j in range(len(table)) if table[j, 0] == k]
        res = cs.consensus_numerics(obj.engine(), sharded.matrix(obj.engine(), jobs), k, norm_ds, kw,
                                    density_threshold=float(g["dt"]))
        e = np.linalg.norm(res["median_spectra"] - g["cspectra_k%d" % k]) / np.linalg.norm(g["cspectra_k%d" % k])
        assert e < (3e-4 if k == 4 else 1e-4), (k, e)
        print("rank %d/%d: consensus(k=%d) from the device slab matches the reference (rel-L2 %.2e)" % (rank, world, k, e), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
