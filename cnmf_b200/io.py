"""On-disk codecs of the cNMF file ledger (host code).

``.df.npz``  = np.savez_compressed(data=values, index=index, columns=columns), read back with
               allow_pickle (reference cnmf.py:31-40) -- byte-layout compatible, so files written by
               either implementation are readable by the other.
``.txt``     = tab-separated DataFrame.to_csv (cnmf.py:34-35).
cells x genes matrices (``norm_counts``, ``tpm``): the reference stores AnnData ``.h5ad`` through
scanpy (cnmf.py:410,561).  anndata / h5py are optional here: with anndata installed real ``.h5ad``
files are read and written; without it the same matrix goes to ``<path>.npz`` (dense, with
obs / var names) and a note is printed once.
"""
import os
import warnings

import numpy as np
import pandas as pd


def save_df_to_npz(obj, filename):
    np.savez_compressed(filename, data=obj.values, index=obj.index.values, columns=obj.columns.values)


def save_df_to_text(obj, filename):
    obj.to_csv(filename, sep="\t")


def load_df_from_npz(filename):
    with np.load(filename, allow_pickle=True) as f:
        obj = pd.DataFrame(**f)
    return obj


class CellGeneMatrix:
    """Minimal cells x genes container (dense ndarray or scipy sparse) with obs / var names."""

    def __init__(self, X, obs_names, var_names):
        self.X = X
        self.obs_names = pd.Index(obs_names)
        self.var_names = pd.Index(var_names)

    @property
    def shape(self):
        return self.X.shape

    @property
    def is_sparse(self):
        return hasattr(self.X, "tocsr")

    def dense(self, dtype=np.float64):
        X = self.X.toarray() if hasattr(self.X, "toarray") else np.asarray(self.X)
        return np.ascontiguousarray(X, dtype=dtype)

    def subset_genes(self, names):
        idx = self.var_names.get_indexer(list(names))
        if (idx < 0).any():
            raise KeyError("genes not found: %s" % list(np.asarray(names)[idx < 0][:4]))
        return CellGeneMatrix(self.X[:, idx], self.obs_names, self.var_names[idx]), idx


def _have_anndata():
    try:
        import anndata  # noqa: F401
        return True
    except Exception:
        return False


_warned = False


def write_matrix(path, mat):
    global _warned
    if _have_anndata():
        import anndata
        ad = anndata.AnnData(X=mat.X, obs=pd.DataFrame(index=mat.obs_names), var=pd.DataFrame(index=mat.var_names))
        ad.write(path)
        return path
    if not _warned:
        warnings.warn("anndata is not installed: cells x genes matrices are stored as '<name>.h5ad.npz' "
                      "instead of '.h5ad'", UserWarning)
        _warned = True
    names = dict(obs=np.asarray(mat.obs_names, dtype=object), var=np.asarray(mat.var_names, dtype=object))
    if hasattr(mat.X, "tocsr"):       # sparse stays sparse (the reference keeps CSR unless --densify, cnmf.py:399-405)
        csr = mat.X.tocsr()
        np.savez(path + ".npz", csr_data=csr.data.astype(np.float64), csr_indices=csr.indices, csr_indptr=csr.indptr,
                 csr_shape=np.asarray(csr.shape), **names)
    else:
        np.savez(path + ".npz", X=mat.dense(np.float64), **names)
    return path + ".npz"


def read_matrix(path):
    if os.path.exists(path) and _have_anndata():
        import anndata
        ad = anndata.read_h5ad(path)
        return CellGeneMatrix(ad.X, ad.obs.index, ad.var.index)
    if os.path.exists(path + ".npz"):
        with np.load(path + ".npz", allow_pickle=True) as f:
            if "csr_data" in f:
                import scipy.sparse as sp
                X = sp.csr_matrix((f["csr_data"], f["csr_indices"], f["csr_indptr"]), shape=tuple(f["csr_shape"]))
                return CellGeneMatrix(X, f["obs"], f["var"])
            return CellGeneMatrix(f["X"], f["obs"], f["var"])
    if os.path.exists(path):
        raise RuntimeError("%s is an .h5ad file but anndata is not installed" % path)
    raise FileNotFoundError(path)


def read_counts(counts_fn):
    """Input counts as accepted by the reference's prepare() (cnmf.py:383-402): .h5ad, df.npz, or tab-delimited text.
    10x .mtx directories need scanpy and are not supported here."""
    if counts_fn.endswith(".h5ad"):
        return read_matrix(counts_fn)
    if counts_fn.endswith(".mtx") or counts_fn.endswith(".mtx.gz"):
        raise NotImplementedError("10x mtx input needs scanpy.read_10x_mtx, which is outside the accelerated path")
    if counts_fn.endswith(".npz"):
        df = load_df_from_npz(counts_fn)
    else:
        df = pd.read_csv(counts_fn, sep="\t", index_col=0)
    return CellGeneMatrix(df.values, df.index, df.columns)
