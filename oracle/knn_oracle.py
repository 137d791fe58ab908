"""CPU oracle for the k-NN conditioning build — TEST INFRASTRUCTURE ONLY (never imported by ic_gan_b200/).

Restates ILSVRC_HDF5_feats._obtain_nns (data_utils/datasets_common.py:695-745) + make_hdf5_nns.run
(data_utils/make_hdf5_nns.py:132-133).  The reference delegates the search to Faiss ``IndexFlatL2`` (un-vendored
dependency, pinned faiss-gpu=1.7.0 in environment.yml:19,35; not installable here), whose published algorithm is exact
brute-force squared-L2 search returning the k+1 smallest distances in ascending order.  Faiss leaves rounding/tie order
to its SGEMM; this oracle fixes them: distances in float64 from the float32 features, ties broken by the lower index.
PARITY PINNING: against the reference's own in-tree sklearn fallback path (same function, faiss_lib=False), which
yields the neighbour SET per row (unordered) — checked in oracle/make_golden_extra.py; the order is pinned by definition.
"""
from __future__ import annotations

import numpy as np


def normalize_features(feats):
    """datasets_common.py:422-428: float64 normalise; :726-729: float32 cast handed to the index."""
    f = np.asarray(feats, dtype=np.float64)
    f = f / np.linalg.norm(f, axis=1, keepdims=True)
    return f.astype(np.float32)


def obtain_nns(x32: np.ndarray, k_nn: int, block: int = 512):
    """-> (sample_nns int64 [N,k], sample_nns_radius float64 [N]) for float32 features x32 [N,d]."""
    x32 = np.ascontiguousarray(x32, dtype=np.float32)
    n, _ = x32.shape
    x = x32.astype(np.float64)
    sq = (x * x).sum(1)
    kk = k_nn + 1  # the query itself is its own 0-NN (datasets_common.py:714-716)
    nns = np.full((n, k_nn), -1, dtype=np.int64)
    radii = np.zeros(n, dtype=np.float64)
    extra = 16
    for b0 in range(0, n, block):
        b1 = min(n, b0 + block)
        d2 = sq[b0:b1, None] + sq[None, :] - 2.0 * (x[b0:b1] @ x.T)  # float64 shortlist
        m = min(n, kk + extra)
        short = np.argpartition(d2, m - 1, axis=1)[:, :m]
        for r in range(b1 - b0):
            cand = short[r]
            diff = x[b0 + r][None, :] - x[cand]
            ex = (diff * diff).sum(1)  # exact difference form
            order = np.lexsort((cand, ex))  # by distance, ties -> lower index
            cand, ex = cand[order][:kk], ex[order][:kk]
            keep = cand[cand != (b0 + r)]  # drop the query by value (datasets_common.py:739-743)
            nns[b0 + r, :min(k_nn, keep.size)] = keep[:k_nn]
            radii[b0 + r] = float(np.sqrt(np.float32(ex[min(kk, ex.size) - 1])))  # float32 sqrt as Faiss' output
    return nns, radii


def obtain_nns_rows(x32: np.ndarray, rows, k_nn: int):
    """The same answer for a subset of query rows only (exact float64 difference form against the whole database):
    lets tests and bench.py check builds whose full oracle would take hours (N = 100 k ... 1.28 M)."""
    x32 = np.ascontiguousarray(x32, dtype=np.float32)
    rows = np.asarray(rows, dtype=np.int64)
    nns = np.full((rows.size, k_nn), -1, dtype=np.int64)
    radii = np.zeros(rows.size, dtype=np.float64)
    n = x32.shape[0]
    kk = k_nn + 1
    for o, r in enumerate(rows):
        q = x32[r].astype(np.float64)
        ex = np.empty(n, dtype=np.float64)
        for c0 in range(0, n, 65536):  # chunked: no [N, d] float64 temporary
            diff = x32[c0:c0 + 65536].astype(np.float64) - q[None, :]
            ex[c0:c0 + 65536] = (diff * diff).sum(1)
        m = min(n, kk + 16)
        cand = np.argpartition(ex, m - 1)[:m]
        order = np.lexsort((cand, ex[cand]))
        cand = cand[order][:kk]
        keep = cand[cand != r]
        nns[o, :min(k_nn, keep.size)] = keep[:k_nn]
        radii[o] = float(np.sqrt(np.float32(ex[cand[min(kk, cand.size) - 1]])))
    return nns, radii
