"""k-NN instance-conditioning build on B200 (host side).

Mirrors ``ILSVRC_HDF5_feats._obtain_nns`` (data_utils/datasets_common.py:695-745) and the arrays ``make_hdf5_nns.run``
writes (data_utils/make_hdf5_nns.py:132-172): ``sample_nns`` int64 [N, k] and ``sample_nns_radius`` float64 [N].
Query rows shard across ranks (database replicated) with no collective during the search.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from ._lib import call, ptr, stream_ptr


def normalize_features(feats: np.ndarray) -> np.ndarray:
    """float64 L2-normalisation then the float32 cast Faiss receives (datasets_common.py:422-428, :726-729)."""
    f = np.asarray(feats, dtype=np.float64)
    f = f / np.linalg.norm(f, axis=1, keepdims=True)
    return f.astype(np.float32)


class KNNResult:
    def __init__(self, nns, radii, stats):
        self.sample_nns, self.sample_nns_radius, self.stats = nns, radii, stats


def obtain_nns(feats: torch.Tensor, k_nn: int = 50, rows: Optional[Tuple[int, int]] = None, candidates: int = 64,
               passes: int = 3, query_block: int = 32768) -> KNNResult:
    """Exact k nearest neighbours (squared L2, self excluded) of feats[rows] within feats ([N, d] float32 CUDA tensor).

    Returns int64 [n, k_nn] indices ordered by (distance, index) and float64 [n] radii.  Every row is certified exact by
    the safety-margin test of icgan_knn_rerank or recomputed by the float64 brute-force kernel."""
    assert feats.is_cuda and feats.dtype == torch.float32 and feats.dim() == 2
    X = feats.contiguous()
    N, d = X.shape
    q0, q1 = rows if rows is not None else (0, N)
    C = min(candidates, 64)
    if not (1 <= k_nn < C):
        raise ValueError("need 1 <= k_nn < candidates <= 64")
    dev = X.device
    hi = torch.empty(N, d, device=dev, dtype=torch.bfloat16)
    lo = torch.empty(N, d, device=dev, dtype=torch.bfloat16)
    norms = torch.empty(N, device=dev, dtype=torch.float32)
    call("icgan_knn_prepare", ptr(X), ptr(hi), ptr(lo), ptr(norms), N, d, stream_ptr())
    nn = torch.empty(q1 - q0, k_nn, device=dev, dtype=torch.int64)
    rad = torch.empty(q1 - q0, device=dev, dtype=torch.float64)
    flags = torch.empty(q1 - q0, device=dev, dtype=torch.int32)
    max_err = torch.zeros(1, device=dev, dtype=torch.float32)
    margin = 1e-4 if passes == 3 else 2e-2
    blocks = []
    for b0 in range(q0, q1, query_block):
        b1 = min(q1, b0 + query_block)
        ci = torch.empty(b1 - b0, C, device=dev, dtype=torch.int32)
        cd = torch.empty(b1 - b0, C, device=dev, dtype=torch.float32)
        call("icgan_knn_coarse", ptr(hi), ptr(lo), ptr(norms), N, d, b0, b1, C, passes, ptr(ci), ptr(cd), stream_ptr())
        blocks.append((b0, b1, ci, cd))
    for attempt in range(3):
        max_err.zero_()
        for b0, b1, ci, cd in blocks:
            o = b0 - q0
            call("icgan_knn_rerank", ptr(X), N, d, b0, b1, C, k_nn, ptr(ci), ptr(cd), ptr(nn[o:]), ptr(rad[o:]),
                 ptr(flags[o:]), ptr(max_err), float(margin), stream_ptr())
        err = float(max_err.item())
        if 2.0 * err <= margin:
            break
        margin = 4.0 * err  # the coarse pass was less accurate than assumed: re-certify with a wider margin
    bad = torch.nonzero(flags).flatten().tolist()
    if bad:
        scratch = torch.empty(N, device=dev, dtype=torch.float64)
        for r in bad:
            call("icgan_knn_exact_row", ptr(X), N, d, q0 + r, k_nn, ptr(scratch), ptr(nn[r:]), ptr(rad[r:]), stream_ptr())
    stats = {"max_coarse_error": err, "margin": margin, "uncertified_rows": len(bad), "candidates": C, "passes": passes}
    return KNNResult(nn, rad, stats)
