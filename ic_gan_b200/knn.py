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
               passes: int = 3, query_block: int = 0, timing: Optional[list] = None) -> KNNResult:
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
    if query_block <= 0:  # whole waves of the persistent sweep kernel: 128 query rows per CTA, one CTA per SM, two rounds
        query_block = 2 * 128 * torch.cuda.get_device_properties(dev).multi_processor_count
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
        if timing is not None:  # bench.py: CUDA events around the distance sweep on the launching stream
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        call("icgan_knn_coarse", ptr(hi), ptr(lo), ptr(norms), N, d, b0, b1, C, passes, ptr(ci), ptr(cd), stream_ptr())
        if timing is not None:
            ev[1].record()
            timing.append(ev)
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
    bad = torch.nonzero(flags).flatten().tolist()  # ONE host sync for the whole build; the launches below are asynchronous
    if bad:
        scratch = torch.empty(N, device=dev, dtype=torch.float64)
        for r in bad:
            call("icgan_knn_exact_row", ptr(X), N, d, q0 + r, k_nn, ptr(scratch), ptr(nn[r:]), ptr(rad[r:]), stream_ptr())
    stats = {"max_coarse_error": err, "margin": margin, "uncertified_rows": len(bad), "candidates": C, "passes": passes}
    return KNNResult(nn, rad, stats)


# ---------------------------------------------------------------------------------------------------------------------
# On-disk form of the neighbour tables.  make_hdf5_nns.run (data_utils/make_hdf5_nns.py:153-172)
# writes an HDF5 file with two datasets, `sample_nns` int64 [N, k] (chunks (chunk_size, k)) and `sample_nns_radius`
# float64 [N] (chunks (chunk_size,)), optional lzf compression; ILSVRC_HDF5_feats reads them back whole
# (data_utils/datasets_common.py:437-439).  With h5py importable the writer below produces exactly that file.  This
# image ships no h5py/HDF5 library, so the same two arrays -- same names, dtypes and shapes -- then go into a NumPy
# .npz next to the requested path (suffix .npz), and the reader accepts either.
def nns_filename(out_path, resolution=256, which_dataset="imagenet", split="train", test_part=False,
                 feature_extractor="selfsupervised", backbone_feature_extractor="resnet50", k_nn=50):
    """The reference's file name (make_hdf5_nns.py:140-150)."""
    prefix = {"imagenet": "ILSVRC", "imagenet_lt": "ILSVRC", "coco": "COCO"}.get(which_dataset, which_dataset)
    return "%s/%s%i%s%s%s_feats_%s_%s_nn_k%i.hdf5" % (
        out_path, prefix, resolution, "longtail" if which_dataset == "imagenet_lt" else "",
        "_val" if split == "val" else "", "_test" if test_part else "", feature_extractor, backbone_feature_extractor, k_nn)


def save_nns(path: str, sample_nns, sample_nns_radius, chunk_size: int = 500, compression=None) -> str:
    nns = np.ascontiguousarray(torch.as_tensor(sample_nns).cpu().numpy(), dtype=np.int64)
    rad = np.ascontiguousarray(torch.as_tensor(sample_nns_radius).cpu().numpy(), dtype=np.float64)
    if nns.ndim != 2 or rad.shape != (nns.shape[0],):
        raise ValueError("sample_nns must be [N, k] and sample_nns_radius [N]")
    try:
        import h5py
    except ImportError:
        h5py = None
    if h5py is not None:
        with h5py.File(path, "w") as f:
            c = max(1, min(chunk_size, nns.shape[0]))
            f.create_dataset("sample_nns", nns.shape, dtype="int64", maxshape=nns.shape, chunks=(c, nns.shape[1]),
                             compression=compression)[...] = nns
            f.create_dataset("sample_nns_radius", rad.shape, dtype="float", maxshape=rad.shape, chunks=(c,),
                             compression=compression)[...] = rad
        return path
    out = path + ".npz"
    with open(out, "wb") as f:
        np.savez(f, sample_nns=nns, sample_nns_radius=rad)
    return out


def load_nns(path: str):
    """-> (sample_nns int64 [N,k], sample_nns_radius float64 [N]) from the HDF5 file or its .npz stand-in."""
    import os
    if os.path.exists(path) and not path.endswith(".npz"):
        import h5py  # an .hdf5 file can only have been written with h5py present
        with h5py.File(path, "r") as f:
            return f["sample_nns"][:], f["sample_nns_radius"][:]
    data = np.load(path if path.endswith(".npz") else path + ".npz")
    return data["sample_nns"].astype(np.int64), data["sample_nns_radius"].astype(np.float64)
