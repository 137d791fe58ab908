"""k-NN conditioning build: oracle vs golden (CPU) and the B200 kernels vs oracle / golden (GPU), bit-exact indices."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import knn_oracle as K
from tests.helpers import GOLD


def _golden_features():
    meta = json.load(open(os.path.join(GOLD, "knn_n1500_k50.json")))
    rng = np.random.default_rng(meta["seed"])
    raw = rng.standard_normal((meta["n"], meta["d"]))
    raw[7] = raw[3]
    raw[900] = raw[901] * 2
    data = np.load(os.path.join(GOLD, "knn_n1500_k50.npz"))
    return meta, K.normalize_features(raw), data


def test_oracle_matches_golden():
    meta, x32, data = _golden_features()
    nns, radii = K.obtain_nns(x32, meta["k"])
    assert np.array_equal(nns, data["nns"].astype(np.int64))
    assert np.array_equal(radii, data["radii"])
    # the reference's sklearn path reports the float64 k-th distance; ours is the float32 sqrt Faiss would return
    assert np.allclose(radii, data["ref_radius_sklearn"], rtol=0, atol=1e-6)
    assert nns[3, 0] == 7 and nns[7, 0] == 3  # exact duplicates are each other's nearest neighbour at distance 0
    assert all(i not in nns[i] for i in range(meta["n"]))


def test_oracle_edge_cases():
    rng = np.random.default_rng(0)
    x = K.normalize_features(rng.standard_normal((40, 64)))
    nns, radii = K.obtain_nns(x, 5)
    d = ((x[:, None, :].astype(np.float64) - x[None, :, :].astype(np.float64)) ** 2).sum(-1)
    for i in range(40):
        order = np.lexsort((np.arange(40), d[i]))
        expect = [j for j in order[:6] if j != i][:5]
        assert nns[i].tolist() == expect
        assert radii[i] == float(np.sqrt(np.float32(d[i][order[5]])))


@pytest.mark.gpu
def test_gpu_knn_matches_golden_bit_exact(cuda_device):
    from ic_gan_b200 import knn
    meta, x32, data = _golden_features()
    res = knn.obtain_nns(torch.from_numpy(x32).to(cuda_device), meta["k"])
    print("knn stats", res.stats)
    assert np.array_equal(res.sample_nns.cpu().numpy(), data["nns"].astype(np.int64))
    assert np.array_equal(res.sample_nns_radius.cpu().numpy(), data["radii"])


@pytest.mark.gpu
@pytest.mark.parametrize("n,d,k,passes", [(4097, 2048, 50, 3), (3000, 512, 20, 3), (2500, 2048, 50, 1), (300, 128, 10, 3)])
def test_gpu_knn_matches_oracle(cuda_device, n, d, k, passes):
    """ragged sizes (N not a multiple of the 128/256 tiles), clustered data with near-ties, sharded query ranges."""
    from ic_gan_b200 import knn
    rng = np.random.default_rng(n)
    centers = rng.standard_normal((8, d))
    raw = centers[rng.integers(0, 8, n)] * 0.7 + rng.standard_normal((n, d))
    raw[n // 2] = raw[n // 2 - 1]
    x32 = K.normalize_features(raw)
    nns, radii = K.obtain_nns(x32, k)
    X = torch.from_numpy(x32).to(cuda_device)
    res = knn.obtain_nns(X, k, passes=passes)
    print("knn stats", res.stats)
    assert np.array_equal(res.sample_nns.cpu().numpy(), nns)
    assert np.array_equal(res.sample_nns_radius.cpu().numpy(), radii)
    # query sharding (the multi-GPU partition): rows [a, b) computed alone give the same answer
    a, b = n // 3, n // 3 + 777 if n > 1200 else n // 3 + 50
    part = knn.obtain_nns(X, k, rows=(a, b), passes=passes)
    assert np.array_equal(part.sample_nns.cpu().numpy(), nns[a:b])


@pytest.mark.gpu
def test_gpu_knn_properties_at_scale(cuda_device):
    """size the oracle cannot reach quickly: self never listed, radii non-decreasing vs listed distances, exact re-check."""
    from ic_gan_b200 import knn
    n, d, k = 60000, 2048, 50
    g = torch.Generator(device=cuda_device).manual_seed(6)
    X = torch.nn.functional.normalize(torch.randn(n, d, device=cuda_device, generator=g, dtype=torch.float64), dim=1).float()
    res = knn.obtain_nns(X, k)
    print("knn stats", res.stats)
    nn = res.sample_nns
    assert int((nn == torch.arange(n, device=cuda_device)[:, None]).sum()) == 0
    assert int((nn < 0).sum()) == 0
    rows = torch.randint(0, n, (64,), device=cuda_device)
    d2 = ((X[rows].double()[:, None, :] - X.double()[None, :, :][:, :, :]) ** 2).sum(-1) if False else None
    for r in rows.tolist()[:16]:
        dist = ((X[r].double()[None, :] - X.double()) ** 2).sum(1)
        order = torch.argsort(dist, stable=True)[: k + 1]
        expect = [j for j in order.tolist() if j != r][:k]
        assert nn[r].tolist() == expect


def test_nns_file_roundtrip(tmp_path):
    """sample_nns / sample_nns_radius on disk: the reference's dataset names, dtypes and shapes (make_hdf5_nns.py:153-172)."""
    from ic_gan_b200 import knn
    rng = np.random.default_rng(1)
    nns = rng.integers(0, 1000, (37, 50))
    rad = rng.random(37)
    name = knn.nns_filename(str(tmp_path), 256, "imagenet", "train", False, "selfsupervised", "resnet50", 50)
    assert name.endswith("ILSVRC256_feats_selfsupervised_resnet50_nn_k50.hdf5")
    out = knn.save_nns(name, torch.from_numpy(nns), torch.from_numpy(rad))
    a, b = knn.load_nns(name)
    assert a.dtype == np.int64 and b.dtype == np.float64 and a.shape == (37, 50) and b.shape == (37,)
    assert np.array_equal(a, nns) and np.array_equal(b, rad) and os.path.exists(out)


def test_oracle_rows_agree_with_full_oracle():
    meta, x32, data = _golden_features()
    rows = [0, 3, 7, 900, 901, 1499]
    nns, radii = K.obtain_nns_rows(x32, rows, meta["k"])
    assert np.array_equal(nns, data["nns"].astype(np.int64)[rows]) and np.array_equal(radii, data["radii"][rows])


@pytest.mark.gpu
def test_gpu_knn_100k_rows_match_oracle(cuda_device):
    """BASELINE config 5's oracle-checkable size: N = 100 000, d = 2048, k = 50; every row certified or recomputed on the
    device, 48 query rows (first/last tiles + random) bit-exact against the float64 oracle."""
    from ic_gan_b200 import knn
    n, d, k = 100_000, 2048, 50
    g = torch.Generator().manual_seed(6)
    x = torch.randn(n, d, generator=g)
    x32 = K.normalize_features(x.numpy())
    res = knn.obtain_nns(torch.from_numpy(x32).to(cuda_device), k)
    print("knn 100k stats", res.stats)
    rows = np.concatenate([np.arange(8), np.arange(n - 8, n), np.random.default_rng(0).integers(0, n, 32)])
    want_nn, want_r = K.obtain_nns_rows(x32, rows, k)
    got_nn = res.sample_nns.cpu().numpy()[rows]
    got_r = res.sample_nns_radius.cpu().numpy()[rows]
    assert np.array_equal(got_nn, want_nn)
    assert np.array_equal(got_r, want_r)
