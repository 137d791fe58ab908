"""Conditioning sampler (ic_gan_b200/sampler.py) against golden vectors recorded from the live reference
(oracle/make_golden_sampler.py): for a given np.random.seed the selected instances, neighbour labels and gathered
feature rows are bit-identical to data_utils/datasets_common.py:525-622 and data_utils/utils.py:830-901."""
import os

import numpy as np
import pytest
import torch

from ic_gan_b200.sampler import ConditioningSampler, sample_conditioning_values
from oracle.make_golden_sampler import synth_tables

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sampler.npz")


@pytest.fixture(scope="module")
def tables():
    feats, feats_hflip, labels, nns = synth_tables()
    return torch.from_numpy(feats), torch.from_numpy(feats_hflip), labels, nns


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


def _eq(t, ref):
    assert t is not None
    got = t.cpu().numpy()
    assert got.shape == ref.shape and got.dtype == ref.dtype, (got.shape, got.dtype, ref.shape, ref.dtype)
    assert np.array_equal(got, ref)


def test_instance_balance_matches_reference(tables, gold):
    feats, _, labels, nns = tables
    s = ConditioningSampler(feats, nns, labels)
    np.random.seed(123)
    lab, f = s.sample_conditioning_instance_balance(16)
    _eq(lab, gold["ib_labels"])
    _eq(f, gold["ib_feats"])
    np.random.seed(124)
    lab, f = s.sample_conditioning_instance_balance(9, weights=gold["ibw_weights"])
    _eq(lab, gold["ibw_labels"])
    _eq(f, gold["ibw_feats"])


def test_nnclass_balance_matches_reference(tables, gold):
    feats, _, labels, nns = tables
    s = ConditioningSampler(feats, nns, labels)
    np.random.seed(125)
    lab, f = s.sample_conditioning_nnclass_balance(12, weights=None, num_classes=7)
    _eq(lab, gold["nb_labels"])
    _eq(f, gold["nb_feats"])
    np.random.seed(126)
    lab, f = s.sample_conditioning_nnclass_balance(12, weights=[1, 2, 3, 4, 3, 2, 1], num_classes=7)
    _eq(lab, gold["nbw_labels"])
    _eq(f, gold["nbw_feats"])


def test_restricted_instances_and_label_free_stream(tables, gold):
    feats, _, labels, nns = tables
    s = ConditioningSampler(feats, nns, labels, possible_sampling_idxs=[3, 7, 11, 19, 42, 99])
    np.random.seed(127)
    lab, f = s.sample_conditioning_instance_balance(10)
    _eq(lab, gold["sub_labels"])
    _eq(f, gold["sub_feats"])
    # without labels the neighbour draws still advance the numpy stream exactly as the reference does
    s2 = ConditioningSampler(feats, nns, labels=None)
    np.random.seed(128)
    lab, f = s2.sample_conditioning_instance_balance(6)
    assert lab is None
    _eq(f, gold["nolab_feats"])
    assert np.random.randint(1 << 30) == int(gold["nolab_next_draw"][0])


def test_dispatcher_matches_reference(tables, gold):
    feats, _, labels, nns = tables
    s = ConditioningSampler(feats, nns, labels)
    z_ = torch.zeros(8, 4)
    z_.sample_ = lambda: None
    np.random.seed(129)
    z, lab, f = sample_conditioning_values(z_, None, batch_size=8, dataset=s, class_cond=True, instance_cond=True,
                                           nn_sampling_strategy="instance_balance")
    assert z is z_
    _eq(lab, gold["disp_labels"])
    _eq(f, gold["disp_feats"])
    np.random.seed(129)
    z, f2 = sample_conditioning_values(z_, None, batch_size=8, dataset=s, class_cond=False, instance_cond=True)
    _eq(f2, gold["disp_feats"])
    with pytest.raises(ValueError):
        sample_conditioning_values(z_, None, batch_size=8, dataset=s, instance_cond=True, nn_sampling_strategy="nope")


def test_hflip_augmentation_follows_the_hdf5_read_path(tables):
    """datasets_common.py:655-672: one np.random.randint(2) per gathered instance; a 1 selects the flipped-image
    features when feature_augmentation is on.  Rows must come from the matching table and the draw count must be one
    per instance (the in-memory path draws nothing)."""
    feats, feats_hflip, labels, nns = tables
    s = ConditioningSampler(feats, nns, labels, feats_hflip=feats_hflip, feature_augmentation=True, draw_hflip=True)
    idx = np.array([5, 9, 5, 200, 17, 17, 3])
    np.random.seed(7)
    flips = np.array([np.random.randint(2) == 1 for _ in idx])
    after = np.random.randint(1 << 30)
    np.random.seed(7)
    out = s.get_instance_features(idx)
    assert np.random.randint(1 << 30) == after
    assert flips.any() and not flips.all()
    want = torch.where(torch.from_numpy(flips)[:, None], feats_hflip[idx], feats[idx])
    assert torch.equal(out, want)
    plain = ConditioningSampler(feats, nns, labels)
    np.random.seed(7)
    plain.get_instance_features(idx)
    np.random.seed(7)
    first = np.random.randint(1 << 30)
    np.random.seed(7)
    plain.get_instance_features(idx)
    assert np.random.randint(1 << 30) == first  # no draws consumed
    with pytest.raises(ValueError):
        ConditioningSampler(feats, nns, labels, feature_augmentation=True)
    with pytest.raises(ValueError):
        ConditioningSampler(feats, nns[:10], labels)


def test_normalize_option_and_device_residency(tables):
    feats, _, labels, nns = tables
    s = ConditioningSampler(feats * 3.0, nns, labels, normalize=True)
    np.random.seed(1)
    _, f = s.sample_conditioning_instance_balance(32)
    assert torch.allclose(f.norm(dim=1), torch.ones(32), atol=1e-6)
    assert f.device == s.device and f.dtype == torch.float32


@pytest.mark.gpu
def test_gpu_resident_tables_match_reference(cuda_device, gold):
    """Row a26 on the device: tables resident on cuda:0, draws in the reference's numpy order, gathers on the GPU; the
    results (labels, feature rows) are bit-identical to the live reference's golden vectors and stay on the device."""
    feats, feats_hflip, labels, nns = synth_tables()
    s = ConditioningSampler(torch.from_numpy(feats).to(cuda_device), nns, labels)
    assert s.nns_dev is not None and s.nns_dev.is_cuda
    np.random.seed(123)
    lab, f = s.sample_conditioning_instance_balance(16)
    assert lab.is_cuda and f.is_cuda
    _eq(lab, gold["ib_labels"]); _eq(f, gold["ib_feats"])
    np.random.seed(125)
    lab, f = s.sample_conditioning_nnclass_balance(12, weights=None, num_classes=7)
    _eq(lab, gold["nb_labels"]); _eq(f, gold["nb_feats"])
    z_ = torch.zeros(8, 4, device=cuda_device)
    z_.sample_ = lambda: None
    np.random.seed(129)
    z, lab, f = sample_conditioning_values(z_, None, batch_size=8, dataset=s, class_cond=True, instance_cond=True)
    _eq(lab, gold["disp_labels"]); _eq(f, gold["disp_feats"])


def test_ragged_neighbour_rows_use_per_instance_draws(tables):
    feats, _, labels, nns = tables
    ragged = [np.asarray(r)[: 3 + (i % 5)] for i, r in enumerate(nns)]
    s = ConditioningSampler(feats, ragged, labels)
    assert s.nns_dev is None
    np.random.seed(3)
    sel = np.random.randint(0, len(ragged), size=6)
    want = [np.random.choice(ragged[i]) for i in sel]
    np.random.seed(3)
    lab, _ = s.sample_conditioning_instance_balance(6)
    assert lab.tolist() == [int(labels[j]) for j in want]
