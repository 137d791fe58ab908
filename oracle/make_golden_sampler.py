"""TEST INFRASTRUCTURE. Golden vectors for ic_gan_b200/sampler.py from the LIVE reference (this container only):
imports /root/reference/data_utils/datasets_common.py (h5py is not installed here and is only touched by the HDF5 read
paths, so an empty stand-in module satisfies the import), builds an in-memory ILSVRC_HDF5_feats by attribute injection
(the pattern SURVEY.md section 8c proposes) and records what its samplers return for fixed numpy seeds.

    python oracle/make_golden_sampler.py   ->  tests/golden/sampler.npz"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def synth_tables(n=300, d=32, k=5, num_classes=7, seed=0):
    """The same synthetic tables the test regenerates: unit-norm features, labels, neighbour lists."""
    rs = np.random.RandomState(seed)
    feats = rs.randn(n, d).astype(np.float32)
    feats /= np.linalg.norm(feats, axis=1, keepdims=True)
    feats_hflip = rs.randn(n, d).astype(np.float32)
    feats_hflip /= np.linalg.norm(feats_hflip, axis=1, keepdims=True)
    labels = rs.randint(0, num_classes, size=n).astype(np.int64)
    labels[:num_classes] = np.arange(num_classes)  # every class present
    nns = rs.randint(0, n, size=(n, k)).astype(np.int64)
    return feats, feats_hflip, labels, nns


def main():
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    try:  # data_utils/resnet.py:45 imports a helper that newer torchvision moved to torch.hub
        import torchvision.models.utils  # noqa: F401
    except ImportError:
        shim = types.ModuleType("torchvision.models.utils")
        shim.load_state_dict_from_url = torch.hub.load_state_dict_from_url
        sys.modules["torchvision.models.utils"] = shim
    sys.path.insert(0, REF)
    from data_utils import datasets_common as dc  # noqa: E402  (the reference, unmodified)
    from data_utils import utils as du  # noqa: E402

    feats, feats_hflip, labels, nns = synth_tables()
    ds = object.__new__(dc.ILSVRC_HDF5_feats)
    ds.load_features = True
    ds.load_in_mem_feats = True
    ds.feats = torch.from_numpy(feats)
    ds.labels = labels
    ds.load_labels = True
    ds.label_onehot = False
    ds.load_in_mem_labels = True
    ds.sample_nns = [list(r) for r in nns]
    ds.possible_sampling_idxs = np.arange(len(feats))
    ds._feature_dim = feats.shape[1]
    out = {}
    np.random.seed(123)
    lab, f = ds.sample_conditioning_instance_balance(16)
    out["ib_labels"], out["ib_feats"] = lab.numpy(), f.numpy()
    w = np.linspace(1.0, 2.0, len(feats))
    w /= w.sum()
    np.random.seed(124)
    lab, f = ds.sample_conditioning_instance_balance(9, weights=w)
    out["ibw_labels"], out["ibw_feats"], out["ibw_weights"] = lab.numpy(), f.numpy(), w
    np.random.seed(125)
    lab, f = ds.sample_conditioning_nnclass_balance(12, weights=None, num_classes=7)
    out["nb_labels"], out["nb_feats"] = lab.numpy(), f.numpy()
    np.random.seed(126)
    lab, f = ds.sample_conditioning_nnclass_balance(12, weights=[1, 2, 3, 4, 3, 2, 1], num_classes=7)
    out["nbw_labels"], out["nbw_feats"] = lab.numpy(), f.numpy()
    # restricted instance set (the reference's kmeans / subsampling path only changes possible_sampling_idxs)
    ds.possible_sampling_idxs = np.array([3, 7, 11, 19, 42, 99])
    np.random.seed(127)
    lab, f = ds.sample_conditioning_instance_balance(10)
    out["sub_labels"], out["sub_feats"] = lab.numpy(), f.numpy()
    ds.possible_sampling_idxs = np.arange(len(feats))
    # no labels loaded: the neighbour draw still consumes the stream (datasets_common.py:562-563)
    ds.load_labels = False
    np.random.seed(128)
    lab, f = ds.sample_conditioning_instance_balance(6)
    assert lab is None
    out["nolab_feats"] = f.numpy()
    out["nolab_next_draw"] = np.array([np.random.randint(1 << 30)])
    ds.load_labels = True

    # dispatcher (data_utils/utils.py:830-901) with the reference's Distribution objects
    sys.path.insert(0, os.path.join(REF, "BigGAN_PyTorch"))
    z_ = torch.zeros(8, 4)
    z_.sample_ = lambda: None  # the dispatcher only calls sample_(); the noise itself is not part of this golden
    np.random.seed(129)
    z, lab, f = du.sample_conditioning_values(z_, None, batch_size=8, dataset=ds, class_cond=True, instance_cond=True,
                                              nn_sampling_strategy="instance_balance")
    out["disp_labels"], out["disp_feats"] = lab.numpy(), f.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sampler.npz"), **out)
    print("wrote tests/golden/sampler.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
