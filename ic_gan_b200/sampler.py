"""Device-resident conditioning sampler: the step just before the G/D hot path.

Restates, on tables that live on the GPU, what the reference does on the host with numpy + HDF5 reads every step:
  * ILSVRC_HDF5_feats.sample_conditioning_instance_balance   data_utils/datasets_common.py:525-576
  * ILSVRC_HDF5_feats.sample_conditioning_nnclass_balance    data_utils/datasets_common.py:578-622
  * ILSVRC_HDF5_feats.get_instance_features                  data_utils/datasets_common.py:647-679
  * data_utils.utils.sample_conditioning_values              data_utils/utils.py:830-901 (the dispatch on class_cond /
    instance_cond / nn_sampling_strategy)

The index draws use numpy's legacy global stream in exactly the reference's order (np.random.randint / np.random.choice,
one call per instance), so for a given np.random.seed the chosen instances, neighbours and labels are identical to the
reference's; only the row gathers (features, labels) happen on the device, and the results stay there -- no per-step
HDF5 open, no host->device copy of [B, 2048] features.  Normalised features are expected (the reference normalises at
load time for in-memory tables and per read otherwise; pass `normalize=True` to do it here)."""
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor


class ConditioningSampler:
    def __init__(self, feats: Tensor, sample_nns, labels: Optional[Sequence[int]] = None,
                 possible_sampling_idxs: Optional[Sequence[int]] = None, feats_hflip: Optional[Tensor] = None,
                 feature_augmentation: bool = False, draw_hflip: bool = False, normalize: bool = False,
                 device: Optional[torch.device] = None):
        """feats [N, D] (float); sample_nns: N rows of neighbour indices (ragged allowed); labels [N] ints or None.
        draw_hflip: consume one np.random.randint(2) per gathered instance like the reference's HDF5 read path
        (datasets_common.py:655-672) -- the in-memory path (:651-652) draws nothing; feature_augmentation selects the
        flipped-image feature table on a 1."""
        dev = device if device is not None else feats.device
        self.device = dev
        f = feats.to(dev, torch.float32)
        fh = feats_hflip.to(dev, torch.float32) if feats_hflip is not None else None
        if normalize:
            f = f / f.norm(dim=1, keepdim=True)
            fh = fh / fh.norm(dim=1, keepdim=True) if fh is not None else None
        self.feats, self.feats_hflip = f.contiguous(), fh
        if feature_augmentation and fh is None:
            raise ValueError("feature_augmentation needs feats_hflip")
        self.feature_augmentation, self.draw_hflip = bool(feature_augmentation), bool(draw_hflip)
        self.sample_nns = [np.asarray(r) for r in sample_nns]  # host: only used to draw indices
        if len(self.sample_nns) != f.shape[0]:
            raise ValueError(f"sample_nns has {len(self.sample_nns)} rows, features {f.shape[0]}")
        # Rectangular table (the make_hdf5_nns output, [N, k]): keep it on the device and draw ALL neighbour positions of a
        # batch with one vectorised np.random.randint -- the legacy numpy stream yields the same numbers as the
        # reference's per-instance np.random.choice(row) calls (tests/test_sampler.py pins that), without a Python loop
        # of B RNG calls on the step's critical path.
        widths = {len(r) for r in self.sample_nns}
        self.nn_width = widths.pop() if len(widths) == 1 else None
        self.nns_dev = None
        if self.nn_width:
            self.nns_dev = torch.as_tensor(np.stack(self.sample_nns).astype(np.int64), device=dev)
        self.labels_host = None if labels is None else np.asarray(labels).astype(np.int64)
        self.labels = None if labels is None else torch.as_tensor(self.labels_host, device=dev)
        n = f.shape[0]
        self.possible_sampling_idxs = np.arange(n) if possible_sampling_idxs is None else np.array(possible_sampling_idxs)

    # ------------------------------------------------------------------------------------------------ gathers
    def get_instance_features(self, index) -> Tensor:
        """Rows `index` of the (optionally flip-augmented) feature table, on the device (datasets_common.py:647-679)."""
        idx = np.atleast_1d(np.asarray(index)).astype(np.int64)
        flips = None
        if self.draw_hflip:  # one draw per instance, in order, as the reference's loop does (same stream, one call)
            flips = np.random.randint(2, size=len(idx)) == 1
        ti = torch.as_tensor(idx, device=self.device)
        out = self.feats.index_select(0, ti)
        if self.feature_augmentation and flips is not None and flips.any():
            tf = torch.as_tensor(flips, device=self.device)
            out = torch.where(tf[:, None], self.feats_hflip.index_select(0, ti), out)
        return out

    # ------------------------------------------------------------------------------------------------ samplers
    def sample_conditioning_instance_balance(self, batch_size: int, weights=None) -> Tuple[Optional[Tensor], Tensor]:
        """p(h) first, then a neighbour's label: datasets_common.py:525-576."""
        if weights is None:
            sel = np.random.randint(0, len(self.possible_sampling_idxs), size=batch_size)
            sel = self.possible_sampling_idxs[sel]
        else:
            sel = np.random.choice(self.possible_sampling_idxs, batch_size, replace=True, p=weights)
        instance_gen = self.get_instance_features(sel)
        # the neighbour is drawn even without labels, as in the reference (it advances the stream)
        if self.nns_dev is not None:
            pos = np.random.randint(0, self.nn_width, size=batch_size)
            if self.labels is None:
                return None, instance_gen
            chosen = self.nns_dev[torch.as_tensor(np.asarray(sel, dtype=np.int64), device=self.device),
                                  torch.as_tensor(pos, device=self.device)]
            return self.labels.index_select(0, chosen), instance_gen
        chosen = [np.random.choice(self.sample_nns[i]) for i in sel]  # ragged rows: per-instance draws
        labels_gen = None
        if self.labels is not None:
            labels_gen = self.labels.index_select(0, torch.as_tensor(np.asarray(chosen, dtype=np.int64), device=self.device))
        return labels_gen, instance_gen

    def sample_conditioning_nnclass_balance(self, batch_size: int, weights=None, num_classes: int = 1000
                                            ) -> Tuple[Tensor, Tensor]:
        """p(y) first, then an image of that class, then an instance that has it as a neighbour: :578-622."""
        if self.labels_host is None:
            raise ValueError("nnclass_balance needs labels")
        if weights is not None:
            weights = np.array(weights) / sum(weights)
        chosen_class = np.random.choice(range(num_classes), batch_size, replace=True, p=weights)
        nn_idxs = []
        for lab in chosen_class:
            chosen_xnn = np.random.choice((self.labels_host == lab).nonzero()[0])
            nn_idxs.append(np.random.choice(self.sample_nns[chosen_xnn]))
        instance_gen = self.get_instance_features(nn_idxs)
        labels_gen = torch.as_tensor(np.asarray(chosen_class, dtype=np.int64), device=self.device)
        return labels_gen, instance_gen


def sample_conditioning_values(z_, y_, ddp=False, batch_size=1, weights_sampling=None, dataset=None,
                               constant_conditioning=False, class_cond=True, instance_cond=False,
                               nn_sampling_strategy="instance_balance"):
    """data_utils/utils.py:830-901 with `dataset` a ConditioningSampler (or the reference dataset: same method names)."""
    with torch.no_grad():
        z_.sample_()
        if not class_cond and not instance_cond:
            return z_
        if class_cond and not instance_cond:
            y_.sample_()
            if constant_conditioning:
                return z_, torch.zeros_like(y_)
            return (z_, y_) if ddp else (z_, y_.data.clone())
        if nn_sampling_strategy == "instance_balance":
            fn = dataset.sample_conditioning_instance_balance
        elif nn_sampling_strategy == "nnclass_balance":
            fn = dataset.sample_conditioning_nnclass_balance
        else:
            raise ValueError(f"unknown nn_sampling_strategy {nn_sampling_strategy!r}")
        labels_g, f_g = fn(batch_size, weights_sampling)
        if instance_cond and not class_cond:
            return z_, f_g
        return z_, labels_g, f_g
