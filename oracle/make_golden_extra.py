"""Golden vectors beyond the BigGAN networks: the k-NN conditioning build (called from oracle/make_golden.py)."""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np


def _import_datasets_common(ref):
    """data_utils/datasets_common.py imports h5py unconditionally (line 33) and resnet.py imports a torchvision symbol that
    moved; both are harness-side shims, the reference files are untouched."""
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.ModuleType("h5py")
    import torchvision.models as tvm
    if not hasattr(tvm, "utils"):
        import torch.hub
        shim = types.ModuleType("torchvision.models.utils")
        shim.load_state_dict_from_url = torch.hub.load_state_dict_from_url
        sys.modules["torchvision.models.utils"] = shim
        tvm.utils = shim
    sys.path.insert(0, ref)
    import data_utils.datasets_common as dc
    return dc


def knn_golden(ref, gold):
    from oracle import knn_oracle as K
    dc = _import_datasets_common(ref)
    import torch
    n, d, k = 1500, 2048, 50
    rng = np.random.default_rng(6)
    raw = rng.standard_normal((n, d))
    raw[7] = raw[3]          # exact duplicates: distance 0 ties, resolved by the lower index
    raw[900] = raw[901] * 2  # same direction => identical after normalisation
    x32 = K.normalize_features(raw)
    nns, radii = K.obtain_nns(x32, k)
    # the reference's own in-tree path (sklearn fallback of _obtain_nns): neighbour SET per row, unordered
    ds = object.__new__(dc.ILSVRC_HDF5_feats)
    f64 = raw / np.linalg.norm(raw, axis=1, keepdims=True)
    ds.feats = torch.from_numpy(f64)
    ds.num_imgs = n
    ds._obtain_nns(k_nn=k, faiss_lib=False, gpu=False)
    mism = 0
    for i in range(n):
        ref_set = set(ds.sample_nns[i][:k]) if len(ds.sample_nns[i]) >= k else set(ds.sample_nns[i])
        mine = set(nns[i].tolist())
        if len(ds.sample_nns[i]) == k and ref_set != mine:
            mism += 1
    # rows with exact ties at the k-th boundary may legitimately differ as a set; there are none in this data besides dups
    assert mism <= 4, f"oracle neighbour sets differ from the reference sklearn path on {mism} rows"
    ref_r = np.asarray(ds.sample_nn_radius)
    np.savez_compressed(os.path.join(gold, "knn_n1500_k50.npz"), nns=nns.astype(np.int32), radii=radii,
                        ref_radius_sklearn=ref_r)
    with open(os.path.join(gold, "knn_n1500_k50.json"), "w") as f:
        json.dump({"n": n, "d": d, "k": k, "seed": 6, "rows_differing_from_reference_sklearn_set": mism,
                   "note": "features = default_rng(6).standard_normal((n,d)); row7=row3; row900=2*row901; "
                           "float64 normalise -> float32"}, f, indent=1)
    print(f"[golden] knn_n1500_k50: oracle sets == reference sklearn sets on {n - mism}/{n} rows")


def stylegan_ops_golden(ref, gold):
    """bias_act / upfirdn2d: the reference's own impl='ref' outputs and autograd gradients (1st and 2nd order)."""
    import torch
    from oracle import stylegan_ops_oracle as S
    sys.path.insert(0, os.path.join(ref, "stylegan2_ada_pytorch"))
    import warnings
    warnings.filterwarnings("ignore")
    from torch_utils.ops import bias_act as RB, upfirdn2d as RU
    out = {}
    g = torch.Generator().manual_seed(17)
    x = torch.randn(3, 6, 5, 7, generator=g) * 2
    b = torch.randn(6, generator=g)
    gy = torch.randn(3, 6, 5, 7, generator=g)
    gg = torch.randn(3, 6, 5, 7, generator=g)
    out["ba_x"], out["ba_b"], out["ba_gy"], out["ba_gg"] = x, b, gy, gg
    for act in RB.activation_funcs:
        for tag, kw in (("def", {}), ("clamp", dict(gain=1.7, clamp=1.1, alpha=0.3))):
            xr, br = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
            y = RB.bias_act(xr, br, act=act, impl="ref", **kw)
            dx, db = torch.autograd.grad(y, [xr, br], gy, create_graph=True)
            ddx = torch.autograd.grad(dx, xr, gg, allow_unused=True)[0] if dx.requires_grad else None
            xo, bo = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
            yo = S.bias_act(xo, bo, act=act, **kw)
            dxo, dbo = torch.autograd.grad(yo, [xo, bo], gy, create_graph=True)
            assert (yo - y).abs().max() <= 1e-6 and (dxo - dx).abs().max() <= 1e-5 and (dbo - db).abs().max() <= 1e-4, act
            k = f"ba_{act}_{tag}"
            out[k + "_y"], out[k + "_dx"], out[k + "_db"] = y.detach(), dx.detach(), db.detach()
            out[k + "_ddx"] = torch.zeros_like(x) if ddx is None else ddx.detach()
    f = RU.setup_filter([1, 3, 3, 1])
    assert (f - S.setup_filter([1, 3, 3, 1])).abs().max() == 0
    out["uf_f"] = f
    for site in S.UPFIRDN_SITES:
        hw = 9 if site["odd"] else 8
        xi = torch.randn(2, 5, hw, hw, generator=g)
        xr = xi.clone().requires_grad_(True)
        y = RU.upfirdn2d(xr, f, up=site["up"], down=site["down"], padding=site["padding"], gain=site["gain"], impl="ref")
        gyy = torch.randn(y.shape, generator=g)
        dx = torch.autograd.grad(y, xr, gyy)[0]
        yo = S.upfirdn2d(xi, f, up=site["up"], down=site["down"], padding=site["padding"], gain=site["gain"])
        assert yo.shape == y.shape and (yo - y).abs().max() <= 1e-6, site["name"]
        k = "uf_" + site["name"]
        out[k + "_x"], out[k + "_y"], out[k + "_gy"], out[k + "_dx"] = xi, y.detach(), gyy, dx.detach()
    # library helpers with their padding arithmetic
    xi = torch.randn(2, 3, 8, 8, generator=g)
    out["ufh_x"] = xi
    out["ufh_up"] = RU.upsample2d(xi, f, impl="ref")
    out["ufh_down"] = RU.downsample2d(xi, f, impl="ref")
    out["ufh_filt"] = RU.filter2d(xi, f, impl="ref")
    out["ufh_flip"] = RU.upfirdn2d(xi, RU.setup_filter([1, 2, 4]), up=[2, 1], down=[1, 2], padding=[1, 0, 2, 1],
                                   flip_filter=True, gain=1.5, impl="ref")
    np.savez_compressed(os.path.join(gold, "stylegan_ops.npz"), **{k: v.numpy() for k, v in out.items()})
    print(f"[golden] stylegan_ops: {len(out)} arrays; oracle == reference impl='ref'")


def stylegan_conv_golden(ref, gold):
    """conv2d_resample / modulated_conv2d of the live reference (CPU), incl. first- and second-order gradients."""
    import torch
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import stylegan_ops_oracle as S
    sys.path.insert(0, os.path.join(ref, "stylegan2_ada_pytorch"))
    from torch_utils.ops import conv2d_resample as RC, upfirdn2d as RU
    from training.networks import modulated_conv2d as Rmod
    f = RU.setup_filter([1, 3, 3, 1])
    g = torch.Generator().manual_seed(29)
    out = {"f": f}
    for name, ci, co, k, up, down, pad, hw in S.CONV_SITES:
        x = torch.randn(2, ci, hw, hw, generator=g)
        w = torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5
        xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        y = RC.conv2d_resample(xr, wr, f=f, up=up, down=down, padding=pad)
        gy = torch.randn(y.shape, generator=g)
        dx, dw = torch.autograd.grad(y, [xr, wr], gy, create_graph=True)
        v = torch.randn(dx.shape, generator=g)
        (ddw,) = torch.autograd.grad(dx, wr, v)  # second order: d<dx, v>/dw  (R1-style)
        yo = S.conv2d_resample(x, w, f, up=up, down=down, padding=pad)
        assert yo.shape == y.shape and (yo - y).abs().max() <= 2e-5, name
        k_ = "cr_" + name
        out.update({k_ + "_x": x, k_ + "_w": w, k_ + "_y": y.detach(), k_ + "_gy": gy, k_ + "_dx": dx.detach(),
                    k_ + "_dw": dw.detach(), k_ + "_v": v, k_ + "_ddw": ddw.detach()})
    # modulated conv: training (non-fused) and inference (fused) forms, with the gradients the loss terms need
    for name, up, pad, k in (("mod_plain", 1, 1, 3), ("mod_up", 2, 1, 3), ("mod_rgb", 1, 0, 1)):
        ci, co, hw = 8, 6, 8
        x = torch.randn(3, ci, hw, hw, generator=g)
        w = torch.randn(co, ci, k, k, generator=g)
        st = torch.randn(3, ci, generator=g) + 1.0
        ohw = hw * up
        noise = torch.randn(3, 1, ohw, ohw, generator=g) * 0.1
        dem = name != "mod_rgb"
        xr, wr, sr = x.clone().requires_grad_(True), w.clone().requires_grad_(True), st.clone().requires_grad_(True)
        y = Rmod(xr, wr, sr, noise=noise if dem else None, up=up, padding=pad, resample_filter=f, demodulate=dem,
                 fused_modconv=False)
        yf = Rmod(x, w, st, noise=noise if dem else None, up=up, padding=pad, resample_filter=f, demodulate=dem,
                  fused_modconv=True)
        assert (y - yf).abs().max() <= 2e-5
        gy = torch.randn(y.shape, generator=g)
        dx, dw, ds = torch.autograd.grad(y, [xr, wr, sr], gy, create_graph=True)
        (dds,) = torch.autograd.grad(dx.square().sum(), sr)  # path-length style second-order term
        yo = S.modulated_conv2d(x, w, st, noise if dem else None, up=up, padding=pad, resample_filter=f, demodulate=dem)
        assert (yo - y).abs().max() <= 5e-5, name
        out.update({name + "_x": x, name + "_w": w, name + "_s": st, name + "_noise": noise, name + "_y": y.detach(),
                    name + "_gy": gy, name + "_dx": dx.detach(), name + "_dw": dw.detach(), name + "_ds": ds.detach(),
                    name + "_dds": dds.detach()})
    np.savez_compressed(os.path.join(gold, "stylegan_conv.npz"), **{k: v.numpy() for k, v in out.items()})
    print(f"[golden] stylegan_conv: {len(out)} arrays; oracle == reference")


def main(ref, gold):
    knn_golden(ref, gold)
    stylegan_ops_golden(ref, gold)
    stylegan_conv_golden(ref, gold)
