"""Freeze the call signatures of the reference's public surface on the hot path (build container only; needs
/root/reference):   python oracle/make_golden_signatures.py   ->  tests/golden/reference_signatures.json
tests/test_boundary_surface.py holds the B200 drop-in modules to them (same parameter names, order and literal defaults),
which is what lets BigGAN_PyTorch/trainer.py and stylegan2_ada_pytorch/training/training_loop.py call them unchanged."""
from __future__ import annotations

import contextlib
import inspect
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ICGAN_REFERENCE", "/root/reference")

SURFACE = {  # reference module -> (B200 module, [qualified callables])
    "BigGAN_PyTorch.BigGAN": ("ic_gan_b200.biggan.model", [
        "Generator.__init__", "Generator.forward", "Discriminator.__init__", "Discriminator.forward", "G_D.__init__",
        "G_D.forward", "G_arch", "D_arch"]),
    "BigGAN_PyTorch.layers": ("ic_gan_b200.biggan.layers", [
        "SNConv2d.__init__", "SNConv2d.forward", "SNLinear.__init__", "SNLinear.forward", "SNEmbedding.__init__",
        "SNEmbedding.forward", "Attention.__init__", "Attention.forward", "ccbn.__init__", "ccbn.forward", "bn.__init__",
        "bn.forward", "GBlock.__init__", "GBlock.forward", "DBlock.__init__", "DBlock.forward"]),
    "train_fns": ("ic_gan_b200.biggan.train_fns", ["GAN_training_function"]),
    "utils": ("ic_gan_b200.biggan.train_fns", ["ema.__init__", "ema.update", "toggle_grad"]),
    "training.networks": ("ic_gan_b200.stylegan2.networks", [
        "modulated_conv2d", "FullyConnectedLayer.__init__", "Conv2dLayer.__init__", "MappingNetwork.__init__",
        "MappingNetwork.forward", "SynthesisLayer.__init__", "SynthesisLayer.forward", "ToRGBLayer.__init__",
        "SynthesisBlock.__init__", "SynthesisBlock.forward", "SynthesisNetwork.__init__", "SynthesisNetwork.forward",
        "Generator.__init__", "Generator.forward", "DiscriminatorBlock.__init__", "MinibatchStdLayer.__init__",
        "DiscriminatorEpilogue.__init__", "Discriminator.__init__", "Discriminator.forward"]),
    "training.loss": ("ic_gan_b200.stylegan2.loss", [
        "StyleGAN2Loss.__init__", "StyleGAN2Loss.run_G", "StyleGAN2Loss.run_D", "StyleGAN2Loss.accumulate_gradients"]),
    "torch_utils.ops.bias_act": ("ic_gan_b200.stylegan2.ops.bias_act", ["bias_act"]),
    "torch_utils.ops.upfirdn2d": ("ic_gan_b200.stylegan2.ops.upfirdn2d", [
        "setup_filter", "upfirdn2d", "filter2d", "upsample2d", "downsample2d"]),
    "torch_utils.ops.conv2d_resample": ("ic_gan_b200.stylegan2.ops.conv2d_resample", ["conv2d_resample"]),
    "torch_utils.ops.conv2d_gradfix": ("ic_gan_b200.stylegan2.ops.conv2d_gradfix", [
        "conv2d", "conv_transpose2d", "no_weight_gradients"]),
    "torch_utils.ops.fma": ("ic_gan_b200.stylegan2.ops.fma", ["fma"]),
}


def describe(fn):
    out = []
    for name, p in inspect.signature(fn).parameters.items():
        if name == "self":
            continue
        d = p.default
        if d is inspect.Parameter.empty:
            dv = "<required>"
        elif isinstance(d, (int, float, str, bool, type(None))):
            dv = d
        elif isinstance(d, (list, tuple)) and all(isinstance(v, (int, float, str, bool)) for v in d):
            dv = list(d)
        else:
            dv = "<object>"  # nn.ReLU(), dicts, functions ...: presence is checked, value is not
        out.append([name, str(p.kind).split(".")[-1], dv])
    return out


def resolve(mod, dotted):
    """The undecorated callable: stylegan2's @persistence.persistent_class replaces a class by a subclass whose __init__
    is (*args, **kwargs) (the real one is the base's), and @misc.profiled_function wraps functions in a closure."""
    obj = mod
    for part in dotted.split("."):
        if isinstance(obj, type) and "_orig_module_src" in vars(obj) and len(obj.__mro__) > 1:
            obj = obj.__mro__[1]
        obj = getattr(obj, part)
    if isinstance(obj, type):
        return obj
    if getattr(obj, "__closure__", None) and obj.__code__.co_name == "decorator":
        for cell in obj.__closure__:
            if inspect.isfunction(cell.cell_contents):
                return cell.cell_contents
    return inspect.unwrap(obj)


def main():
    import importlib
    sys.path[:0] = [REF, os.path.join(REF, "BigGAN_PyTorch"), os.path.join(REF, "stylegan2_ada_pytorch")]
    table = {}
    with contextlib.redirect_stdout(io.StringIO()):
        for ref_mod, (mine, names) in SURFACE.items():
            m = importlib.import_module(ref_mod)
            for n in names:
                table[f"{ref_mod}:{n}"] = {"b200": f"{mine}:{n}", "params": describe(resolve(m, n))}
    path = os.path.join(ROOT, "tests", "golden", "reference_signatures.json")
    with open(path, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print(f"wrote {len(table)} signatures to {path}")


if __name__ == "__main__":
    main()
