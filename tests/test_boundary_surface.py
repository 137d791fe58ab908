"""Drop-in boundary: every public callable of the reference on the hot path has a B200 counterpart
with the same parameter names, order, kinds and literal defaults (tests/golden/reference_signatures.json, frozen from the
live reference by oracle/make_golden_signatures.py).  Extra trailing keyword parameters with defaults are allowed
(compute_dtype, grad_sync, lazy_losses): callers written for the reference never pass them."""
import importlib
import inspect
import json
import os

import pytest

from oracle.make_golden_signatures import describe, resolve
from tests.helpers import GOLD

with open(os.path.join(GOLD, "reference_signatures.json")) as _f:
    TABLE = json.load(_f)


@pytest.mark.parametrize("ref_name", sorted(TABLE))
def test_signature_matches_reference(ref_name):
    entry = TABLE[ref_name]
    mod_name, dotted = entry["b200"].split(":")
    mine = describe(resolve(importlib.import_module(mod_name), dotted))
    ref = entry["params"]
    has_varkw = any(kind == "VAR_KEYWORD" for _, kind, _ in mine)
    assert len(mine) >= len([p for p in ref if p[1] != "VAR_KEYWORD"]) or has_varkw, (ref_name, mine, ref)
    mine_by_name = {n: (i, kind, d) for i, (n, kind, d) in enumerate(mine)}
    pos = 0
    for name, kind, default in ref:
        if kind in ("VAR_KEYWORD", "VAR_POSITIONAL"):
            continue
        assert name in mine_by_name, f"{ref_name}: parameter `{name}` of the reference is missing"
        i, mkind, mdefault = mine_by_name[name]
        assert i == pos, f"{ref_name}: parameter `{name}` is at position {i}, the reference has it at {pos}"
        if default == "<required>":
            assert mdefault == "<required>", f"{ref_name}: `{name}` must stay a required argument"
        elif default != "<object>":
            assert mdefault == default, f"{ref_name}: default of `{name}` is {mdefault!r}, reference {default!r}"
        pos += 1
    for name, kind, default in mine[pos:]:  # extensions must be optional
        assert default != "<required>" or kind in ("VAR_KEYWORD", "VAR_POSITIONAL"), \
            f"{ref_name}: extra parameter `{name}` has no default"


def test_operand_staleness_is_tracked_per_parameter():
    """ops.weight_stamp: an optimiser step marks only ITS parameters stale (G's operand copies survive D's step); a call
    without an optimiser invalidates everything.  Host logic only -- no kernels involved."""
    import torch
    from ic_gan_b200 import ops
    a, b = torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(3))
    a.grad, b.grad = torch.ones(3), torch.ones(3)
    sa, sb = ops.weight_stamp(a), ops.weight_stamp(b)
    torch.optim.SGD([a], lr=0.1).step()          # the package registers a global post-step hook
    assert ops.weight_stamp(a) != sa and ops.weight_stamp(b)[1:] == sb[1:]
    sa, sb = ops.weight_stamp(a), ops.weight_stamp(b)
    ops.invalidate_operands()
    assert ops.weight_stamp(a) != sa and ops.weight_stamp(b) != sb


def test_operand_buffers_are_rewritten_in_place():
    """SNState._keep / _slot: operand copies keep their storage across rebuilds (a captured CUDA graph bakes the address)."""
    import torch
    from ic_gan_b200 import ops
    st = ops.SNState(module=None, kind="conv")
    first = st._keep("wk_fwd", torch.arange(6.0).reshape(2, 3).t())   # non-contiguous value -> stored densely
    assert first.is_contiguous() and st.wk_fwd is first
    again = st._keep("wk_fwd", torch.ones(3, 2))
    assert again is first and float(first.sum()) == 6.0
    other = st._keep("wk_fwd", torch.ones(4, 2))                        # a new shape gets a new buffer
    assert other is not first and st.wk_fwd is other
    slot = st._slot("wk_dgrad", (2, 2), torch.float32, torch.device("cpu"))
    assert st._slot("wk_dgrad", (2, 2), torch.float32, torch.device("cpu")) is slot
