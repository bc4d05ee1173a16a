"""CPU tier: triad_b200.producers.patch_dropout reproduces the reference's patch dropout (src/model.py:268-308)
bit for bit for the same seed — same Bernoulli stream, kept patches first in their original order, zero rows
behind — and passes gradients to the kept patches only."""
import os
import sys
import types

import pytest
import torch

from triad_b200.producers import patch_dropout

REF_SRC = "/root/reference/src"


def _loop_version(x, drop_rate):
    """The reference's algorithm, restated (model.py:283-307)."""
    B, N, D = x.shape
    keep = torch.bernoulli(torch.ones(B, N, dtype=x.dtype) * (1 - drop_rate)).bool()
    kept = [x[i][keep[i]] for i in range(B)]
    m = max(t.size(0) for t in kept)
    return torch.stack([torch.cat([t, torch.zeros(m - t.size(0), D, dtype=x.dtype)]) for t in kept])


@pytest.mark.parametrize("rate", [0.1, 0.5, 0.9])
def test_matches_the_loop_formulation(rate):
    x = torch.randn(7, 33, 5)
    torch.manual_seed(3)
    a = patch_dropout(x, rate)
    torch.manual_seed(3)
    b = _loop_version(x, rate)
    assert torch.equal(a, b)
    assert patch_dropout(x, rate, training=False) is x and patch_dropout(x, 0) is x


def test_gradients_reach_only_kept_patches():
    x = torch.randn(4, 20, 3, requires_grad=True)
    torch.manual_seed(5)
    y = patch_dropout(x, 0.4)
    y.sum().backward()
    torch.manual_seed(5)
    keep = torch.bernoulli(torch.ones(4, 20) * 0.6).bool()
    assert torch.equal(x.grad, keep[:, :, None].expand_as(x).float())


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference tree not present (GPU box)")
def test_matches_the_reference_method():
    peft = types.ModuleType("peft")
    for n in ("LoraConfig", "get_peft_model", "TaskType"):
        setattr(peft, n, object)
    sys.modules.setdefault("peft", peft)
    sys.path.insert(0, REF_SRC)
    try:
        import model as ref_model
    finally:
        sys.path.remove(REF_SRC)

    class Stub:
        training = True
    x = torch.randn(6, 256, 16)
    torch.manual_seed(11)
    want = ref_model.ViTLoRAEmbedder.patch_dropout(Stub(), x, 0.25)
    torch.manual_seed(11)
    got = patch_dropout(x, 0.25)
    assert torch.equal(got, want)
