"""Producers of the hot path's inputs: the fused projection head against torch.nn (fp32 and the
autocast recipe) and the device patch dropout against the reference's loop (src/model.py:268-308), bit for bit for the
same seed.  GPU tier, except the check that CPU tensors are refused (no fallback)."""
import pytest
import torch

from triad_b200.producers import ProjectionHead, patch_dropout
from tests.helpers import rel_err


def test_cpu_tensors_are_refused():
    with pytest.raises(RuntimeError):
        patch_dropout(torch.randn(2, 8, 16), 0.5)
    head = ProjectionHead(64, 32)
    with pytest.raises(RuntimeError):
        head(torch.randn(2, 3, 64).bfloat16())


def _loop_version(x, drop_rate):
    """The reference's algorithm, restated (model.py:283-307)."""
    B, N, D = x.shape
    keep = torch.bernoulli(torch.ones(B, N, device=x.device, dtype=x.dtype) * (1 - drop_rate)).bool()
    kept = [x[i][keep[i]] for i in range(B)]
    m = max(t.size(0) for t in kept)
    return torch.stack([torch.cat([t, torch.zeros(m - t.size(0), D, dtype=x.dtype, device=x.device)]) for t in kept])


@pytest.mark.gpu
@pytest.mark.parametrize("rate", [0.1, 0.5, 0.9])
@pytest.mark.parametrize("dtype,shape", [(torch.bfloat16, (7, 256, 512)), (torch.float32, (5, 33, 12)), (torch.bfloat16, (3, 1030, 64))])
def test_patch_dropout_matches_the_loop_formulation(rate, dtype, shape):
    x = torch.randn(*shape, device="cuda").to(dtype)
    torch.manual_seed(3)
    a = patch_dropout(x, rate)
    torch.manual_seed(3)
    b = _loop_version(x, rate)
    assert torch.equal(a, b)
    assert patch_dropout(x, rate, training=False) is x and patch_dropout(x, 0) is x


@pytest.mark.gpu
def test_patch_dropout_matches_the_staged_reference_method():
    from oracle import ref_loader
    ref = ref_loader.load()
    if ref is None:
        pytest.skip("oracle/_ref not staged")

    class Stub:
        training = True
    x = torch.randn(6, 256, 64, device="cuda").bfloat16()
    torch.manual_seed(11)
    want = ref[0].ViTLoRAEmbedder.patch_dropout(Stub(), x, 0.25)
    torch.manual_seed(11)
    got = patch_dropout(x, 0.25)
    assert torch.equal(got, want)


@pytest.mark.gpu
def test_patch_dropout_gradients_reach_only_kept_patches():
    x = torch.randn(4, 20, 8, device="cuda", requires_grad=True)
    torch.manual_seed(5)
    y = patch_dropout(x, 0.4)
    y.sum().backward()
    torch.manual_seed(5)
    keep = torch.bernoulli(torch.ones(4, 20, device="cuda") * 0.6).bool()
    assert torch.equal(x.grad, keep[:, :, None].expand_as(x).float())


def _torch_head(head, x, autocast):
    """The reference's three modules (model.py:68): under autocast (bf16 Linear, fp32 LayerNorm) or in fp32."""
    if autocast:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return head.projection2(head.layer_norm(head.projection1(x)))
    return head.projection2(head.layer_norm(head.projection1(x.float())))


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,Din,Dout", [(4, 250, 768, 512), (3, 77, 768, 512), (2, 256, 384, 512), (1, 5, 64, 64),
                                          (2, 130, 1024, 256), (1, 1, 128, 16), (2, 300, 768, 320)])
def test_projection_head_matches_torch_nn(B, N, Din, Dout):
    torch.manual_seed(Din + N)
    head = ProjectionHead(Din, Dout).cuda()
    with torch.no_grad():                                   # non-trivial affine parameters
        head.layer_norm.weight.uniform_(0.5, 1.5)
        head.layer_norm.bias.uniform_(-0.3, 0.3)
    x = torch.randn(B, N, Din, device="cuda").bfloat16()
    with torch.no_grad():
        got = head(x)
        ac = _torch_head(head, x, True)
        f32 = _torch_head(head, x, False)
    assert got.shape == (B, N, Dout) and got.dtype == torch.bfloat16 and got.is_contiguous()
    assert rel_err(got, ac) < 4e-3                          # same recipe, different accumulation order: bf16 roundings
    assert rel_err(got, f32) < 1e-2                         # the north star's bf16 bar against fp32 torch.nn
    # + F.normalize (retrieval.py:93-94)
    emb = head.embed(x)
    want = torch.nn.functional.normalize(f32, dim=2)
    assert rel_err(emb, want) < 1e-2
    assert (emb.float().norm(dim=2) - 1).abs().max().item() < 1e-2


@pytest.mark.gpu
def test_projection_head_backward_matches_autograd():
    torch.manual_seed(0)
    head = ProjectionHead(256, 128).cuda()
    ref = ProjectionHead(256, 128).cuda()
    ref.load_state_dict(head.state_dict())
    x = torch.randn(3, 40, 256, device="cuda").bfloat16().requires_grad_()
    x2 = x.detach().clone().requires_grad_()
    w = torch.randn(3, 40, 128, device="cuda")
    (head(x).float() * w).sum().backward()
    (_torch_head(ref, x2, False) * w).sum().backward()
    assert rel_err(x.grad, x2.grad) < 2e-2
    for (n, a), (_, b) in zip(head.named_parameters(), ref.named_parameters()):
        assert rel_err(a.grad, b.grad) < 2e-2, n
