"""Producers of the hot path's inputs: the projection head shared by the reference's three
embedders and its patch dropout, each as one kernel.

``ProjectionHead``  src/model.py:32-34,68 (audio), :81-83,116 (text), :253-255,326 (visual):
    ``projection2(layer_norm(projection1(x)))``.  The module keeps the reference's attribute names and parameter
    shapes (``projection1`` / ``layer_norm`` / ``projection2``, so a checkpoint's ``state_dict`` loads unchanged);
    the forward is ONE launch of ``triad_project_tokens`` (tcgen05 GEMM -> LayerNorm epilogue -> tcgen05 GEMM ->
    bias / optional L2 normalisation), emitting the bf16 ``[B, N, D]`` K-major tensor the similarity kernel reads.
    The arithmetic follows the reference under autocast (bf16 GEMMs with fp32 accumulation, LayerNorm in fp32).
    Backward (training): closed-form through the three layers with library GEMMs (plain matmuls; the hidden
    activations are recomputed, not stored).

``patch_dropout``  src/model.py:268-308 without its per-image Python loop: the same ``torch.bernoulli`` call (same
    random stream, so the same output for the same seed), then ``triad_patch_compact`` moves the kept patches to
    the front of every image and zero-fills the rest.  The one host synchronisation left is the one that sizes the
    output (the batch's longest kept list), which the reference's ``max(...)`` needs too.

CUDA only, like the rest of the package: CPU tensors raise.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from ._lib import check

HIDDEN = 512          # the reference hard-codes the head's hidden width (model.py:32, :81, :253)


def project_tokens(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, ln_g: torch.Tensor, ln_b: torch.Tensor, ln_eps: float,
                   w2: torch.Tensor, b2: torch.Tensor, normalize: bool = False) -> torch.Tensor:
    """Raw call: x (..., Din) bf16 -> (..., Dout) bf16."""
    lib = _lib.load()
    ops._require_cuda(x, w1, b1, ln_g, ln_b, w2, b2)
    if x.dtype != torch.bfloat16:
        raise TypeError("project_tokens computes in bf16 (the reference's autocast dtype); cast the encoder output first")
    Din = x.shape[-1]
    Dout = w2.shape[0]
    if w1.shape != (HIDDEN, Din) or w2.shape[1] != HIDDEN:
        raise ValueError(f"expected projection1 ({HIDDEN},{Din}) and projection2 (Dout,{HIDDEN}), got {tuple(w1.shape)} / {tuple(w2.shape)}")
    x2 = x.contiguous().view(-1, Din)
    M = x2.shape[0]
    w1b, w2b = w1.detach().to(torch.bfloat16).contiguous(), w2.detach().to(torch.bfloat16).contiguous()
    f32 = lambda t: t.detach().to(torch.float32).contiguous()                                         # noqa: E731
    b1f, gf, bf, b2f = f32(b1), f32(ln_g), f32(ln_b), f32(b2)
    out = torch.empty(M, Dout, dtype=torch.bfloat16, device=x.device)
    with ops._on(x):
        ws = ops._Workspace.get(lib.triad_project_workspace_bytes(), x.device, "project")
        check(lib.triad_project_tokens(x2.data_ptr(), w1b.data_ptr(), b1f.data_ptr(), gf.data_ptr(), bf.data_ptr(), float(ln_eps),
                                       w2b.data_ptr(), b2f.data_ptr(), M, Din, Dout, 1 if normalize else 0,
                                       out.data_ptr(), ws.data_ptr(), ws.numel(), ops._stream(x.device)), "triad_project_tokens")
    return out.view(*x.shape[:-1], Dout)


class _ProjectFn(torch.autograd.Function):
    """Forward: the fused kernel.  Backward: the chain rule through Linear -> LayerNorm -> Linear with library GEMMs
    (the bf16 hidden activations are recomputed from x)."""

    @staticmethod
    def forward(ctx, x, w1, b1, g, b, w2, b2, eps):
        ctx.save_for_backward(x, w1, b1, g, b, w2)
        ctx.eps = eps
        ctx.b2_dtype = b2.dtype
        return project_tokens(x, w1, b1, g, b, eps, w2, b2, normalize=False)

    @staticmethod
    def backward(ctx, gy):
        x, w1, b1, g, b, w2 = ctx.saved_tensors
        Din = x.shape[-1]
        x2 = x.reshape(-1, Din)
        gy2 = gy.reshape(-1, gy.shape[-1]).to(torch.bfloat16)
        w1b, w2b = w1.to(torch.bfloat16), w2.to(torch.bfloat16)
        h = (x2 @ w1b.t() + b1.to(torch.bfloat16)).float()                    # Linear 1 output (bf16-rounded), as fp32
        mu = h.mean(dim=1, keepdim=True)
        rstd = torch.rsqrt(h.var(dim=1, unbiased=False, keepdim=True) + ctx.eps)
        xh = (h - mu) * rstd
        ln = (xh * g.float() + b.float()).to(torch.bfloat16)
        gw2 = gy2.t() @ ln                                                    # (Dout, 512)
        gb2 = gy2.float().sum(dim=0)
        gln = (gy2 @ w2b).float()                                             # (M, 512)
        gg = (gln * xh).sum(dim=0)
        gb = gln.sum(dim=0)
        gxh = gln * g.float()
        gh = (gxh - gxh.mean(dim=1, keepdim=True) - xh * (gxh * xh).mean(dim=1, keepdim=True)) * rstd
        ghb = gh.to(torch.bfloat16)
        gw1 = ghb.t() @ x2
        gb1 = gh.sum(dim=0)
        gx = (ghb @ w1b).view_as(x)
        return (gx, gw1.to(w1.dtype), gb1.to(b1.dtype), gg.to(g.dtype), gb.to(b.dtype), gw2.to(w2.dtype),
                gb2.to(ctx.b2_dtype), None)


class ProjectionHead(torch.nn.Module):
    """projection1 -> layer_norm -> projection2 of the reference's embedders, as one kernel.

    ``head(x)`` for x (B, N, Din) bf16 returns (B, N, embedding_dim) bf16; ``head.embed(x)`` additionally L2-normalises
    every token (retrieval.py:93-94)."""

    def __init__(self, in_features: int, embedding_dim: int = 512):
        super().__init__()
        self.projection1 = torch.nn.Linear(in_features, HIDDEN)
        self.layer_norm = torch.nn.LayerNorm(HIDDEN)
        self.projection2 = torch.nn.Linear(HIDDEN, embedding_dim)

    def _args(self):
        return (self.projection1.weight, self.projection1.bias, self.layer_norm.weight, self.layer_norm.bias,
                self.projection2.weight, self.projection2.bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        w1, b1, g, b, w2, b2 = self._args()
        if torch.is_grad_enabled() and (x.requires_grad or any(t.requires_grad for t in self._args())):
            return _ProjectFn.apply(x, w1, b1, g, b, w2, b2, self.layer_norm.eps)
        return project_tokens(x, w1, b1, g, b, self.layer_norm.eps, w2, b2, normalize=False)

    @torch.no_grad()
    def embed(self, x: torch.Tensor) -> torch.Tensor:
        w1, b1, g, b, w2, b2 = self._args()
        return project_tokens(x, w1, b1, g, b, self.layer_norm.eps, w2, b2, normalize=True)


class _PatchCompact(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, keep, max_len):
        lib = _lib.load()
        B, N, D = x.shape
        x = x.contiguous()
        keep8 = keep.to(torch.uint8).contiguous()
        out = torch.empty(B, max_len, D, dtype=x.dtype, device=x.device)
        with ops._on(x):
            check(lib.triad_patch_compact(x.data_ptr(), keep8.data_ptr(), B, N, D, x.element_size(), max_len, out.data_ptr(),
                                          ops._stream(x.device)), "triad_patch_compact")
        ctx.save_for_backward(keep)
        ctx.n = N
        return out

    @staticmethod
    def backward(ctx, gy):
        (keep,) = ctx.saved_tensors
        B, L, D = gy.shape
        pos = keep.cumsum(dim=1) - 1
        gx = gy.new_zeros(B, ctx.n, D)
        bi, ni = keep.nonzero(as_tuple=True)
        gx[bi, ni] = gy[bi, pos[bi, ni]]
        return gx, None, None


def patch_dropout(x: torch.Tensor, drop_rate: float, training: bool = True) -> torch.Tensor:
    """(B, N, D) -> (B, max_kept, D): kept patches compacted to the front of each image, zero rows behind.
    Differentiable w.r.t. ``x`` (gradients reach the kept patches only), like the reference's indexing."""
    if not training or drop_rate == 0:
        return x
    ops._require_cuda(x)
    B, N, D = x.shape
    keep = torch.bernoulli(torch.ones(B, N, device=x.device, dtype=x.dtype) * (1 - drop_rate)).bool()
    max_len = int(keep.sum(dim=1).max().item())                   # the one host synchronisation (output shape)
    return _PatchCompact.apply(x, keep, max_len)
