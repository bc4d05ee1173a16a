"""Producers of the hot path's inputs (first piece): the reference's patch dropout
(src/model.py:268-308) without its per-sample Python loop.

The reference draws a Bernoulli keep-mask, then for each of the B images boolean-indexes the kept patch
embeddings (one device synchronisation per image), pads every image to the batch's longest kept list with zero
rows and stacks.  Those zero rows DO take part in the max over patches downstream (model.py:296-307), so the
layout matters: kept patches first, in their original order, zeros after.  Here the same mask (same
``torch.bernoulli`` call, so the same random stream and therefore the same output for the same seed) is
compacted with one cumulative sum and one scatter; the only synchronisation left is the one that sizes the
output (the batch's longest kept list), which the reference's ``max(...)`` also needs.
"""
from __future__ import annotations

import torch


def patch_dropout(x: torch.Tensor, drop_rate: float, training: bool = True) -> torch.Tensor:
    """(B, N, D) -> (B, max_kept, D): kept patches compacted to the front of each image, zero rows behind.
    Differentiable w.r.t. ``x`` (gradients reach the kept patches only), like the reference's indexing."""
    if not training or drop_rate == 0:
        return x
    B, N, D = x.shape
    keep = torch.bernoulli(torch.ones(B, N, device=x.device, dtype=x.dtype) * (1 - drop_rate)).bool()
    pos = keep.cumsum(dim=1) - 1                                   # slot of every kept patch inside its image
    max_len = int(keep.sum(dim=1).max().item())                   # the one host synchronisation (output shape)
    out = x.new_zeros(B, max_len, D)
    bi, ni = keep.nonzero(as_tuple=True)
    out = out.index_put((bi, pos[bi, ni]), x[bi, ni])
    return out
