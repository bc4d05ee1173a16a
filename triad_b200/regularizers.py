"""The regularisation terms of the reference's loss: src/model.py:394-428
(audio-visual) and :516-542 (text-visual).

Two kinds of term read the token-similarity tensor the fused path never builds:

* ``l_nonneg = mean(clamp(S, lo, 0)^2)`` over ALL Bq*Bv*Nq*Nv pairs (model.py:411-412, lo=-60;
  :525-526, lo=-20).  Its gradient is dense (every negative similarity), so its backward is two
  real GEMMs.  ``DenseNonNeg`` streams the batch in image chunks: N = dL/d<q,v> for the chunk comes
  straight out of the tcgen05 forward kernel (``triad_nonneg_fused_chunk``: the epilogue writes
  coef*T*min(S,0) instead of reducing the tile, and the sums for the value and dL/dT), then two
  library GEMMs (dQ += N V, dV = N^T Q).  For fp32 inputs / shapes the tensor-core kernel does not
  take: library GEMM (raw <q,v>) -> ``triad_nonneg_chunk`` (elementwise, in place) -> the same two
  GEMMs.  Nothing of size B^2*Nq*Nv is ever resident (see DESIGN.md §4 K5 for why the two backward
  GEMMs are not fused into a flash-attention-style kernel at D = 512).
* terms on the B POSITIVE pairs only (token_sims[i,i]): temporal smoothness (model.py:394-408)
  and patch-usage sparsity (:528-541).  They touch B*Nq*Nv elements — 1/B of the tensor — and are
  written with the reference's own ATen ops on a batched GEMM of the diagonal blocks.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from ._lib import check

#: bytes of one raw-similarity chunk (rows x images-in-chunk x Nv).  2 GiB: the three GEMMs stay large and the
#: fp32 accumulation of dQ across chunks (one extra pass over dQ per chunk) is paid 4x at the B=256 shape
CHUNK_BYTES = 2 << 30


def nonneg_chunk(S: torch.Tensor, T: torch.Tensor, lo: float, coef: float, write_grad: bool,
                 sums: torch.Tensor) -> None:
    """In place on one contiguous chunk of raw dot products; accumulates into sums (fp64 [2])."""
    lib = _lib.load()
    ops._require_cuda(S, T, sums)
    if not S.is_contiguous():
        raise ValueError("nonneg_chunk needs a contiguous chunk")
    with ops._on(S):
        ws = ops._Workspace.get(lib.triad_nonneg_workspace_bytes(), S.device, "nonneg")
        check(lib.triad_nonneg_chunk(S.data_ptr(), S.numel(), ops._dtype_code(S), T.data_ptr(), float(lo), float(coef),
                                     1 if write_grad else 0, sums.data_ptr(), ws.data_ptr(), ws.numel(),
                                     ops._stream(S.device)), "triad_nonneg_chunk")


#: use the fused tcgen05 forward (triad_nonneg_fused_chunk) when the shape allows; False forces the library-GEMM +
#: elementwise path (kept for fp32 inputs, odd shapes and as a cross-check)
USE_FUSED = True


def fused_supported(q: torch.Tensor, v: torch.Tensor) -> bool:
    """Shapes the tcgen05 dense-regulariser forward takes.  Patch counts that are not a multiple of 8 (patch dropout:
    Nv ~ 190-210, model.py:296-307) are handled by nonneg_sweep, which pads the images with ZERO patches: a zero
    patch has S = 0, clamp(0, lo, 0)^2 = 0 and a zero gradient, so value and gradients are exactly those of the
    unpadded problem (the mean's denominator is passed explicitly)."""
    D, Nv = q.shape[-1], v.shape[1]
    return q.dtype == torch.bfloat16 and D % 64 == 0 and D <= 512 and (Nv + 7) // 8 * 8 <= 256


def nonneg_fused_chunk(q: torch.Tensor, vc: torch.Tensor, T: torch.Tensor, lo: float, coef: float, write_grad: bool,
                       sums: torch.Tensor):
    """N = dl_nonneg/d<q,v> for all rows of q against the images vc ([M, jc*Nv] bf16, or None when write_grad is
    False), straight from the tensor-core forward; accumulates into sums (fp64 [2])."""
    lib = _lib.load()
    Bq, Nq, D = q.shape
    jc, Nv, _ = vc.shape
    N = torch.empty(Bq * Nq, jc * Nv, dtype=torch.bfloat16, device=q.device) if write_grad else None
    with ops._on(q):
        ws = ops._Workspace.get(lib.triad_nonneg_fused_workspace_bytes(), q.device, "nonneg_fused")
        check(lib.triad_nonneg_fused_chunk(q.data_ptr(), vc.data_ptr(), T.data_ptr(), Bq, jc, Nq, Nv, D, float(lo), float(coef),
                                           None if N is None else N.data_ptr(), jc * Nv, 1 if write_grad else 0,
                                           sums.data_ptr(), ws.data_ptr(), ws.numel(), ops._stream(q.device)),
              "triad_nonneg_fused_chunk")
    return N


def nonneg_sweep(q: torch.Tensor, v: torch.Tensor, T: torch.Tensor, lo: float, numel: float, need_grads: bool,
                 chunk_bytes: int):
    """One sweep over image chunks.  Returns (sums fp64 [2] = {sum clamp^2, dl/dT for l = sum clamp^2 / numel},
    dq fp32 [Bq*Nq, D] or None, dv [Bv,Nv,D] in the input dtype or None); the gradients are those of
    sum clamp(T<q,v>, lo, 0)^2 / numel.  `numel` is the caller's normaliser (all pairs of the GLOBAL batch when the
    rows are one rank's shard)."""
    q, v = q.contiguous(), v.contiguous()
    Bq, Nq, D = q.shape
    Bv, Nv_true, _ = v.shape
    fused = USE_FUSED and fused_supported(q, v)
    if fused and Nv_true % 8:                                # zero patches up to a multiple of 8 (exact, see fused_supported)
        v = torch.nn.functional.pad(v, (0, 0, 0, 8 - Nv_true % 8))
    Nv = v.shape[1]
    M = Bq * Nq
    q2 = q.view(M, D)
    sums = torch.zeros(2, dtype=torch.float64, device=q.device)
    jc = max(1, min(Bv, int(chunk_bytes) // max(1, M * Nv * q.element_size())))
    dq32 = torch.zeros(M, D, dtype=torch.float32, device=q.device) if need_grads else None
    dv = torch.empty_like(v) if need_grads else None
    for j0 in range(0, Bv, jc):
        vc = v[j0:j0 + jc].reshape(-1, D)
        if fused:       # the tcgen05 forward writes N itself (no S chunk, no elementwise pass)
            S = nonneg_fused_chunk(q, v[j0:j0 + jc], T, lo, 2.0 / numel, need_grads, sums)
        else:
            S = torch.mm(q2, vc.t())                        # raw <q,v>, rounded to the input dtype like the reference's matmul
            nonneg_chunk(S, T, lo, 2.0 / numel, need_grads, sums)  # in place: S -> N = dl_nonneg/d<q,v>
        if need_grads:
            dq32.add_(torch.mm(S, vc))
            dv[j0:j0 + jc] = torch.mm(S.t(), q2).view(-1, Nv, D)
    if dv is not None and Nv != Nv_true:
        dv = dv[:, :Nv_true].contiguous()
    return sums, dq32, dv


class DenseNonNeg(torch.autograd.Function):
    """l_nonneg = mean(clamp(T*<q,v>, lo, 0)^2) over all token pairs, with dq, dv, dT.

    The gradients are produced during the forward sweep (one pass over the similarities serves both the value
    and the gradient), saved, and scaled by the incoming gradient in backward."""

    @staticmethod
    def forward(ctx, q, v, temperature, lo, chunk_bytes):
        ops._require_cuda(q, v)
        if q.dtype != v.dtype or q.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("DenseNonNeg supports float32 and bfloat16 embeddings of one dtype")
        Bq, Nq, D = q.shape
        Bv, Nv, _ = v.shape
        numel = float(Bq * Nq) * Bv * Nv
        T = ops.temperature_tensor(temperature, q.device)
        need = any(ctx.needs_input_grad[:3])
        sums, dq32, dv = nonneg_sweep(q, v, T, lo, numel, need, chunk_bytes)
        value = (sums[0] / numel).to(torch.float32)
        if need:
            ctx.save_for_backward(dq32.to(q.dtype).view(Bq, Nq, D), dv, sums[1].to(torch.float32))
        ctx.need = need
        ctx.t_shape = temperature.shape if isinstance(temperature, torch.Tensor) else None
        ctx.t_dtype = temperature.dtype if isinstance(temperature, torch.Tensor) else None
        return value

    @staticmethod
    def backward(ctx, gl):
        if not ctx.need:
            return None, None, None, None, None
        dq, dv, dT = ctx.saved_tensors
        gq = dq * gl.to(dq.dtype) if ctx.needs_input_grad[0] else None
        gv = dv * gl.to(dv.dtype) if ctx.needs_input_grad[1] else None
        gT = None
        if ctx.needs_input_grad[2] and ctx.t_shape is not None:
            gT = (dT * gl).reshape(ctx.t_shape).to(ctx.t_dtype)
        return gq, gv, gT, None, None


def nonneg_pressure(q, v, temperature, lo: float, chunk_bytes=None) -> torch.Tensor:
    return DenseNonNeg.apply(q, v, temperature, float(lo), int(CHUNK_BYTES if chunk_bytes is None else chunk_bytes))


def positive_pair_token_sims(q, v, temperature) -> torch.Tensor:
    """token_sims[i,i] for every i: (B,Nq,Nv) = T * q_i v_i^T — the diagonal blocks the reference stacks
    at model.py:405 / :528-534 (same rounding: matmul output in the input dtype, then * T)."""
    if q.shape[0] != v.shape[0]:
        raise ValueError("the positive-pair regularisers need as many queries as images")
    return torch.bmm(q, v.transpose(1, 2)) * temperature


def temporal_smoothness(diag: torch.Tensor) -> torch.Tensor:
    """model.py:394-408 on the diagonal blocks."""
    d = diag[:, 1:] - diag[:, :-1]
    return torch.mean(d ** 2)


def patch_sparsity(diag: torch.Tensor, threshold: float) -> torch.Tensor:
    """model.py:536-541: softmax over patches, usage fraction per patch (padded tokens included, as in the
    reference), squared excess over the threshold."""
    probs = torch.softmax(diag, dim=-1)
    frac = probs.sum(dim=1) / probs.shape[1]
    return (torch.relu(frac - threshold) ** 2).mean()
