"""The regularisation terms of the reference's loss: src/model.py:394-428
(audio-visual) and :516-542 (text-visual).

Two kinds of term read the token-similarity tensor the fused path never builds:

* ``l_nonneg = mean(clamp(S, lo, 0)^2)`` over ALL Bq*Bv*Nq*Nv pairs (model.py:411-412, lo=-60;
  :525-526, lo=-20).  Its gradient is dense (every negative similarity), so its backward is two
  real GEMMs.  ``DenseNonNeg`` streams the batch in image chunks: N = dL/d<q,v> for the chunk comes
  straight out of the tcgen05 forward kernel (``triad_nonneg_fused_chunk``: the epilogue writes
  coef*T*min(S,0) instead of reducing the tile, and the sums for the value and dL/dT) — or, in
  training, out of the SAME pass as the max-mean forward (``triad_maxmean_fwd_nonneg``) — then the
  two hand-written tcgen05 GEMMs ``triad_dense_grad_gemm`` (dQ = N V, dV = N^T Q; MN-major operands,
  no transposed copies).  For fp32 inputs / shapes the tensor-core kernels do not take: library GEMM
  (raw <q,v>) -> ``triad_nonneg_chunk`` (elementwise, in place) -> two library GEMMs.  Nothing of size B^2*Nq*Nv is ever resident (see DESIGN.md §4 K5 for why the two backward
  GEMMs are not fused into a flash-attention-style kernel at D = 512).
* terms on the B POSITIVE pairs only (token_sims[i,i]): temporal smoothness (model.py:394-408)
  and patch-usage sparsity (:528-541).  They touch B*Nq*Nv elements — 1/B of the tensor:
  ``PositivePairTerm`` = one small batched library GEMM of the diagonal blocks, ONE kernel
  (``triad_pospair_terms``: value, d value / d<q,v> and d value / dT in a single pass) and two small
  batched GEMMs in backward, instead of the ~40 ATen kernels of the reference's autograd graph.
"""
from __future__ import annotations

import time

import torch

from . import _lib, ops
from ._lib import check

#: upper bound on the bytes of one chunk of N = dL/d<q,v> (rows x images-in-chunk x Nv, bf16/fp32).  When the whole
#: batch fits (cfg 2: 8.4 GB, cfg 3: 10.3 GB of the 180 GB) there is ONE chunk: dQ = N V is a single GEMM with the
#: contraction over all images (fp32 accumulation inside the GEMM, one rounding), and nothing is accumulated across
#: chunks.  Larger batches fall back to several chunks with an fp32 dQ accumulator.  The budget is also capped at a
#: quarter of the memory that is free at the time of the call (chunk_budget).
CHUNK_BYTES = 16 << 30


#: cudaMemGetInfo is a driver query that can take milliseconds (tens of them while another client — nvidia-smi, an NVML
#: poller — talks to the driver): asked once per training step it showed up as 10-50 ms host stalls in bench.py's
#: full-loss block.  The answer is cached per device for this many seconds.
FREE_MEMORY_TTL_S = 2.0
_free_cache = {}


def _free_bytes(device) -> int:
    dev = torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    now = time.monotonic()
    hit = _free_cache.get(key)
    if hit is None or now - hit[0] > FREE_MEMORY_TTL_S:
        hit = (now, int(torch.cuda.mem_get_info(dev)[0]))
        _free_cache[key] = hit
    return hit[1]


def chunk_budget(device, chunk_bytes: int) -> int:
    if chunk_bytes < (1 << 20):
        return int(chunk_bytes)
    return max(1 << 20, min(int(chunk_bytes), _free_bytes(device) // 4))


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


#: the two backward GEMMs dQ = N V, dV = N^T Q on the hand-written tcgen05 kernel (triad_dense_grad_gemm, MN-major
#: operands); False: the library's GEMM (kept as a cross-check and for fp32)
USE_OWN_GEMM = True


def grad_gemms(N: torch.Tensor, q2: torch.Tensor, v2: torch.Tensor, need_dq: bool = True, need_dv: bool = True):
    """dQ [M,D] = N v2, dV [Kc,D] = N^T q2 for N = dL/d<q,v> [M,Kc]."""
    own = USE_OWN_GEMM and N.dtype == torch.bfloat16 and q2.dtype == torch.bfloat16 and q2.shape[1] % 8 == 0 and q2.shape[1] <= 512 \
        and N.stride(1) == 1 and N.stride(0) % 8 == 0
    if own:
        return (ops.dense_grad_gemm(N, v2, 0) if need_dq else None, ops.dense_grad_gemm(N, q2, 1) if need_dv else None)
    return (torch.mm(N, v2) if need_dq else None, torch.mm(N.t(), q2) if need_dv else None)


#: let compute_all_similarities_* produce N in the same pass as the max-mean reduction (triad_maxmean_fwd_nonneg)
MERGE_FORWARD = True


def merged_forward_ok(q: torch.Tensor, v: torch.Tensor) -> bool:
    """One pass for the max-mean forward and the dense regulariser: the shapes the tcgen05 kernel takes WITHOUT
    padding (zero patches would take part in the max over patches) and N for all images within the chunk budget."""
    if not (MERGE_FORWARD and USE_FUSED and fused_supported(q, v)) or v.shape[1] % 8:
        return False
    n_bytes = q.shape[0] * q.shape[1] * v.shape[0] * v.shape[1] * 2
    return n_bytes <= chunk_budget(q.device, CHUNK_BYTES)


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
    dq [Bq*Nq, D] (input dtype when one chunk held all images, else the fp32 accumulator) or None, dv [Bv,Nv,D] in
    the input dtype or None); the gradients are those of
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
    jc = max(1, min(Bv, chunk_budget(q.device, chunk_bytes) // max(1, M * Nv * q.element_size())))
    single = jc >= Bv                                        # one chunk: dQ = N V in one GEMM, no accumulator
    dq = torch.zeros(M, D, dtype=torch.float32, device=q.device) if (need_grads and not single) else None
    dv = torch.empty_like(v) if need_grads else None
    for j0 in range(0, Bv, jc):
        vc = v[j0:j0 + jc].reshape(-1, D)
        if fused:       # the tcgen05 forward writes N itself (no S chunk, no elementwise pass)
            S = nonneg_fused_chunk(q, v[j0:j0 + jc], T, lo, 2.0 / numel, need_grads, sums)
        else:
            S = torch.mm(q2, vc.t())                        # raw <q,v>, rounded to the input dtype like the reference's matmul
            nonneg_chunk(S, T, lo, 2.0 / numel, need_grads, sums)  # in place: S -> N = dl_nonneg/d<q,v>
        if need_grads:
            dq_c, dv_c = grad_gemms(S, q2, vc)
            if single:
                dq = dq_c
            else:
                dq.add_(dq_c)
            dv[j0:j0 + jc] = dv_c.view(-1, Nv, D)
    if dv is not None and Nv != Nv_true:
        dv = dv[:, :Nv_true].contiguous()
    return sums, dq, dv


def scale_by(x: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """x * s for a 0-dim DEVICE scalar s (the upstream gradient inside backward).  torch's broadcast of a 0-dim CUDA
    tensor takes the strided TensorIterator path (0.87 TB/s on a 65 MB gradient); this is one vectorised stream."""
    lib = _lib.load()
    x = x.contiguous()
    if x.dtype not in (torch.float32, torch.bfloat16):
        return x * s.to(x.dtype)
    s32 = s.detach().to(device=x.device, dtype=torch.float32).reshape(())
    y = torch.empty_like(x)
    with ops._on(x):
        check(lib.triad_scale(x.data_ptr(), y.data_ptr(), x.numel(), ops._dtype_code(x), s32.data_ptr(),
                              ops._stream(x.device)), "triad_scale")
    return y


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
        gq = scale_by(dq, gl) if ctx.needs_input_grad[0] else None
        gv = scale_by(dv, gl) if ctx.needs_input_grad[1] else None
        gT = None
        if ctx.needs_input_grad[2] and ctx.t_shape is not None:
            gT = (dT * gl).reshape(ctx.t_shape).to(ctx.t_dtype)
        return gq, gv, gT, None, None


class DenseNonNegFromN(torch.autograd.Function):
    """DenseNonNeg when N = dL/d<q,v> and its sums already came out of the max-mean forward
    (ops.maxmean_fwd_nonneg): only the two GEMMs dQ = N V, dV = N^T Q are left."""

    @staticmethod
    def forward(ctx, q, v, temperature, N, sums):
        Bq, Nq, D = q.shape
        Bv, Nv, _ = v.shape
        numel = float(Bq * Nq) * Bv * Nv
        need = any(ctx.needs_input_grad[:3])
        if need:
            q2, v2 = q.contiguous().view(Bq * Nq, D), v.contiguous().view(Bv * Nv, D)
            dq, dv = grad_gemms(N, q2, v2, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
            dq = dq.view(Bq, Nq, D) if dq is not None else None
            dv = dv.view(Bv, Nv, D) if dv is not None else None
            ctx.save_for_backward(dq, dv, sums[1].to(torch.float32))
        ctx.need = need
        ctx.t_shape = temperature.shape if isinstance(temperature, torch.Tensor) else None
        ctx.t_dtype = temperature.dtype if isinstance(temperature, torch.Tensor) else None
        return (sums[0] / numel).to(torch.float32)

    @staticmethod
    def backward(ctx, gl):
        if not ctx.need:
            return None, None, None, None, None
        dq, dv, dT = ctx.saved_tensors
        gq = scale_by(dq, gl) if dq is not None else None
        gv = scale_by(dv, gl) if dv is not None else None
        gT = None
        if ctx.needs_input_grad[2] and ctx.t_shape is not None:
            gT = (dT * gl).reshape(ctx.t_shape).to(ctx.t_dtype)
        return gq, gv, gT, None, None


def nonneg_pressure(q, v, temperature, lo: float, chunk_bytes=None, precomputed=None) -> torch.Tensor:
    """precomputed: (N, sums) from the merged forward (TokenSims.take_nonneg) for this very (q, v, temperature, lo)."""
    if precomputed is not None:
        return DenseNonNegFromN.apply(q, v, temperature, precomputed[0], precomputed[1])
    return DenseNonNeg.apply(q, v, temperature, float(lo), int(CHUNK_BYTES if chunk_bytes is None else chunk_bytes))


def positive_pair_token_sims(q, v, temperature) -> torch.Tensor:
    """token_sims[i,i] for every i: (B,Nq,Nv) = T * q_i v_i^T — the diagonal blocks the reference stacks
    at model.py:405 / :528-534 (same rounding: matmul output in the input dtype, then * T).  Inspection aid; the loss
    path uses PositivePairTerm."""
    if q.shape[0] != v.shape[0]:
        raise ValueError("the positive-pair regularisers need as many queries as images")
    return torch.bmm(q, v.transpose(1, 2)) * temperature


SMOOTHNESS, SPARSITY = 0, 1


class PositivePairTerm(torch.autograd.Function):
    """Temporal smoothness (mode SMOOTHNESS, model.py:394-408) or patch-usage sparsity (mode SPARSITY,
    model.py:528-541) of the positive pairs, with dq, dv, dT.  Forward: one batched library GEMM of the diagonal
    blocks (raw <q_i, v_i>, rounded to the input dtype like the reference's matmul) and triad_pospair_terms, which
    returns the value together with G = d value / d raw and d value / dT; backward: dq_i = G_i v_i, dv_i = G_i^T q_i."""

    @staticmethod
    def forward(ctx, q, v, temperature, mode, threshold):
        ops._require_cuda(q, v)
        if q.dtype != v.dtype or q.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("PositivePairTerm supports float32 and bfloat16 embeddings of one dtype")
        if q.shape[0] != v.shape[0]:
            raise ValueError("the positive-pair regularisers need as many queries as images")
        lib = _lib.load()
        q, v = q.contiguous(), v.contiguous()
        B, Nq, _ = q.shape
        Nv = v.shape[1]
        T = ops.temperature_tensor(temperature, q.device)
        raw = torch.bmm(q, v.transpose(1, 2))
        G = torch.empty_like(raw)
        sums = torch.empty(2, dtype=torch.float64, device=q.device)
        with ops._on(q):
            ws = ops._Workspace.get(lib.triad_pospair_workspace_bytes(B), q.device, "pospair")
            check(lib.triad_pospair_terms(raw.data_ptr(), ops._dtype_code(raw), T.data_ptr(), int(mode), float(threshold),
                                          B, Nq, Nv, G.data_ptr(), sums.data_ptr(), ws.data_ptr(), ws.numel(),
                                          ops._stream(q.device)), "triad_pospair_terms")
        ctx.save_for_backward(q, v, G, sums)
        ctx.t_shape = temperature.shape if isinstance(temperature, torch.Tensor) else None
        ctx.t_dtype = temperature.dtype if isinstance(temperature, torch.Tensor) else None
        return sums[0].to(torch.float32)

    @staticmethod
    def backward(ctx, gl):
        q, v, G, sums = ctx.saved_tensors
        Gs = scale_by(G, gl) if (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]) else None
        gq = torch.bmm(Gs, v) if ctx.needs_input_grad[0] else None
        gv = torch.bmm(Gs.transpose(1, 2), q) if ctx.needs_input_grad[1] else None
        gT = None
        if ctx.needs_input_grad[2] and ctx.t_shape is not None:
            gT = (sums[1].to(torch.float32) * gl).reshape(ctx.t_shape).to(ctx.t_dtype)
        return gq, gv, gT, None, None


def temporal_smoothness(q, v, temperature) -> torch.Tensor:
    """model.py:394-408 on the diagonal blocks."""
    return PositivePairTerm.apply(q, v, temperature, SMOOTHNESS, 0.0)


def patch_sparsity(q, v, temperature, threshold: float) -> torch.Tensor:
    """model.py:536-541: softmax over patches, usage fraction per patch (padded tokens included, as in the
    reference), squared excess over the threshold."""
    return PositivePairTerm.apply(q, v, temperature, SPARSITY, float(threshold))
