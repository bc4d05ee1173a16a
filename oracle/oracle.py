"""CPU restatement of TRIAD's dense max-mean similarity + symmetric InfoNCE path.

THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It is the checker the
CUDA path is compared against (``tests/``, ``__graft_entry__.smoke()``) and the
CPU baseline that ``bench.py`` times (``cpu_baseline`` / ``--impl reference``).
Nothing under ``triad_b200/`` imports it and there is no CPU fallback.

Parity pinning: the reference ships no tests and no golden vectors (SURVEY.md
§4), so this restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF:
``oracle/gen_golden.py`` imports ``/root/reference/src/{model,retrieval}.py``
in the build container, runs its unmodified functions on seeded inputs and
commits the results under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks every function below against those fixtures.

Each function cites the reference lines (relative to /root/reference) it follows.
All arithmetic is plain torch on CPU tensors (the reference *is* torch code, so
torch's CPU kernels are the reference arithmetic); float64 is used where a
"mathematically exact" companion value is useful for tolerance bookkeeping.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

__all__ = [
    "token_sims_for_query",
    "maxmean_forward",
    "infonce",
    "similarity_stats",
    "maxmean_backward",
    "contrastive_step_closed_form",
    "reference_step_autograd",
    "regularization_av",
    "regularization_tv",
    "aggregate_pair",
    "recall_at_k",
    "similarity_matrix",
    "row_scale_from_mask",
    "make_inputs",
]


# --------------------------------------------------------------------------------------
# inputs (seeded N(0,1)/sqrt(D) embeddings, right-padded masks)
# --------------------------------------------------------------------------------------
def make_inputs(B: int, Nq: int, Nv: int, D: int, dtype: torch.dtype, seed: int,
                masked: bool = False, min_len: int = 1, Bv: Optional[int] = None):
    """Seeded synthetic embeddings.  Values are generated in fp32 and cast once, so every
    consumer (oracle, CUDA path, reference) sees the same representable numbers."""
    g = torch.Generator().manual_seed(seed)
    Bv = B if Bv is None else Bv
    q = (torch.randn(B, Nq, D, generator=g) / math.sqrt(D)).to(dtype)
    v = (torch.randn(Bv, Nv, D, generator=g) / math.sqrt(D)).to(dtype)
    mask = None
    if masked:
        lens = torch.randint(min_len, Nq + 1, (B,), generator=g)
        lens[0] = Nq  # the tokenizer pads to the batch-longest caption (model.py:102-109)
        mask = (torch.arange(Nq)[None, :] < lens[:, None]).to(torch.int64)
    return q, v, mask


def row_scale_from_mask(mask: Optional[torch.Tensor], B: int, Nq: int) -> torch.Tensor:
    """Per-token weight of the (masked) mean: 1/Nq (model.py:391) or mask/clamp(sum mask,1e-7)
    (model.py:509-512)."""
    if mask is None:
        return torch.full((B, Nq), 1.0 / Nq, dtype=torch.float32)
    m = mask.to(torch.float32)
    return m / m.sum(dim=1, keepdim=True).clamp(min=1e-7)


# --------------------------------------------------------------------------------------
# forward: token similarity -> max over patches -> (masked) mean over tokens
# --------------------------------------------------------------------------------------
def token_sims_for_query(q_i: torch.Tensor, v: torch.Tensor, T) -> torch.Tensor:
    """S[j,a,p] = T * <q_i[a], v[j,p]> for one query against every image.

    model.py:384-387 / :502-505.  With bf16 inputs the reference's matmul output is bf16
    (fp32 accumulate, one rounding) and the product with the fp32 0-dim temperature is
    computed in fp32 and rounded to bf16 again; fp32 inputs stay fp32 throughout."""
    T32 = torch.as_tensor(T, dtype=torch.float32)
    acc = torch.matmul(q_i.to(torch.float32), v.to(torch.float32).transpose(1, 2))  # (Bv,Nq,Nv)
    if q_i.dtype == torch.bfloat16:
        s = acc.to(torch.bfloat16)
        return (s.to(torch.float32) * T32).to(torch.bfloat16)
    return acc * T32


def maxmean_forward(q: torch.Tensor, v: torch.Tensor, T, mask: Optional[torch.Tensor] = None
                    ) -> Dict[str, torch.Tensor]:
    """compute_all_similarities_av (model.py:370-392) / _tv (model.py:490-514) without ever
    holding more than one query's token-similarity slab.

    Returns
      clip      (Bq,Bv) fp32 : aggregated similarity before the reference's final rounding
      clip_ref  (Bq,Bv)      : the dtype the reference returns (bf16 for AV-bf16, else fp32)
      idx       (Bq,Bv,Nq) int64 : argmax patch, first index among ties (torch.max semantics)
      rowmax    (Bq,Bv,Nq) fp32  : max_p S (already rounded like the reference's token_sims)
    """
    Bq, Nq, _ = q.shape
    Bv = v.shape[0]
    scale = row_scale_from_mask(mask, Bq, Nq)
    idx = torch.empty(Bq, Bv, Nq, dtype=torch.int64)
    rowmax = torch.empty(Bq, Bv, Nq, dtype=torch.float32)
    for i in range(Bq):
        s = token_sims_for_query(q[i], v, T)
        m, ix = torch.max(s, dim=2)  # model.py:389 / :507
        idx[i] = ix
        rowmax[i] = m.to(torch.float32)
    if mask is None:
        clip = rowmax.sum(dim=2) / Nq                       # model.py:391
        clip_ref = clip.to(q.dtype)
    else:
        m32 = mask.to(torch.float32)
        clip = (rowmax * m32[:, None, :]).sum(dim=2) / m32.sum(dim=1).clamp(min=1e-7)[:, None]
        clip_ref = clip                                     # model.py:509-512 promotes to fp32
    return {"clip": clip, "clip_ref": clip_ref, "idx": idx, "rowmax": rowmax, "row_scale": scale}


# --------------------------------------------------------------------------------------
# symmetric InfoNCE (model.py:453-459, :572-578) and the statistics block (:435-450)
# --------------------------------------------------------------------------------------
def infonce(clip: torch.Tensor) -> Dict[str, torch.Tensor]:
    """loss = mean_i( -log_softmax_row(clip)[i,i] - log_softmax_col(clip)[i,i] ) / 2 and its
    closed-form gradient g = (softmax_row + softmax_col - 2 I) / (2B)  (SURVEY.md §8 a3)."""
    c = clip.to(torch.float64)
    B = c.shape[0]
    lse_r = torch.logsumexp(c, dim=1)
    lse_c = torch.logsumexp(c, dim=0)
    d = torch.diagonal(c)
    loss = ((lse_r - d).sum() + (lse_c - d).sum()) / (2 * B)
    g = (torch.exp(c - lse_r[:, None]) + torch.exp(c - lse_c[None, :])
         - 2 * torch.eye(B, dtype=torch.float64)) / (2 * B)
    return {"loss": loss, "g": g, "lse_row": lse_r, "lse_col": lse_c}


def similarity_stats(clip: torch.Tensor, prefix: str) -> Dict[str, float]:
    """The six floats of model.py:435-450,463-470 (unbiased std, hardest negative = max off-diag)."""
    c = clip.to(torch.float64)
    B = c.shape[0]
    d = torch.diagonal(c)
    off = c[~torch.eye(B, dtype=torch.bool)]
    pm, nm = d.mean().item(), off.mean().item()
    return {
        f"{prefix}_pos_sim_mean": pm,
        f"{prefix}_pos_sim_std": d.std().item() if B > 1 else float("nan"),
        f"{prefix}_neg_sim_mean": nm,
        f"{prefix}_neg_sim_std": off.std().item() if off.numel() > 1 else float("nan"),
        f"{prefix}_separation": pm - nm,
        f"{prefix}_hardest_negative": off.max().item() if off.numel() else float("nan"),
    }


# --------------------------------------------------------------------------------------
# backward through max-mean (autograd of model.py:387-391)
# --------------------------------------------------------------------------------------
def maxmean_backward(q: torch.Tensor, v: torch.Tensor, idx: torch.Tensor, g: torch.Tensor, T,
                     row_scale: torch.Tensor, clip: torch.Tensor
                     ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """dq[i,a] = T s[i,a] sum_j g[i,j] v[j,idx[i,j,a]];  dv[j,p] = sum_{i,a:idx=p} T s[i,a] g[i,j] q[i,a];
    dT = sum g*clip / T.  float64 accumulation, returned as float64."""
    Bq, Nq, D = q.shape
    Bv, Nv, _ = v.shape
    T = float(T)
    q64, v64, g64 = q.to(torch.float64), v.to(torch.float64), g.to(torch.float64)
    s64 = row_scale.to(torch.float64)
    dq = torch.zeros(Bq, Nq, D, dtype=torch.float64)
    dv = torch.zeros(Bv * Nv, D, dtype=torch.float64)
    jbase = (torch.arange(Bv) * Nv)[:, None]
    for i in range(Bq):
        flat = (jbase + idx[i]).reshape(-1)                       # (Bv*Nq,)
        w = (T * g64[i][:, None] * s64[i][None, :])               # (Bv,Nq)
        gathered = v64.reshape(Bv * Nv, D)[flat].reshape(Bv, Nq, D)
        dq[i] = (w[:, :, None] * gathered).sum(dim=0)
        contrib = (w[:, :, None] * q64[i][None, :, :]).reshape(Bv * Nq, D)
        dv.index_add_(0, flat, contrib)
    dT = (g64 * clip.to(torch.float64)).sum() / T
    return dq, dv.reshape(Bv, Nv, D), dT


def contrastive_step_closed_form(q, v, T, mask=None):
    """Forward + closed-form backward of the contrastive-only objective (no reg terms)."""
    fwd = maxmean_forward(q, v, T, mask)
    nce = infonce(fwd["clip"])
    dq, dv, dT = maxmean_backward(q, v, fwd["idx"], nce["g"], T, fwd["row_scale"], fwd["clip"])
    out = dict(fwd)
    out.update(nce)
    out.update({"dq": dq, "dv": dv, "dT": dT})
    return out


def reference_step_autograd(q: torch.Tensor, v: torch.Tensor, T: torch.Tensor,
                            mask: Optional[torch.Tensor] = None):
    """The reference's own way of doing one contrastive fwd+bwd step, restated: broadcast both
    operands to (Bq,Bv,N,D), one batched matmul that MATERIALISES the (Bq,Bv,Nq,Nv) tensor,
    max / mean, two log-softmaxes, autograd backward (model.py:384-391, :453-459).  This is
    what ``bench.py`` times as the CPU baseline ("port").  q, v, T need requires_grad."""
    Bq, Bv = q.shape[0], v.shape[0]
    lhs = q[:, None].expand(Bq, Bv, *q.shape[1:])
    rhs = v[None].expand(Bq, Bv, *v.shape[1:]).transpose(2, 3)
    tok = torch.matmul(lhs, rhs) * T
    best = tok.max(dim=3).values
    if mask is None:
        clip = best.mean(dim=2)
    else:
        m = mask[:, None, :].to(torch.float32).expand(-1, Bv, -1)
        clip = (best * m).sum(dim=2) / m.sum(dim=2).clamp(min=1e-7)
    tgt = torch.arange(Bq)
    lp_r = torch.log_softmax(clip, dim=1)
    lp_c = torch.log_softmax(clip.t(), dim=1)
    loss = (-(lp_r[tgt, tgt]) - lp_c[tgt, tgt]).mean() / 2
    loss.backward()
    return loss.detach(), clip.detach(), tok.detach()


# --------------------------------------------------------------------------------------
# regularisers that consume the dense token-similarity tensor
# --------------------------------------------------------------------------------------
def regularization_av(tok: torch.Tensor, T: torch.Tensor):
    """model.py:417-428 with :404-408: 20*relu(-log T)^2 + 0.15*mean(clamp(S,-60,0)^2)
    + 0.01*mean((S_ii[1:]-S_ii[:-1])^2).  Returns (reg, 0.01*l_smooth)."""
    B = tok.shape[0]
    l_nonneg = tok.clamp(min=-60, max=0).pow(2).mean()
    l_cal = torch.clamp(-torch.log(T), min=0) ** 2
    diag = tok[torch.arange(B), torch.arange(B)]                  # (B,Na,Nv)
    l_smooth = (diag[:, 1:] - diag[:, :-1]).pow(2).mean()
    return 20 * l_cal + 0.15 * l_nonneg + 0.01 * l_smooth, 0.01 * l_smooth


def regularization_tv(tok: torch.Tensor, thr: float, weight: float):
    """model.py:516-542: 0.15*mean(clamp(S,-20,0)^2) + weight*mean(relu(sum_t softmax_p(S_ii)/Nt - thr)^2)."""
    B = tok.shape[0]
    l_nonneg = tok.clamp(min=-20, max=0).pow(2).mean()
    diag = tok[torch.arange(B), torch.arange(B)]                  # (B,Nt,Nv)
    frac = torch.softmax(diag, dim=-1).sum(dim=1) / diag.shape[1]
    sparsity = torch.relu(frac - thr).pow(2).mean()
    return 0.15 * l_nonneg + weight * sparsity


# --------------------------------------------------------------------------------------
# retrieval use of the same math (retrieval.py:106-115, :190-198, :117-144)
# --------------------------------------------------------------------------------------
def aggregate_pair(q_feats: torch.Tensor, v_feats: torch.Tensor, temperature: float,
                   direction: str = "q2v") -> float:
    """q2v: mean_q max_p (q.v/T)  (aggregator_av_a2v / tv_t2v);  v2q: mean_p max_q (…_v2a / …_v2t).
    Note the retrieval path DIVIDES by the temperature (retrieval.py:108) where training multiplies."""
    s = torch.matmul(q_feats, v_feats.t()) / temperature
    best = s.max(dim=1).values if direction == "q2v" else s.max(dim=0).values
    return best.mean().item()


def recall_at_k(sim: np.ndarray) -> Dict[str, float]:
    """retrieval.py:117-144: rank of the diagonal under a descending argsort of each row."""
    n = sim.shape[0]
    ranks = np.empty(n, dtype=np.int64)
    for i in range(n):
        order = np.argsort(-sim[i])
        ranks[i] = int(np.nonzero(order == i)[0][0])
    return {f"r{k}": float(np.mean(ranks < k)) for k in (1, 5, 10, 20)}


def similarity_matrix(f1: torch.Tensor, f2: torch.Tensor, T) -> torch.Tensor:
    """model.py:355-368: per-pair L2-normalised bmm times temperature."""
    a = torch.nn.functional.normalize(f1, dim=-1)
    b = torch.nn.functional.normalize(f2, dim=-1)
    return torch.bmm(a, b.transpose(1, 2)) * T
