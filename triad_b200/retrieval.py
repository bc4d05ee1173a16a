"""Retrieval use of the max-mean kernel: drop-ins for src/retrieval.py's aggregators
(:106-115, :190-198), recall@k (:117-144) and the two metric drivers (:146-188, :250-292).

The reference fills each N x N similarity matrix with a Python double loop — 10^6 iterations,
each an H2D copy, a matmul and an ``.item()``.  Here a whole matrix is a handful of launches of
the same forward kernel used for training (both directions are max-mean with the operands
swapped), and the rank of the diagonal is computed on the device.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Sequence

import numpy as np
import torch

from . import _lib, ops
from ._lib import check


def _scores(q_feats: torch.Tensor, v_feats: torch.Tensor, temperature, direction: int) -> torch.Tensor:
    """scores[n] for one query (Nq,D) against a gallery (n,Nv,D); retrieval divides by T."""
    lib = _lib.load()
    q = q_feats.contiguous()
    gal = v_feats.contiguous() if v_feats.dim() == 3 else v_feats.contiguous().unsqueeze(0)
    if q.dtype != gal.dtype:
        gal = gal.to(q.dtype)
    n_img, Nv, D = gal.shape
    Nq = q.shape[0]
    dt = ops._dtype_code(q)
    T = ops.temperature_tensor(temperature, q.device)
    out = torch.empty(n_img, dtype=torch.float32, device=q.device)
    nws = lib.triad_retrieve_workspace_bytes(Nq, n_img, Nv, D, dt)
    with ops._on(q):
        ws = ops._Workspace.get(nws, q.device, "retrieve")
        check(lib.triad_retrieve_scores(q.data_ptr(), Nq, gal.data_ptr(), n_img, Nv, D, dt, T.data_ptr(), 1, direction,
                                        out.data_ptr(), ws.data_ptr(), ws.numel(), ops._stream(q.device)),
              "triad_retrieve_scores")
    return out


def aggregator_av_a2v(a_feats, v_feats, temperature) -> float:
    """mean over audio frames of max over patches of (a.v/T) — retrieval.py:106-110."""
    return _scores(a_feats, v_feats, temperature, 0).item()


def aggregator_av_v2a(a_feats, v_feats, temperature) -> float:
    """mean over patches of max over audio frames — retrieval.py:112-115."""
    return _scores(a_feats, v_feats, temperature, 1).item()


def aggregator_tv_t2v(t_feats, v_feats, temperature) -> float:
    """retrieval.py:190-193."""
    return _scores(t_feats, v_feats, temperature, 0).item()


def aggregator_tv_v2t(t_feats, v_feats, temperature) -> float:
    """retrieval.py:195-198."""
    return _scores(t_feats, v_feats, temperature, 1).item()


def retrieve_topk(q_feats: torch.Tensor, gallery: torch.Tensor, temperature, k: int, direction: int = 0):
    """Top-k gallery items for one query (BASELINE cfg 5): (scores fp32 [k], ids int32 [k])."""
    lib = _lib.load()
    s = _scores(q_feats, gallery, temperature, direction)
    n = s.numel()
    out_s = torch.empty(k, dtype=torch.float32, device=s.device)
    out_i = torch.empty(k, dtype=torch.int32, device=s.device)
    nws = lib.triad_topk_workspace_bytes(n, k)
    with ops._on(s):
        ws = ops._Workspace.get(nws, s.device, "topk")
        check(lib.triad_topk(s.data_ptr(), n, k, out_s.data_ptr(), out_i.data_ptr(), ws.data_ptr(), ws.numel(),
                             ops._stream(s.device)), "triad_topk")
    return out_s, out_i


def pairwise_similarity(q_list: Sequence[torch.Tensor], v_list: Sequence[torch.Tensor], temperature,
                        direction: int = 0, device="cuda") -> torch.Tensor:
    """sim[i,j] = aggregator(q_list[i], v_list[j]) for every pair, on the device.

    direction 0: mean_q max_v (a2v / t2v); direction 1: mean_v max_q (v2a / v2t).  Items may have
    different lengths (text is truncated to its valid tokens, retrieval.py:243-244): they are
    grouped by length and each (query-length, gallery-length) group is one forward launch."""
    T = ops.temperature_tensor(temperature, torch.device(device))
    rows, cols = (q_list, v_list) if direction == 0 else (v_list, q_list)
    n_r, n_c = len(rows), len(cols)
    sim = torch.empty(n_r, n_c, dtype=torch.float32, device=device)
    by_len_r: Dict[int, List[int]] = defaultdict(list)
    by_len_c: Dict[int, List[int]] = defaultdict(list)
    for i, t in enumerate(rows):
        by_len_r[t.shape[0]].append(i)
    for j, t in enumerate(cols):
        by_len_c[t.shape[0]].append(j)
    for nr, ri in by_len_r.items():
        R = torch.stack([rows[i] for i in ri]).to(device)
        scale = ops.row_scale(None, R.shape[0], nr, R.device)
        for nc, ci in by_len_c.items():
            Cm = torch.stack([cols[j] for j in ci]).to(device=device, dtype=R.dtype)
            clip, _ = ops.maxmean_fwd(R, Cm, scale, T, want_idx=False, flags=_lib.FWD_DIVIDE_BY_T)
            sim[torch.tensor(ri, device=device)[:, None], torch.tensor(ci, device=device)[None, :]] = clip
    return sim if direction == 0 else sim.t().contiguous()


def compute_recall_at_k(sim_matrix) -> Dict[str, float]:
    """R@1/5/10/20 with the diagonal as ground truth — retrieval.py:117-144.  Accepts the numpy
    array the reference passes or a CUDA tensor; ranks are computed on the device."""
    lib = _lib.load()
    if isinstance(sim_matrix, np.ndarray):
        sim = torch.from_numpy(np.ascontiguousarray(sim_matrix, dtype=np.float32)).cuda()
    else:
        sim = sim_matrix.detach().float().contiguous()
    N = sim.shape[0]
    ranks = torch.empty(N, dtype=torch.int32, device=sim.device)
    with ops._on(sim):
        check(lib.triad_diag_ranks(sim.data_ptr(), N, ranks.data_ptr(), ops._stream(sim.device)), "triad_diag_ranks")
    r = ranks.cpu().numpy()
    return {"r1": float(np.mean(r < 1)), "r5": float(np.mean(r < 5)),
            "r10": float(np.mean(r < 10)), "r20": float(np.mean(r < 20))}


def _metrics(q_list, v_list, temperature, device, tag_q: str, tag_v: str) -> Dict[str, float]:
    fwd = compute_recall_at_k(pairwise_similarity(q_list, v_list, temperature, 0, device))
    # the reference's V->Q matrix is indexed [image i, query j] (retrieval.py:170-175)
    rev = compute_recall_at_k(pairwise_similarity(q_list, v_list, temperature, 1, device).t().contiguous())
    out = {}
    for k in (1, 5, 10, 20):
        out[f"{tag_q}->{tag_v}_r{k}"] = fwd[f"r{k}"]
    for k in (1, 5, 10, 20):
        out[f"{tag_v}->{tag_q}_r{k}"] = rev[f"r{k}"]
    return out


def metrics_from_features(q_list, v_list, temperature, device="cuda", kind: str = "av") -> Dict[str, float]:
    """The eight recalls of retrieval.py:177-187 / :281-291 from already-embedded items."""
    return _metrics(q_list, v_list, temperature, device, "A" if kind == "av" else "T", "V")


def compute_av_retrieval_metrics(model, dataset, subset_file, device="cuda"):
    """retrieval.py:146-188 with the 2 x 10^6-iteration aggregator loops replaced.  Subset selection
    and embedding are the caller-side steps of the reference (retrieval.py:9-30, :32-104), kept as is:
    pass the reference module's helpers through ``model``/``dataset`` unchanged."""
    import importlib
    ref = importlib.import_module("retrieval")      # the reference's own module, for its embed helpers
    indices = ref.select_subset_indices(dataset, subset_file, subset_size=1000)
    audio_feats, video_feats, _ = ref.embed_av_subset(model, dataset, indices, device=device, batch_size=8)
    return metrics_from_features(audio_feats, video_feats, model.temperature.item(), device, "av")


def compute_tv_retrieval_metrics(model, dataset, subset_file, device="cuda"):
    """retrieval.py:250-292, see compute_av_retrieval_metrics."""
    import importlib
    ref = importlib.import_module("retrieval")
    indices = ref.select_subset_indices(dataset, subset_file, subset_size=1000)
    text_feats, image_feats = ref.embed_tv_subset(model, dataset, indices, device=device, batch_size=8)
    return metrics_from_features(text_feats, image_feats, model.temperature.item(), device, "tv")
