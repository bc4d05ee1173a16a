"""Golden-vector case table shared by ``oracle/gen_golden.py`` (which runs the REFERENCE on
these inputs, in the build container) and by the tests (which re-create the same inputs from
the seeds and compare the oracle / the CUDA path with the stored reference outputs).

Test infrastructure only.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch


@dataclass(frozen=True)
class Case:
    name: str
    kind: str            # "av" (unmasked mean) or "tv" (masked mean)
    B: int
    Nq: int
    Nv: int
    D: int
    dtype: str           # "fp32" | "bf16"
    seed: int
    T: float = 1.5
    dup_patches: bool = False    # every 3rd patch duplicates its predecessor -> exact ties
    zero_tail: int = 0           # trailing patches zeroed (patch-dropout padding, model.py:296-307)
    min_len: int = 1             # tv: shortest caption
    project: int = 0             # >0: store grads as D->project random projections (large D)
    scale: float = 1.0           # extra input gain (pushes logits around to exercise softmax)


CASES = [
    Case("av_fp32_tiny",  "av", 4, 10, 32, 64, "fp32", 11),
    Case("av_bf16_tiny",  "av", 4, 10, 32, 64, "bf16", 12),
    Case("av_fp32_cfg1",  "av", 8, 50, 256, 512, "fp32", 13, project=4),
    Case("av_bf16_cfg1",  "av", 8, 50, 256, 512, "bf16", 14, project=4),
    Case("av_bf16_ties",  "av", 5, 33, 48, 64, "bf16", 15, dup_patches=True, scale=3.0),
    Case("av_fp32_ties",  "av", 5, 33, 48, 64, "fp32", 16, dup_patches=True),
    Case("av_bf16_zpad",  "av", 6, 40, 173, 128, "bf16", 17, zero_tail=41),
    Case("av_fp32_T12",   "av", 4, 20, 64, 64, "fp32", 18, T=1.2),
    Case("av_bf16_T20",   "av", 4, 20, 64, 64, "bf16", 19, T=2.0, scale=4.0),
    Case("tv_fp32_tiny",  "tv", 6, 13, 40, 64, "fp32", 21, min_len=1),
    Case("tv_bf16_tiny",  "tv", 6, 13, 40, 64, "bf16", 22, min_len=1),
    Case("tv_bf16_77",    "tv", 8, 77, 256, 512, "bf16", 23, min_len=8, project=4),
    Case("tv_fp32_full",  "tv", 5, 16, 32, 64, "fp32", 24, min_len=16),   # all-valid masks
]

BY_NAME = {c.name: c for c in CASES}


def torch_dtype(c: Case) -> torch.dtype:
    return torch.bfloat16 if c.dtype == "bf16" else torch.float32


def build_inputs(c: Case):
    """(q, v, mask, T) for a case; deterministic in the torch CPU generator."""
    g = torch.Generator().manual_seed(c.seed)
    q = torch.randn(c.B, c.Nq, c.D, generator=g) * (c.scale / math.sqrt(c.D))
    v = torch.randn(c.B, c.Nv, c.D, generator=g) * (c.scale / math.sqrt(c.D))
    if c.dup_patches:
        v[:, 2::3] = v[:, 1::3][:, : v[:, 2::3].shape[1]]
    if c.zero_tail:
        v[:, c.Nv - c.zero_tail:] = 0
    mask: Optional[torch.Tensor] = None
    if c.kind == "tv":
        lens = torch.randint(c.min_len, c.Nq + 1, (c.B,), generator=g)
        lens[0] = c.Nq
        mask = (torch.arange(c.Nq)[None, :] < lens[:, None]).to(torch.int64)
    dt = torch_dtype(c)
    return q.to(dt), v.to(dt), mask, c.T


def projection(c: Case) -> Optional[torch.Tensor]:
    if not c.project:
        return None
    g = torch.Generator().manual_seed(c.seed + 1000)
    return torch.randn(c.D, c.project, generator=g, dtype=torch.float64)
