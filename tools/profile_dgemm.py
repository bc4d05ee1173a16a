#!/usr/bin/env python
"""One launch of each hand-written dense-gradient GEMM at cfg 2, for ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from triad_b200 import ops  # noqa: E402

cfg = bench.CONFIGS["cfg2"]
dev = torch.device("cuda", 0)
(q, v, _), = bench.make_device_inputs(cfg, cfg["B"], 1234, dev, 1)
B, Nq, Nv, D = cfg["B"], cfg["Nq"], cfg["Nv"], cfg["D"]
N = (torch.randn(B * Nq, B * Nv, device=dev, dtype=torch.bfloat16) * 0.01)
for _ in range(2):
    a = ops.dense_grad_gemm(N, v.view(-1, D), 0)
    b = ops.dense_grad_gemm(N, q.view(-1, D), 1)
    c = torch.mm(N, v.view(-1, D))
torch.cuda.synchronize()
print(a.float().abs().mean().item(), b.float().abs().mean().item())
