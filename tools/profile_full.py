#!/usr/bin/env python
"""One step of the FULL reference loss (regularisers on) for an ncu launch list: python tools/profile_full.py [cfg]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import triad_b200  # noqa: E402

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
dev = torch.device("cuda", 0)
(q, v, mask), = bench.make_device_inputs(cfg, cfg["B"], 1234, dev, 1)
q.requires_grad_(True); v.requires_grad_(True)
m = triad_b200.TriadHotPath(1.5).to(dev)
for _ in range(2):
    q.grad = v.grad = None
    if mask is None:
        clip, tok = m.compute_all_similarities_av(q, v)
        total = m.compute_contrastive_loss_av(clip, tok)[0]
    else:
        clip, tok = m.compute_all_similarities_tv(q, v, mask)
        total = m.compute_contrastive_loss_tv(clip, tok)[0]
    total.backward()
torch.cuda.synchronize()
print("loss", total.item())
