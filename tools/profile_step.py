#!/usr/bin/env python
"""A short, fixed run of the hot path for ncu: `python tools/profile_step.py [cfg] [steps]`
(2 steps of BASELINE cfg 2 by default; no CPU baseline, no clocks sampling)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import triad_b200  # noqa: E402


def main():
    cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dev = torch.device("cuda", 0)
    (q, v, mask), = bench.make_device_inputs(cfg, cfg["B"], 1234, dev, 1)
    q.requires_grad_(True); v.requires_grad_(True)
    m = triad_b200.TriadHotPath(1.5).to(dev)
    m.triad_regularizers = False
    for _ in range(steps):
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


if __name__ == "__main__":
    main()
