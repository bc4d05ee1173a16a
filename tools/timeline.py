#!/usr/bin/env python
"""Kernel timeline of one steady-state step (torch.profiler / CUPTI activity records, no replay): start, duration
and the idle gap before every kernel — where the step's time goes BETWEEN kernels.

    python tools/timeline.py [cfg2] [steps=6] [full]       # full: with the reference's regularisers on
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
import triad_b200  # noqa: E402


def main():
    cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    dev = torch.device("cuda", 0)
    sets = bench.make_device_inputs(cfg, cfg["B"], 1234, dev, 3)
    for s in sets:
        s[0].requires_grad_(True); s[1].requires_grad_(True)
    m = triad_b200.TriadHotPath(1.5).to(dev)
    m.triad_regularizers = len(sys.argv) > 3 and sys.argv[3] == "full"

    def step(q, v, mask):
        q.grad = v.grad = m.temperature.grad = None
        if mask is None:
            clip, tok = m.compute_all_similarities_av(q, v)
            total = m.compute_contrastive_loss_av(clip, tok)[0]
        else:
            clip, tok = m.compute_all_similarities_tv(q, v, mask)
            total = m.compute_contrastive_loss_tv(clip, tok)[0]
        total.backward()

    for i in range(5):
        step(*sets[i % 3])
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(steps):
            step(*sets[i % 3])
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # the last full step: from the last-but-one forward kernel to the last one
    # (the max-mean forward, kMode 0 or 2; a regulariser-only pass is the same kernel with kMode 1)
    fw = [k for k, e in enumerate(evs) if re.search(r"maxmean_tc_kernel<\d, (false|true), [02]>", e.name)]
    a, b = fw[-2], fw[-1]
    t0 = evs[a].time_range.start
    prev_end = None
    busy = 0.0
    print(f"{'start us':>10} {'gap us':>8} {'dur us':>9}  kernel")
    for e in evs[a:b]:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        gap = (e.time_range.start - prev_end) if prev_end is not None else 0.0
        print(f"{s:10.1f} {gap:8.1f} {d:9.1f}  {e.name[:90]}")
        prev_end = max(prev_end or 0, e.time_range.end)
        busy += d
    span = evs[b].time_range.start - t0
    print(f"step span {span:.1f} us, kernel time {busy:.1f} us, idle {span - busy:.1f} us")


if __name__ == "__main__":
    main()
