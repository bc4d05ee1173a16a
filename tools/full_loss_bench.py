#!/usr/bin/env python
"""Times one training step of the FULL reference loss (contrastive + regularisers)
next to the contrastive-only step at a BASELINE shape.

    python tools/full_loss_bench.py [cfg2|cfg3] [iters]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import triad_b200  # noqa: E402


def main():
    cfg = dict(bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"])
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dev = torch.device("cuda", 0)
    (q, v, mask), = bench.make_device_inputs(cfg, cfg["B"], 1234, dev, 1)
    q.requires_grad_(True); v.requires_grad_(True)
    m = triad_b200.TriadHotPath(1.5).to(dev)

    def step():
        q.grad = v.grad = m.temperature.grad = None
        if mask is None:
            clip, tok = m.compute_all_similarities_av(q, v)
            total = m.compute_contrastive_loss_av(clip, tok)[0]
        else:
            clip, tok = m.compute_all_similarities_tv(q, v, mask)
            total = m.compute_contrastive_loss_tv(clip, tok)[0]
        total.backward()
        return total

    from triad_b200 import regularizers as R
    for reg, own in ((False, True), (True, True), (True, False)):
        m.triad_regularizers = reg
        R.USE_OWN_GEMM = own
        loss = step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        B = cfg["B"]
        print(f"{cfg['name']}\n  regularisers={reg} own_gemm={own}: loss {loss.item():.6f}  {ms:.3f} ms/step  {B * B / ms / 1e3:.2f} M pairs/s  "
              f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")


if __name__ == "__main__":
    main()
