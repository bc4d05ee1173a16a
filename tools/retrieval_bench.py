#!/usr/bin/env python
"""BASELINE cfg 5: one query (audio Nq=250 or text Nq=77) against a gallery of n_img images x Nv patches,
forward-only max-mean scores + top-k.  Reports ms/query, TFLOP/s and gallery GB/s (the HBM floor is
gallery_bytes / 6.5 TB/s).

    python tools/retrieval_bench.py [n_img=100000] [Nv=1024] [Nq=77,250] [iters=3]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from triad_b200 import retrieval as R  # noqa: E402


def main():
    n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    Nv = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    nqs = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "77,250").split(",")]
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    D, k = 512, 10
    dev = torch.device("cuda", 0)
    t0 = time.time()
    gal = torch.empty(n_img, Nv, D, dtype=torch.bfloat16, device=dev)
    g = torch.Generator(device=dev).manual_seed(7)
    step = max(1, (1 << 30) // (Nv * D * 2))                       # fill 1 GiB at a time
    for i in range(0, n_img, step):
        blk = torch.randn(min(step, n_img - i), Nv, D, generator=g, device=dev, dtype=torch.float32)
        gal[i:i + blk.shape[0]] = torch.nn.functional.normalize(blk, dim=2).to(torch.bfloat16)
    torch.cuda.synchronize()
    gbytes = gal.numel() * 2
    print(f"gallery {n_img} x {Nv} x {D} bf16 = {gbytes / 1e9:.1f} GB (filled in {time.time() - t0:.1f} s)")
    for Nq in nqs:
        q = torch.nn.functional.normalize(torch.randn(Nq, D, generator=g, device=dev), dim=1).to(torch.bfloat16)
        s, ids = R.retrieve_topk(q, gal, 1.5, k)                    # warm-up
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            s, ids = R.retrieve_topk(q, gal, 1.5, k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        flops = 2.0 * Nq * n_img * Nv * D
        # spot check of the winner against a plain fp32 evaluation
        j = int(ids[0])
        ref = ((q.float() @ gal[j].float().t()) / 1.5).max(dim=1).values.mean().item()
        print(f"Nq={Nq:4d}: {ms:9.3f} ms/query  {flops / ms / 1e9:8.1f} TFLOP/s  {gbytes / ms / 1e6:8.1f} GB/s of gallery  "
              f"top1 id {j} score {s[0].item():.5f} (fp32 check {ref:.5f})")


if __name__ == "__main__":
    main()
