#!/usr/bin/env python
"""One launch of the merged forward (max-mean + dense regulariser's N) at a BASELINE shape, for ncu:
   python tools/profile_merged.py [cfg2]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from triad_b200 import ops  # noqa: E402


def main():
    cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
    dev = torch.device("cuda", 0)
    (q, v, _), = bench.make_device_inputs(cfg, cfg["B"], 1234, dev, 1)
    B, Nq, Nv = cfg["B"], cfg["Nq"], cfg["Nv"]
    T = torch.tensor(1.5, device=dev)
    scale = ops.row_scale(None, B, Nq, dev)
    coef = 2.0 / (float(B * Nq) * B * Nv)
    for _ in range(2):
        out = ops.maxmean_fwd_nonneg(q, v, scale, T, -60.0, coef)
    torch.cuda.synchronize()
    print("sums", out[3].tolist())


if __name__ == "__main__":
    main()
