#!/usr/bin/env python
"""Which kernels pull the SM clock down?  Runs 0.5 s loops of (a) the forward only, (b) the backward only (dq + sort +
gather + dT), (c) the whole step, from an idle GPU, and prints per-iteration times early / late plus NVML clock and
power samples.  cfg 2 shape."""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from triad_b200 import _lib, ops  # noqa: E402


def main():
    cfg = bench.CONFIGS["cfg2"]
    dev = torch.device("cuda", 0)
    (q, v, _), = bench.make_device_inputs(cfg, cfg["B"], 1234, dev, 1)
    B, Nq = cfg["B"], cfg["Nq"]
    scale = ops.row_scale(None, B, Nq, dev)
    T = torch.tensor(1.5, device=dev)
    clip, idx = ops.maxmean_fwd(q, v, scale, T)
    g, sums, out = ops.contrastive_head(clip, T)
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)

    def fwd():
        ops.maxmean_fwd(q, v, scale, T)

    def bwd():
        ops.maxmean_bwd(q, v, idx, g, clip, scale, T, flags=_lib.BWD_UNIFORM_SCALE)

    def both():
        fwd(); bwd()

    def fwd_1cta():                          # cta_group::1: every CTA loads whole V tiles (2x the L2->SM traffic of the pair)
        ops.maxmean_fwd(q, v, scale, T, flags=_lib.FWD_FORCE_1CTA)

    def fwd_noidx():                         # forward only (no argmax pass in the epilogue): the epilogue's share of the power
        ops.maxmean_fwd(q, v, scale, T, want_idx=False)

    runs = (("fwd only", fwd, 200), ("bwd only", bwd, 280), ("fwd+bwd", both, 110))
    if len(sys.argv) > 1 and sys.argv[1] == "fwd-variants":
        runs = (("fwd 2cta", fwd, 200), ("fwd 1cta", fwd_1cta, 160), ("fwd 2cta, no argmax pass", fwd_noidx, 200))
    for name, fn, n in runs:
        torch.cuda.synchronize()
        time.sleep(2.0)
        rows, stop = [], [False]

        def poll():
            while not stop[0]:
                rows.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), round(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)))
                time.sleep(0.02)
        th = threading.Thread(target=poll, daemon=True)
        th.start()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        ev[0].record()
        for i in range(n):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        stop[0] = True
        th.join()
        t = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
        k = n // 10
        print(f"{name}: first {k}: {sum(t[:k]) / k:.3f} ms   middle: {sum(t[n // 2:n // 2 + k]) / k:.3f}   last {k}: {sum(t[-k:]) / k:.3f}   total {sum(t):.0f} ms")
        print("   clocks/power:", rows[::2])


if __name__ == "__main__":
    main()
