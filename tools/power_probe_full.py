#!/usr/bin/env python
"""Which kernel of the FULL loss makes the board throttle?  0.6 s loops, each from an idle GPU, of the merged forward,
the two hand-written dense-gradient GEMMs, the same GEMM through cuBLAS, and the whole full-loss step; per-iteration
times (early / middle / late / worst) plus NVML SM clock, board power and the clock-event reasons.  cfg 2 shape."""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import triad_b200  # noqa: E402
from triad_b200 import ops  # noqa: E402


def main():
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    cfg = bench.CONFIGS["cfg2"]
    dev = torch.device("cuda", 0)
    (q, v, _), = bench.make_device_inputs(cfg, cfg["B"], 1234, dev, 1)
    B, Nq, Nv, D = cfg["B"], cfg["Nq"], cfg["Nv"], cfg["D"]
    N = torch.randn(B * Nq, B * Nv, device=dev, dtype=torch.bfloat16) * 0.01
    q2, v2 = q.view(-1, D), v.view(-1, D)
    m = triad_b200.TriadHotPath(1.5).to(dev)
    qg, vg = q.clone().requires_grad_(True), v.clone().requires_grad_(True)

    def merged_fwd():
        m.compute_all_similarities_av(qg, vg)

    def full():
        qg.grad = vg.grad = m.temperature.grad = None
        clip, tok = m.compute_all_similarities_av(qg, vg)
        m.compute_contrastive_loss_av(clip, tok)[0].backward()

    runs = (("merged forward", merged_fwd, 150), ("dgemm dQ = N V", lambda: ops.dense_grad_gemm(N, v2, 0), 200),
            ("dgemm dV = N^T Q", lambda: ops.dense_grad_gemm(N, q2, 1), 200), ("cuBLAS N V", lambda: torch.mm(N, v2), 200),
            ("full loss step", full, 60))
    which = sys.argv[1:] or None
    for name, fn, n in runs:
        if which and not any(w in name for w in which):
            continue
        fn(); fn()
        torch.cuda.synchronize()
        time.sleep(2.0)
        rows, stop = [], [False]

        def poll():
            while not stop[0]:
                rows.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), round(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3),
                             hex(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))))
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
        k = max(1, n // 10)
        print(f"{name}: first {k}: {sum(t[:k]) / k:.3f} ms   middle: {sum(t[n // 2:n // 2 + k]) / k:.3f}   last {k}: {sum(t[-k:]) / k:.3f}"
              f"   worst {max(t):.2f}   total {sum(t):.0f} ms")
        print("   per-iteration:", [round(x, 2) for x in t[::max(1, n // 40)]])
        print("   clock MHz / W / reasons:", rows[::2], flush=True)


if __name__ == "__main__":
    main()
