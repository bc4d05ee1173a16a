#!/usr/bin/env python
"""Per-step times of a long back-to-back run (cfg 2): does the step slow down as the power governor settles?
Prints the forward call and the whole step for every step, plus SM clock / power samples taken in-process (NVML)."""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import triad_b200  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 80
    cfg = bench.CONFIGS["cfg2"]
    dev = torch.device("cuda", 0)
    sets = bench.make_device_inputs(cfg, cfg["B"], 1234, dev, 3)
    for s in sets:
        s[0].requires_grad_(True); s[1].requires_grad_(True)
    m = triad_b200.TriadHotPath(1.5).to(dev)
    m.triad_regularizers = False
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    rows, stop = [], [False]

    def poll():
        while not stop[0]:
            rows.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                         pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
            time.sleep(0.01)

    ev = []

    def step(q, v, mask):
        q.grad = v.grad = m.temperature.grad = None
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        clip, tok = m.compute_all_similarities_av(q, v)
        e[1].record()
        m.compute_contrastive_loss_av(clip, tok)[0].backward()
        e[2].record()
        ev.append(e)

    for i in range(3):
        step(*sets[i % 3])
    torch.cuda.synchronize()
    time.sleep(1.0)                       # idle: the governor starts from rest
    ev.clear()
    t = threading.Thread(target=poll, daemon=True)
    t.start()
    t0 = time.perf_counter()
    for i in range(n):
        step(*sets[i % 3])
    torch.cuda.synchronize()
    stop[0] = True
    t.join()
    print("step: fwd ms, rest ms, total ms")
    for i, e in enumerate(ev):
        f, r = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
        if i < 12 or i % 8 == 0 or i >= n - 3:
            print(f"{i:3d}: {f:.3f} {r:.3f} {f + r:.3f}")
    tot = ev[0][0].elapsed_time(ev[-1][2])
    print(f"mean over {n} steps: {tot / n:.3f} ms; first 10: {ev[0][0].elapsed_time(ev[9][2]) / 10:.3f}; last 20: {ev[-20][0].elapsed_time(ev[-1][2]) / 20:.3f}")
    print("clock/power samples (t ms, MHz, W):", [(round((a - t0) * 1e3), c, round(p)) for a, c, p in rows][::3])


if __name__ == "__main__":
    main()
