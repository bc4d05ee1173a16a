#!/usr/bin/env python
"""Does sampling the clocks perturb the timed region?  cfg 2 step, 60 steps per arm, alternating arms:
no sampler / nvidia-smi -lms 25 / nvidia-smi -lms 100 / in-process NVML thread (50 ms)."""
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import triad_b200  # noqa: E402


class NvmlThread:
    def __init__(self, period):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
        self.period, self.rows, self.stop = period, [], False

    def __enter__(self):
        self.t = threading.Thread(target=self.run, daemon=True)
        self.t.start()
        return self

    def run(self):
        nv = self.nv
        while not self.stop:
            self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                              nv.nvmlDeviceGetPowerUsage(self.h) / 1e3,
                              nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            time.sleep(self.period)

    def __exit__(self, *a):
        self.stop = True
        self.t.join()


class Smi:
    def __init__(self, ms):
        self.ms = ms

    def __enter__(self):
        self.p = subprocess.Popen(["nvidia-smi", "-i", "0", f"--query-gpu={bench.ClockSampler.Q}", "--format=csv,noheader,nounits",
                                   "-lms", str(self.ms)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        time.sleep(0.1)
        return self

    def __exit__(self, *a):
        self.p.terminate(); self.p.wait()


class Nothing:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass


def main():
    cfg = bench.CONFIGS["cfg2"]
    dev = torch.device("cuda", 0)
    sets = bench.make_device_inputs(cfg, cfg["B"], 1234, dev, 3)
    for s in sets:
        s[0].requires_grad_(True); s[1].requires_grad_(True)
    m = triad_b200.TriadHotPath(1.5).to(dev)
    m.triad_regularizers = False

    def step(q, v, mask):
        q.grad = v.grad = m.temperature.grad = None
        clip, tok = m.compute_all_similarities_av(q, v)
        m.compute_contrastive_loss_av(clip, tok)[0].backward()

    def timed(K):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            step(*sets[i % 3])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / K

    for i in range(10):
        step(*sets[i % 3])
    arms = [("none", Nothing), ("smi25", lambda: Smi(25)), ("smi100", lambda: Smi(100)), ("nvml50", lambda: NvmlThread(0.05))]
    res = {n: [] for n, _ in arms}
    for rep in range(4):
        for name, mk in arms:
            with mk() as ctx:
                res[name].append(timed(60))
            if name == "nvml50" and rep == 0:
                print("nvml rows", len(ctx.rows), ctx.rows[:3])
    for n, v in res.items():
        print(f"{n:8s} ms/step: " + " ".join(f"{x:.3f}" for x in v))


if __name__ == "__main__":
    main()
