#!/usr/bin/env python
"""Benchmark of the fused max-mean similarity + symmetric InfoNCE path (fwd + bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl triad|reference] [--config cfg2|cfg3|cfg4]

metric: clip-pairs/sec = B_global^2 / t(step)   (BASELINE.json).  One "step" = one forward + backward of
the hot path over one synthetic batch.  N=1 runs BASELINE cfg 2 (B=256, 250 frames x 256 patches, D=512,
bf16); N>1 (launched by torchrun, one rank per GPU) runs cfg 4 (B=8192 global, rows sharded over the ranks,
NCCL all-gather / reduce-scatter).  Rank 0 prints ONE JSON line.

`value`  : device-resident inputs, CUDA-event timing, max over ranks.
`e2e`    : the same step through the public drop-in API starting from PINNED HOST buffers (H2D copies of the
           embeddings and the D2H read of the loss inside the timed region).
`roofline`: the tcgen05 forward kernel against the measured bf16 peak (MEASURED_PEAKS.json).
`cpu_baseline`: the oracle's port of the reference's own (materialising) step on this box's host cores.
`--impl reference`: only that CPU port, as its own JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    "cfg2": dict(B=256, Nq=250, Nv=256, D=512, masked=False,
                 name="cfg2: B=256, 250 HuBERT frames x 256 DINOv2 patches, D=512, bf16 fwd+bwd"),
    "cfg3": dict(B=512, Nq=77, Nv=256, D=512, masked=True,
                 name="cfg3: B=512, 77 text tokens (ragged masks, n_i in [8,77]) x 256 patches, D=512, bf16 fwd+bwd"),
    "cfg4": dict(B=8192, Nq=250, Nv=256, D=512, masked=False,
                 name="cfg4: B=8192 global, 250 x 256, D=512, bf16 fwd+bwd, rows sharded over ranks"),
}


def algorithmic_flops(B_rows, B_cols, n_tokens_total, Nv, D):
    """2*Bcols*(sum n_i)*Nv*D forward + 4*Bcols*(sum n_i)*D sparse backward."""
    return 2.0 * B_cols * n_tokens_total * Nv * D, 4.0 * B_cols * n_tokens_total * D


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["bf16_tflops"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
    return 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, val in zip(names, r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        hi = sorted(sm)[len(sm) // 2:]            # the samples under load are the upper half when the region is short
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "sm_mhz_upper_half_median": statistics.median(hi)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's port of the reference step (materialises token_sims, autograd)
# ------------------------------------------------------------------------------------------------
def cpu_reference_step_factory(cfg, B_sample):
    import torch
    from oracle import oracle as O
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    q, v, mask = O.make_inputs(B_sample, cfg["Nq"], cfg["Nv"], cfg["D"], torch.float32, seed=1234,
                               masked=cfg["masked"], min_len=8)
    T = torch.tensor(1.5, requires_grad=True)
    q.requires_grad_(True)
    v.requires_grad_(True)

    def step():
        q.grad = v.grad = T.grad = None
        loss, _, _ = O.reference_step_autograd(q, v, T, mask)
        return float(loss)
    return step


def cpu_baseline(cfg, budget_s=12.0, B_sample=32):
    step = cpu_reference_step_factory(cfg, B_sample)
    step()                                   # warm-up
    t0, n = time.perf_counter(), 0
    while True:
        step(); n += 1
        el = time.perf_counter() - t0
        if el >= budget_s or n >= 50:
            break
    import torch
    return {"value": B_sample * B_sample * n / el, "unit": "clip-pairs/s", "cores": torch.get_num_threads(),
            "kind": "port",
            "sample": f"oracle.reference_step_autograd (the reference's materialising fwd+bwd, fp32) on a "
                      f"B={B_sample} sub-batch of the same shape (Nq={cfg['Nq']}, Nv={cfg['Nv']}, D={cfg['D']}), "
                      f"{n} steps in {el:.1f} s; pairs/s is per-pair work, so it transfers to the full batch"}


def run_reference(args, cfg_key):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference is
    Python and does not travel to the GPU box), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cfg = CONFIGS[cfg_key]
    Bs = 32
    step = cpu_reference_step_factory(cfg, Bs)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    val = Bs * Bs * args.steps / el
    cores = torch.get_num_threads()
    sample = (f"each step = fwd+bwd of the reference's materialising path (oracle port, fp32) on a B={Bs} sub-batch "
              f"of {cfg['name']}")
    print(json.dumps({
        "impl": "reference", "metric": "clip-pairs/sec", "value": val, "unit": "clip-pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "sample_batch": Bs},
        "cpu_baseline": {"value": val, "unit": "clip-pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "clip-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def make_device_inputs(cfg, B_local, seed, device, n_sets):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    sets = []
    for _ in range(n_sets):
        q = (torch.randn(B_local, cfg["Nq"], cfg["D"], generator=g, device=device) / cfg["D"] ** 0.5).to(torch.bfloat16)
        v = (torch.randn(B_local, cfg["Nv"], cfg["D"], generator=g, device=device) / cfg["D"] ** 0.5).to(torch.bfloat16)
        mask = None
        if cfg["masked"]:
            lens = torch.randint(8, cfg["Nq"] + 1, (B_local,), generator=g, device=device)
            lens[0] = cfg["Nq"]
            mask = (torch.arange(cfg["Nq"], device=device)[None, :] < lens[:, None]).to(torch.int64)
        sets.append((q, v, mask))
    return sets


def run_triad(args, cfg_key):
    import torch
    import torch.distributed as dist
    import triad_b200
    from triad_b200 import _lib
    from triad_b200.dist import CudaKernels, sharded_contrastive_step

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    _lib.check(lib.triad_device_check(local), "triad_device_check")

    cfg = CONFIGS[cfg_key]
    B = cfg["B"]
    assert B % world == 0
    Bl = B // world
    n_sets = 3 if world == 1 else 2
    sets = make_device_inputs(cfg, Bl, 1234 + rank, dev, n_sets)
    in_bytes = sum(t.numel() * t.element_size() for t in sets[0][:2])
    model = triad_b200.TriadHotPath(temperature=1.5).to(dev)
    model.triad_regularizers = False       # BASELINE.json's metric: max-mean similarity + InfoNCE, fwd + bwd
    T = model.temperature
    tokens_local = int(sets[0][2].sum().item()) if cfg["masked"] else Bl * cfg["Nq"]
    fwd_ev = []

    def step_single(q, v, mask, record=False):
        q.grad = v.grad = T.grad = None
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if mask is None:
            clip, tok = model.compute_all_similarities_av(q, v)
        else:
            clip, tok = model.compute_all_similarities_tv(q, v, mask)
        if record:
            e1.record(); fwd_ev.append((e0, e1))
        if mask is None:
            total, con, reg, smooth, stats = model.compute_contrastive_loss_av(clip, tok)
        else:
            total, stats = model.compute_contrastive_loss_tv(clip, tok)
        total.backward()
        return total

    class TimedKernels(CudaKernels):
        """The product kernels, with CUDA events around the forward call (for the roofline entry)."""
        record = False

        def maxmean_fwd(self, q, v, scale, T_):
            if not self.record:
                return super().maxmean_fwd(q, v, scale, T_)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = super().maxmean_fwd(q, v, scale, T_)
            e1.record(); fwd_ev.append((e0, e1))
            return out

    timed_kernels = TimedKernels()

    def step_sharded(q, v, mask, record=False):
        timed_kernels.record = record
        out = sharded_contrastive_step(q, v, T, mask, kernels=timed_kernels)
        return out["loss"]

    step = step_single if world == 1 else step_sharded
    for s in sets:
        s[0].requires_grad_(world == 1); s[1].requires_grad_(world == 1)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, K):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            fn(i)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1) / K], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- device-resident arm -------------------------------------------------------------------
    for i in range(args.warmup):
        step(*sets[i % n_sets])
    launches0 = lib.triad_launch_count()
    with ClockSampler(local) as clocks:
        ms = timed(lambda i: step(*sets[i % n_sets], record=True), args.steps)
    gpu_launches = int(lib.triad_launch_count() - launches0)     # libtriad_b200.so kernels inside the timed region (this rank)
    value = float(B) * B / (ms * 1e-3)

    # ---- end-to-end arm: pinned host inputs -> H2D -> step -> loss D2H --------------------------
    # Every step's embeddings start in PINNED HOST memory and its loss is read back on the host.  As in any
    # input pipeline, the H2D copy of step i+1 is issued on a copy stream while step i computes (double
    # buffered device staging); each step still pays for its own copy and its own D2H read inside the timed
    # region, but copy and compute overlap instead of serialising.
    host = [(q.detach().cpu().pin_memory(), v.detach().cpu().pin_memory(),
             None if m is None else m.cpu().pin_memory()) for (q, v, m) in sets]
    d2h = {"bytes": 0}
    copy_stream = torch.cuda.Stream(device=dev)
    staged = {}

    def issue_copy(i):
        hq, hv, hm = host[i % n_sets]
        with torch.cuda.stream(copy_stream):
            q = hq.to(dev, non_blocking=True)
            v = hv.to(dev, non_blocking=True)
            m = None if hm is None else hm.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged[i] = (q, v, m, ev)

    # The loss of every step is copied to PINNED host memory (D2H, 4 bytes) on the compute stream and read on the
    # host one step later — after the next step has been queued — so the device never idles waiting for Python,
    # exactly like a training loop that logs the loss with a one-step lag.  Every step's value is read inside the
    # timed region (the last one before the closing synchronise).
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
    seen = {"n": 0, "last": None}

    def e2e_step(i):
        if i not in staged:
            issue_copy(i)
        q, v, m, ev = staged.pop(i)
        torch.cuda.current_stream().wait_event(ev)
        for t in (q, v, m):
            if t is not None:
                t.record_stream(torch.cuda.current_stream())
        issue_copy(i + 1)                     # next step's inputs travel while this step computes
        q.requires_grad_(world == 1); v.requires_grad_(world == 1)
        loss = step(q, v, m)
        loss_host[i % 2].copy_(loss.detach().float(), non_blocking=True)      # D2H read of the step's result
        loss_ev[i % 2].record()
        d2h["bytes"] = 4
        if i > 0:                             # consume the previous step's loss on the host
            loss_ev[(i - 1) % 2].synchronize()
            seen["last"] = float(loss_host[(i - 1) % 2]); seen["n"] += 1

    def e2e_run(K):
        staged.clear()
        seen["n"] = 0
        for i in range(K):
            e2e_step(i)
        loss_ev[(K - 1) % 2].synchronize()
        seen["last"] = float(loss_host[(K - 1) % 2]); seen["n"] += 1
        assert seen["n"] == K and seen["last"] == seen["last"]      # K losses read on the host, finite
        staged.clear()                        # the look-ahead copy issued by the last step is not used

    e2e_run(min(args.warmup, 3))
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    e1.record()
    sync_all()
    e2e_t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms = e2e_t.item()
    h2d = in_bytes + (host[0][2].numel() * 8 if host[0][2] is not None else 0)

    # ---- the reference's FULL loss (contrastive + regularisers), reported beside the metric ----
    full = None
    if world == 1:
        model.triad_regularizers = True
        for i in range(2):
            step(*sets[i % n_sets])
        k_full = max(3, min(args.steps, 10))
        full_ms = timed(lambda i: step(*sets[i % n_sets]), k_full)
        model.triad_regularizers = False
        full = {"value": float(B) * B / (full_ms * 1e-3), "unit": "clip-pairs/s", "ms_per_step": full_ms, "steps": k_full,
                "what": "same step with the reference's regularisers on (model.py:394-428 / :516-542): dense non-negative "
                        "pressure (tcgen05 forward in dense-regulariser mode + 2 library GEMMs per image chunk), smoothness / "
                        "sparsity on the positive pairs; not part of BASELINE.json's metric"}

    # ---- single-GPU rate at the multi-GPU workload's per-rank shape (the strong-scaling denominator) ----------
    # N > 1 runs cfg 4 (B = 8192); one GPU's share of that at 8 ranks is 1024 queries x 8192 images.  Timing that
    # shape here gives the per-GPU rate the N-GPU numbers should be compared with (cfg 2's B = 256 step is short
    # enough to run at boost clocks; 0.8 s of continuous tensor work runs under the power cap).
    rank_shape = None
    if world == 1 and cfg_key == "cfg2" and not args.no_rank_shape:
        c4 = CONFIGS["cfg4"]
        Bq4, Bv4 = c4["B"] // 8, c4["B"]
        (q4, _, _), = make_device_inputs(c4, Bq4, 99, dev, 1)
        (_, v4, _), = make_device_inputs(c4, Bv4, 98, dev, 1)
        T4 = torch.tensor(1.5, device=dev)
        ck = CudaKernels()

        def rank_step():
            sc = ck.row_scale(None, Bq4, c4["Nq"], dev)
            clip4, idx4 = ck.maxmean_fwd(q4, v4, sc, T4)
            lse4, cp4 = ck.infonce_partial(clip4, Bv4, 0)
            g4, _ = ck.infonce_finish(clip4, Bv4, 0, lse4, cp4.reshape(1, 2, Bv4))
            return ck.maxmean_bwd(q4, v4, idx4, g4, clip4, sc, T4)

        rank_step()
        rs_ms = timed(lambda i: rank_step(), 2)
        rank_shape = {"workload": f"one rank's share of cfg4 at 8 GPUs: {Bq4} queries x {Bv4} images, fwd + InfoNCE block + bwd, "
                                  "no collectives", "ms_per_step": rs_ms, "value": float(Bq4) * Bv4 / (rs_ms * 1e-3),
                      "unit": "clip-pairs/s per GPU"}
        del q4, v4
        torch.cuda.empty_cache()

    # ---- roofline of the dominant kernel (tcgen05 forward) ---------------------------------------
    peak, peak_sustained, peak_src = measured_peaks()
    roof = None
    if fwd_ev:
        fwd_ms = statistics.mean(a.elapsed_time(b) for a, b in fwd_ev)
        f_fwd, f_bwd = algorithmic_flops(Bl, B, tokens_local, cfg["Nv"], cfg["D"])
        achieved = f_fwd / (fwd_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(cfg_key)
        roof = {"bound": "tensor", "kernel": "maxmean_tc_kernel (fwd call: memset + kernel + finalize_clip)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": f"{peak_src} (burst); sustained {peak_sustained}",
                "frac_of_sustained": achieved / peak_sustained, "ms": fwd_ms, "traffic": traffic,
                "step_flops": f_fwd + f_bwd,
                "step_frac_of_peak": (f_fwd + f_bwd) * world / (ms * 1e-3) / 1e12 / (peak * world)}

    out = None
    if rank == 0:
        out = {
            "metric": "clip-pairs/sec", "value": value, "unit": "clip-pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": cfg["name"], "global_batch": B, "rows_per_rank": Bl,
                       "parallelism": f"row-sharded x{world}" if world > 1 else "single GPU",
                       "l2": f"{n_sets} rotating input sets ({n_sets * in_bytes / 2**20:.0f} MiB) > 126 MB L2, "
                             "so every step reads its embeddings from HBM"},
            "clocks": clocks.summary(),
            "e2e": {"value": float(B) * B / (e2e_ms * 1e-3), "unit": "clip-pairs/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h["bytes"],
                    "pipeline": "pinned-host inputs of step i+1 are copied (copy stream) while step i computes; each "
                                "step's loss is copied to pinned host memory and read on the host one step later"},
            "gpu_launches": gpu_launches,
            "roofline": roof,
        }
        if full is not None:
            out["full_loss"] = full
        if rank_shape is not None:
            out["cfg4_rank_shape_1gpu"] = rank_shape
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(cfg, args.cpu_seconds)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="triad", choices=["triad", "reference"])
    ap.add_argument("--config", default="auto", choices=["auto"] + list(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rank-shape", action="store_true", help="skip the cfg4 per-rank-shape timing at N=1")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world > 1:
        args.gpus = world
    cfg_key = args.config if args.config != "auto" else ("cfg2" if max(args.gpus, world) == 1 else "cfg4")
    if args.impl == "reference":
        run_reference(args, cfg_key)
    else:
        if args.gpus > 1 and world == 1:
            raise SystemExit("for --gpus N > 1 launch with: python -m torch.distributed.run --nnodes=1 "
                             "--nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...")
        run_triad(args, cfg_key)


if __name__ == "__main__":
    main()
