#!/usr/bin/env python
"""Bring-up diagnostics for the GPU box: runs each stage in its own subprocess (with a timeout)
so a faulting kernel variant cannot take the later stages down with it.

    python tools/diag.py            # all stages, log to gpurun_out/diag.log
    python tools/diag.py simt       # one stage in-process
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ["simt", "nce_bwd", "tc1_small", "tc2_small", "tc1_mid", "tc2_mid", "tc_ragged", "time_cfg2"]


def _fwd(q, v, T, mask=None, flags=0, idx=True):
    import torch
    from triad_b200 import ops
    scale = ops.row_scale(mask, q.shape[0], q.shape[1], q.device)
    Tt = torch.tensor(float(T), device=q.device)
    clip, ix = ops.maxmean_fwd(q, v, scale, Tt, want_idx=idx, flags=flags, check_watchdog=True)
    return clip, ix, scale, Tt


def _cmp_fwd(tag, q, v, T, mask, flags, ref_flags=1):
    """compare a kernel variant against the SIMT kernel on the same device data"""
    import torch
    clip, ix, _, _ = _fwd(q, v, T, mask, flags)
    clip_r, ix_r, _, _ = _fwd(q, v, T, mask, ref_flags)
    torch.cuda.synchronize()
    bad = (ix != ix_r).sum().item()
    err = (clip - clip_r).abs().max().item()
    print(f"[{tag}] idx mismatches {bad}/{ix.numel()}  max|clip diff| {err:.3e}  clip[0,:4]={clip[0,:4].tolist()} ref={clip_r[0,:4].tolist()}", flush=True)
    if bad:
        w = (ix != ix_r).nonzero()[:8]
        print("   first mismatches (j, row):", w.tolist(), ix[ix != ix_r][:8].tolist(), ix_r[ix != ix_r][:8].tolist())
    return bad, err


def stage_simt():
    import torch
    from oracle import oracle as O
    from oracle.cases import CASES, build_inputs
    from tests.helpers import load_golden
    import numpy as np
    for c in CASES:
        q, v, mask, T = build_inputs(c)
        gold = load_golden(c.name)
        qd, vd = q.cuda(), v.cuda()
        md = mask.cuda() if mask is not None else None
        clip, ix, _, _ = _fwd(qd, vd, T, md, flags=1)
        B, Nq = c.B, c.Nq
        from triad_b200 import ops
        idx = ops.idx_to_reference_layout(ix, B, Nq).cpu()
        bad = (idx != torch.from_numpy(gold["idx"].astype(np.int64))).sum().item()
        ref = O.maxmean_forward(q, v, T, mask)
        err = (clip.cpu() - ref["clip"]).abs().max().item()
        print(f"[simt {c.name}] idx mismatches vs reference golden {bad}/{idx.numel()}  max|clip-oracle| {err:.3e}", flush=True)


def stage_nce_bwd():
    import torch
    import triad_b200
    from oracle import oracle as O
    for dt in (torch.float32, torch.bfloat16):
        for masked in (False, True):
            q, v, mask = O.make_inputs(7, 19, 45, 64, dt, seed=3, masked=masked)
            ref = O.contrastive_step_closed_form(q, v, 1.5, mask)
            m = triad_b200.TriadHotPath(1.5).cuda()
            m.triad_fwd_flags = 1
            qd, vd = q.cuda().requires_grad_(), v.cuda().requires_grad_()
            if masked:
                total, stats = m.forward_features_tv(qd, vd, mask.cuda())
                con = total
            else:
                total, con, reg, sm, stats = m.forward_features_av(qd, vd)
            con.backward()
            rel = lambda a, b: ((a.detach().double().cpu() - b.double()).norm() / b.double().norm()).item()
            print(f"[nce_bwd {dt} masked={masked}] loss {con.item():.6f} vs {ref['loss'].item():.6f}  dq {rel(qd.grad, ref['dq']):.2e} "
                  f"dv {rel(vd.grad, ref['dv']):.2e}  dT {m.temperature.grad.item():.6e} vs {ref['dT'].item():.6e}", flush=True)
            st = O.similarity_stats(ref["clip"], "tv" if masked else "av")
            print("   stats", {k: f"{stats[k]:.5f}/{st[k]:.5f}" for k in stats})


def _tc_cases(small):
    if small:
        return [(2, 64, 256, 64), (1, 128, 256, 64), (2, 64, 256, 512), (3, 100, 256, 128), (4, 250, 256, 512)]
    return [(16, 250, 256, 512), (40, 77, 256, 512), (64, 250, 256, 512)]


def stage_tc(cta1, small):
    import torch
    from oracle import oracle as O
    flags = 2 if cta1 else 0
    for (B, Nq, Nv, D) in _tc_cases(small):
        q, v, _ = O.make_inputs(B, Nq, Nv, D, torch.bfloat16, seed=B + Nq)
        _cmp_fwd(f"tc{'1' if cta1 else '2'} B={B} Nq={Nq} Nv={Nv} D={D}", q.cuda(), v.cuda(), 1.5, None, flags)


def stage_tc_ragged():
    import torch
    from oracle import oracle as O
    for flags in (2, 0):
        for (B, Nq, Nv, D) in [(5, 33, 173, 128), (6, 77, 200, 512), (3, 50, 16, 64), (9, 7, 255, 512), (4, 300, 96, 256)]:
            q, v, mask = O.make_inputs(B, Nq, Nv, D, torch.bfloat16, seed=B * 7 + Nv, masked=True)
            v[:, Nv - Nv // 5:] = 0        # zero-padded trailing patches
            v[:, 2::3] = v[:, 1::3][:, : v[:, 2::3].shape[1]]   # exact duplicates -> ties
            _cmp_fwd(f"ragged flags={flags} B={B} Nq={Nq} Nv={Nv} D={D}", q.cuda(), v.cuda(), 1.5, mask.cuda(), flags)


def stage_time_cfg2():
    import torch
    import triad_b200
    from oracle import oracle as O
    B, Nq, Nv, D = 256, 250, 256, 512
    q, v, _ = O.make_inputs(B, Nq, Nv, D, torch.bfloat16, seed=1234)
    qd, vd = q.cuda(), v.cuda()
    for flags, name in ((2, "tc 1cta"), (0, "tc 2cta")):
        try:
            _fwd(qd, vd, 1.5, None, flags)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            for _ in range(10):
                clip, ix, scale, Tt = _fwd(qd, vd, 1.5, None, flags)
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / 10
            fl = 2.0 * B * B * Nq * Nv * D
            print(f"[time fwd {name}] {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
        except Exception as e:  # noqa
            print(f"[time fwd {name}] FAILED {e}", flush=True)
    # spot-check the big shape against SIMT on a slice of queries
    bad, err = _cmp_fwd("cfg2 slice 2cta", qd[:24].contiguous(), vd, 1.5, None, 0)
    m = triad_b200.TriadHotPath(1.5).cuda()
    qg, vg = qd.clone().requires_grad_(), vd.clone().requires_grad_()
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        clip, tok = m.compute_all_similarities_av(qg, vg)
        e[1].record()
        total, con, reg, sm, stats = m.compute_contrastive_loss_av(clip, tok)
        e[2].record()
        con.backward()
        e[3].record()
        torch.cuda.synchronize()
        print(f"[time step {it}] fwd {e[0].elapsed_time(e[1]):.3f} loss {e[1].elapsed_time(e[2]):.3f} bwd {e[2].elapsed_time(e[3]):.3f} ms "
              f"wall {1e3 * (time.time() - t0):.2f} ms  loss={con.item():.5f}", flush=True)


def run_stage(name):
    if name == "simt":
        stage_simt()
    elif name == "nce_bwd":
        stage_nce_bwd()
    elif name == "tc1_small":
        stage_tc(True, True)
    elif name == "tc2_small":
        stage_tc(False, True)
    elif name == "tc1_mid":
        stage_tc(True, False)
    elif name == "tc2_mid":
        stage_tc(False, False)
    elif name == "tc_ragged":
        stage_tc_ragged()
    elif name == "time_cfg2":
        stage_time_cfg2()
    else:
        raise SystemExit(f"unknown stage {name}")


def main():
    if len(sys.argv) > 1:
        run_stage(sys.argv[1])
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "diag.log"), "w")
    for s in STAGES:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), s], cwd=ROOT, timeout=150,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            out, rc = r.stdout, r.returncode
        except subprocess.TimeoutExpired as e:
            out, rc = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), "TIMEOUT"
        msg = f"===== stage {s}: rc={rc} ({time.time() - t0:.1f}s) =====\n{out}\n"
        log.write(msg)
        log.flush()
        print(msg, flush=True)


if __name__ == "__main__":
    main()
