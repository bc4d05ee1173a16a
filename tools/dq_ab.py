#!/usr/bin/env python
"""A/B of the dq kernels: the software-pipelined gather (bwd_dq_pipe.cu, variants via TRIAD_DQ_VARIANT)
against the round-1 staged kernel and the generic gather — bit-equality on a set of shapes, then timing at cfg 2.

    python tools/dq_ab.py            # runs itself once per variant
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    from triad_b200 import _lib, ops
    dev = torch.device("cuda", 0)
    var = os.environ.get("TRIAD_DQ_VARIANT", "0")

    def run(Bq, Bv, Nq, Nv, D=512, iters=0):
        g0 = torch.Generator(device=dev).manual_seed(Bq * 7 + Bv * 3 + Nq + Nv)
        q = (torch.randn(Bq, Nq, D, generator=g0, device=dev) / D ** 0.5).to(torch.bfloat16)
        v = (torch.randn(Bv, Nv, D, generator=g0, device=dev) / D ** 0.5).to(torch.bfloat16)
        scale = ops.row_scale(None, Bq, Nq, dev)
        T = torch.tensor(1.5, device=dev)
        clip, idx = ops.maxmean_fwd(q, v, scale, T)
        g = torch.randn(Bq, Bv, generator=g0, device=dev) / (Bq * Bv) ** 0.5
        outs = {}
        for name, fl in (("pipe", 0), ("staged", _lib.BWD_DQ_STAGED), ("generic", _lib.BWD_GENERIC_DQ)):
            dq, _, _ = ops.maxmean_bwd(q, v, idx, g, clip, scale, T, need_dq=True, need_dv=False, need_dT=False, flags=fl)
            outs[name] = dq
        torch.cuda.synchronize()
        ok1 = torch.equal(outs["pipe"], outs["staged"])
        ok2 = torch.equal(outs["pipe"], outs["generic"])
        fin = bool(torch.isfinite(outs["pipe"].float()).all())
        line = f"v{var} Bq={Bq} Bv={Bv} Nq={Nq} Nv={Nv}: pipe==staged {ok1} pipe==generic {ok2} finite {fin}"
        if iters:
            for name, fl in (("pipe", 0), ("staged", _lib.BWD_DQ_STAGED)):
                f = lambda: ops.maxmean_bwd(q, v, idx, g, clip, scale, T, need_dq=True, need_dv=False, need_dT=False, flags=fl)
                for _ in range(3):
                    f()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    f()
                e1.record()
                torch.cuda.synchronize()
                line += f"  {name} {e0.elapsed_time(e1) / iters:.4f} ms"
        print(line, flush=True)
        return ok1 and ok2 and fin

    good = True
    for shp in ((3, 5, 50, 256), (2, 1, 77, 173), (5, 2, 8, 64), (4, 3, 250, 256), (7, 9, 13, 40), (16, 16, 128, 256)):
        good &= run(*shp)
    good &= run(256, 256, 250, 256, iters=20)
    good &= run(128, 1024, 250, 256, iters=5)
    print(f"v{var} ALL_OK={good}", flush=True)


if __name__ == "__main__":
    if os.environ.get("DQ_AB_CHILD"):
        child()
    else:
        for v in (sys.argv[1:] or ["0", "1", "2", "3", "4"]):
            env = dict(os.environ, DQ_AB_CHILD="1", TRIAD_DQ_VARIANT=v)
            try:
                subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, check=False, timeout=150)
            except subprocess.TimeoutExpired:
                print(f"v{v} TIMEOUT", flush=True)
