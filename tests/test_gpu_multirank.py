"""GPU tier, needs >= 2 GPUs (skipped on a one-GPU box): the row-sharded step under NCCL with the product CUDA
kernels — every rank's loss / clip rows / dq / dv / dT share against the single-GPU drop-in on the gathered batch and,
at a small size, against the CPU oracle (bench.verify_sharded); and the gallery-sharded retrieval against the
single-GPU top-k."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(n, *args):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", *args]
    return subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_sharded_step_parity_under_nccl():
    n = 2 if torch.cuda.device_count() < 8 else 8
    r = _torchrun(n, "bench.py", "--gpus", str(n), "--verify")
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["parity"] == "ok" and out["parity_nranks"] == n
    w = out["worst_rel_err_vs_single_gpu"]
    assert w["loss"] < 1e-6 and w["dq"] < 2e-3 and w["dv"] < 2e-3 and w["clip"] < 2e-6


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_sharded_retrieval_under_nccl():
    r = _torchrun(2, os.path.join("tools", "sharded_retrieval_check.py"))
    assert r.returncode == 0, r.stderr[-3000:]
    assert "SHARDED_RETRIEVAL_OK" in r.stdout


def test_sharded_retrieval_single_rank_cuda():
    """world size 1: sharded_retrieve_topk == retrieve_topk (the merge keeps the library's order, ties included)."""
    from triad_b200 import retrieval as R
    from triad_b200.dist import sharded_retrieve_topk
    g = torch.Generator(device="cuda").manual_seed(5)
    q = torch.nn.functional.normalize(torch.randn(77, 128, generator=g, device="cuda"), dim=1).bfloat16()
    gal = torch.nn.functional.normalize(torch.randn(300, 64, 128, generator=g, device="cuda"), dim=2).bfloat16()
    gal[250] = gal[17]
    gal[40] = gal[17]
    s0, i0 = R.retrieve_topk(q, gal, 1.5, 12)
    s1, i1 = sharded_retrieve_topk(q, gal, 1.5, 12, 0)
    assert torch.equal(i1, i0.to(torch.int64)) and torch.equal(s1, s0)
    # two "shards" merged by hand: the same ids as the whole gallery
    from triad_b200.dist import merge_topk
    sa, ia = R.retrieve_topk(q, gal[:130].contiguous(), 1.5, 12)
    sb, ib = R.retrieve_topk(q, gal[130:].contiguous(), 1.5, 12)
    sm, im = merge_topk(torch.cat([sa, sb]), torch.cat([ia.to(torch.int64), ib.to(torch.int64) + 130]), 12)
    assert torch.equal(im, i0.to(torch.int64)) and torch.equal(sm, s0)
