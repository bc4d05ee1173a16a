"""CPU tier: the rounding / first-argmax-threshold model shared by the kernels
(triad_b200/csrc/triad_round.h) is exhaustively self-tested on the host."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_threshold_is_minimal_and_exact(tmp_path):
    exe = tmp_path / "round_selftest"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tools", "round_selftest.cpp")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "bad=0" in r.stdout


def test_rounding_model_matches_torch():
    """bf16(bf16(acc)*T) written with integer ops == what torch does to the reference's token_sims."""
    g = torch.Generator().manual_seed(0)
    acc = torch.randn(200000, generator=g) * 3
    T = torch.tensor(1.5)
    want = (acc.bfloat16().float() * T).bfloat16().float().numpy()
    u = acc.numpy().view(np.uint32).astype(np.uint64)
    def rn(u):
        return ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    b = rn(u).view(np.float32)
    prod = (b * np.float32(1.5)).astype(np.float32)
    got = rn(prod.view(np.uint32).astype(np.uint64)).view(np.float32)
    assert np.array_equal(got, want)
