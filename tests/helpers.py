"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))


def check_inputs_reproduce(gold, q, v):
    """The fixtures store checksums of the seeded inputs; a torch whose CPU generator differs
    would silently invalidate every comparison, so fail loudly instead."""
    assert abs(q.double().sum().item() - float(gold["in_q_sum"])) < 1e-9
    assert abs((q.double() ** 2).sum().item() - float(gold["in_q_sq"])) < 1e-9
    assert abs(v.double().sum().item() - float(gold["in_v_sum"])) < 1e-9
    assert abs((v.double() ** 2).sum().item() - float(gold["in_v_sq"])) < 1e-9


def _as64(x):
    if isinstance(x, torch.Tensor):
        return x.detach().to(device="cpu", dtype=torch.float64).reshape(-1)
    return torch.as_tensor(np.asarray(x), dtype=torch.float64).reshape(-1)


def rel_err(a, b):
    a, b = _as64(a), _as64(b)
    return ((a - b).norm() / b.norm().clamp(min=1e-30)).item()


