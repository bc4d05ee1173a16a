"""Loader of the staged reference (oracle/_ref/, see build_ref.py) — TEST / BASELINE INFRASTRUCTURE ONLY.

load() imports the reference's unmodified model.py / retrieval.py (an empty `peft` stub module is registered
first: the package is not installed and the hot-path methods never touch it) and returns
(model_module, retrieval_module), or None when oracle/_ref/ has not been staged.  make_stub() builds the object
the reference's methods are called on, unbound: they only read self.temperature and the two patch_sparsity_*
floats (src/model.py:348-351)."""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_loaded = None


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "MANIFEST.json"))


def load():
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        return None
    manifest = json.load(open(os.path.join(REF_DIR, "MANIFEST.json")))
    for name, meta in manifest.items():
        with open(os.path.join(REF_DIR, name), "rb") as f:
            if hashlib.sha256(f.read()).hexdigest() != meta["sha256"]:
                raise RuntimeError(f"oracle/_ref/{name} does not match its manifest (re-run oracle/build_ref.py)")
    peft = types.ModuleType("peft")
    for n in ("LoraConfig", "get_peft_model", "TaskType"):
        setattr(peft, n, object)
    sys.modules.setdefault("peft", peft)
    mods = []
    for name in ("model", "retrieval"):
        spec = importlib.util.spec_from_file_location(f"triad_reference_{name}", os.path.join(REF_DIR, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods.append(mod)
    _loaded = tuple(mods)
    return _loaded


def make_stub(M, T: float, regularizers: bool = True, device="cpu"):
    """The `self` of the reference's hot-path methods.  regularizers=False binds zero-returning regularisers, so that
    compute_contrastive_loss_{av,tv} run their own similarity-statistics and InfoNCE lines (src/model.py:430-459,
    :544-578) and nothing else — the like-for-like of the fused max-mean + InfoNCE metric."""
    class Stub:
        pass
    s = Stub()
    s.temperature = torch.nn.Parameter(torch.tensor(float(T), device=device))
    s.patch_sparsity_threshold, s.patch_sparsity_weight = 0.80, 0.01
    if regularizers:
        for n in ("compute_temporal_smoothness_loss", "compute_regularization_losses_av",
                  "compute_regularization_losses_tv"):
            setattr(s, n, types.MethodType(getattr(M, n), s))
    else:
        zero = lambda: torch.zeros((), device=device)                                        # noqa: E731
        s.compute_regularization_losses_av = lambda token_sims: (zero(), zero())
        s.compute_regularization_losses_tv = lambda token_sims: zero()
    return s
