#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/src) on
the seeded inputs of oracle/cases.py.  Runs only in the build container (the reference does
not exist on the GPU box); the resulting fixtures are committed.

    python oracle/gen_golden.py            # rewrites every fixture

The reference's model.py imports `peft`, which is not installed; an empty stub module is
registered first.  MultiModalModel cannot be constructed offline (weights
need the network), and need not be: the hot-path methods only read self.temperature and the
two patch_sparsity_* floats, so they are called unbound on a stub object.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.cases import CASES, build_inputs, projection  # noqa: E402

REF_SRC = "/root/reference/src"


def load_reference():
    peft = types.ModuleType("peft")
    for n in ("LoraConfig", "get_peft_model", "TaskType"):
        setattr(peft, n, object)
    sys.modules.setdefault("peft", peft)
    sys.path.insert(0, REF_SRC)
    import model as ref_model          # noqa: E402
    import retrieval as ref_retrieval  # noqa: E402
    return ref_model, ref_retrieval


def make_stub(M, T: float):
    class Stub:
        pass
    s = Stub()
    s.temperature = torch.nn.Parameter(torch.tensor(float(T)))
    s.patch_sparsity_threshold, s.patch_sparsity_weight = 0.80, 0.01
    for n in ("compute_temporal_smoothness_loss", "compute_regularization_losses_av",
              "compute_regularization_losses_tv"):
        setattr(s, n, types.MethodType(getattr(M, n), s))
    return s


def f32(t):
    return t.detach().to(torch.float32).numpy()


def proj(t, P):
    t64 = t.detach().to(torch.float64)
    return (t64 @ P).numpy() if P is not None else t64.numpy()


def run_case(M, c):
    q, v, mask, T = build_inputs(c)
    q.requires_grad_(True)
    v.requires_grad_(True)
    s = make_stub(M, T)
    if c.kind == "av":
        clip, tok = M.compute_all_similarities_av(s, q, v)
        total, contrastive, reg, smooth, stats = M.compute_contrastive_loss_av(s, clip, tok)
    else:
        clip, tok = M.compute_all_similarities_tv(s, q, v, mask)
        total, stats = M.compute_contrastive_loss_tv(s, clip, tok)
        ar = torch.arange(c.B)
        contrastive = (F.cross_entropy(clip, ar) + F.cross_entropy(clip.t(), ar)) / 2
        reg, smooth = total - contrastive, torch.zeros(())
    best, idx = torch.max(tok, dim=3)
    P = projection(c)

    contrastive.backward(retain_graph=True)
    dq_c, dv_c, dT_c = q.grad.clone(), v.grad.clone(), s.temperature.grad.clone()
    q.grad = None
    v.grad = None
    s.temperature.grad = None
    total.backward()
    dq_t, dv_t, dT_t = q.grad.clone(), v.grad.clone(), s.temperature.grad.clone()

    out = {
        "in_q_sum": np.float64(q.detach().double().sum().item()),
        "in_q_sq": np.float64((q.detach().double() ** 2).sum().item()),
        "in_v_sum": np.float64(v.detach().double().sum().item()),
        "in_v_sq": np.float64((v.detach().double() ** 2).sum().item()),
        "clip": f32(clip),
        "clip_is_bf16": np.bool_(clip.dtype == torch.bfloat16),
        "idx": idx.numpy().astype(np.uint16),
        "rowmax": f32(best),
        "contrastive": np.float64(contrastive.item()),
        "total": np.float64(total.item()),
        "reg": np.float64(float(reg)),
        "smooth": np.float64(float(smooth)),
        "stats_keys": np.array(sorted(stats.keys())),
        "stats_vals": np.array([stats[k] for k in sorted(stats.keys())], dtype=np.float64),
        "dq": proj(dq_c, P), "dv": proj(dv_c, P), "dT": np.float64(dT_c.item()),
        "dq_total": proj(dq_t, P), "dv_total": proj(dv_t, P), "dT_total": np.float64(dT_t.item()),
    }
    if mask is not None:
        out["mask"] = mask.numpy().astype(np.uint8)
    return out


def run_retrieval(M, R):
    """aggregator_* (retrieval.py:106-115,190-198), compute_recall_at_k (:117-144) and
    compute_similarity_matrix (model.py:355-368) on seeded inputs."""
    g = torch.Generator().manual_seed(77)
    out = {}
    Ts = [0.7, 1.5]
    pairs = []
    for n, (nq, nv, d) in enumerate([(50, 256, 512), (13, 40, 64), (1, 7, 32), (250, 1024, 512)]):
        qf = torch.randn(nq, d, generator=g)
        vf = torch.randn(nv, d, generator=g)
        if n != 1:
            qf, vf = F.normalize(qf, dim=1), F.normalize(vf, dim=1)   # AV feats are normalised (:93-94)
        for T in Ts:
            pairs.append([R.aggregator_av_a2v(qf, vf, T), R.aggregator_av_v2a(qf, vf, T),
                          R.aggregator_tv_t2v(qf, vf, T), R.aggregator_tv_v2t(qf, vf, T)])
    out["agg_shapes"] = np.array([(50, 256, 512), (13, 40, 64), (1, 7, 32), (250, 1024, 512)])
    out["agg_T"] = np.array(Ts)
    out["agg_vals"] = np.array(pairs, dtype=np.float64)
    sim = torch.randn(60, 60, generator=g)
    sim += torch.eye(60) * 1.5
    sim[3, 9] = sim[3, 3]            # an exact tie with the diagonal
    rec = R.compute_recall_at_k(sim.numpy())
    out["recall_sim"] = sim.numpy()
    out["recall_vals"] = np.array([rec["r1"], rec["r5"], rec["r10"], rec["r20"]])
    s = make_stub(M, 1.5)
    f1 = torch.randn(3, 9, 32, generator=g)
    f2 = torch.randn(3, 17, 32, generator=g)
    out["simmat"] = M.compute_similarity_matrix(s, f1, f2).detach().numpy()
    return out


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref_model, ref_retrieval = load_reference()
    M = ref_model.MultiModalModel
    dst = os.path.join(ROOT, "tests", "golden")
    os.makedirs(dst, exist_ok=True)
    for c in CASES:
        out = run_case(M, c)
        np.savez_compressed(os.path.join(dst, c.name + ".npz"), **out)
        print(f"{c.name:16s} contrastive={out['contrastive']:.6f} total={out['total']:.6f} "
              f"dT={out['dT']:.6f}")
    np.savez_compressed(os.path.join(dst, "retrieval.npz"), **run_retrieval(M, ref_retrieval))
    print("retrieval ok")
    with open(os.path.join(dst, "PROVENANCE.txt"), "w") as f:
        f.write("Generated by oracle/gen_golden.py from the unmodified reference at /root/reference/src\n"
                f"torch {torch.__version__}, numpy {np.__version__}, CPU, threads={torch.get_num_threads()}\n")


if __name__ == "__main__":
    main()
