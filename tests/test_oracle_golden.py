"""CPU tier: the oracle (oracle/oracle.py) against the golden vectors produced by the
UNMODIFIED reference (oracle/gen_golden.py -> tests/golden/*.npz).  This is what pins the
checker that the CUDA path is later compared with."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle.cases import CASES, build_inputs, projection
from tests.helpers import check_inputs_reproduce, load_golden, rel_err


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_forward_matches_reference(case):
    gold = load_golden(case.name)
    q, v, mask, T = build_inputs(case)
    check_inputs_reproduce(gold, q, v)
    out = O.maxmean_forward(q, v, T, mask)
    # argmax patch indices: bit-exact (torch.max first-index tie-break, model.py:389)
    assert np.array_equal(out["idx"].numpy(), gold["idx"].astype(np.int64))
    # rounded row maxima and clip_sims in the reference's dtype: bit-exact
    assert np.array_equal(out["rowmax"].numpy(), gold["rowmax"])
    assert np.array_equal(out["clip_ref"].float().numpy(), gold["clip"])
    assert bool(gold["clip_is_bf16"]) == (out["clip_ref"].dtype == torch.bfloat16)


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_loss_and_gradients_match_reference(case):
    gold = load_golden(case.name)
    q, v, mask, T = build_inputs(case)
    out = O.contrastive_step_closed_form(q, v, T, mask)
    P = projection(case)
    dq = out["dq"] @ P if P is not None else out["dq"]
    dv = out["dv"] @ P if P is not None else out["dv"]
    fp32 = case.dtype == "fp32"
    # north-star tolerances: 1e-4 relative in fp32, 1e-2 in bf16 (the reference's bf16 autograd
    # rounds every intermediate, the oracle accumulates in fp64)
    tol = 1e-4 if fp32 else 1e-2
    assert abs(out["loss"].item() - float(gold["contrastive"])) <= tol * abs(float(gold["contrastive"]))
    assert rel_err(dq, gold["dq"]) < tol
    assert rel_err(dv, gold["dv"]) < tol
    # dT = sum g*clip / T is a sum of cancelling terms (sum_j g_ij ~ 0): compare against the
    # scale of the terms, not of the tiny result
    scale = (out["g"].abs() * out["clip"].abs().double()).sum().item() / T
    assert abs(out["dT"].item() - float(gold["dT"])) < (1e-4 if fp32 else 2e-2) * scale


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_stats_match_reference(case):
    gold = load_golden(case.name)
    q, v, mask, T = build_inputs(case)
    out = O.maxmean_forward(q, v, T, mask)
    stats = O.similarity_stats(out["clip_ref"], case.kind)
    keys = [str(k) for k in gold["stats_keys"]]
    assert sorted(stats.keys()) == keys
    # the reference takes these statistics on its clip_sims dtype: in bf16 a mean of ~0.4 is only
    # known to 2^-9, so differences of means (separation) are compared on the scale of the means
    scale = max(abs(float(x)) for x in gold["stats_vals"])
    tol = 1e-2 if case.dtype == "bf16" and case.kind == "av" else 1e-4
    for k, ref in zip(keys, gold["stats_vals"]):
        assert abs(stats[k] - float(ref)) <= tol * scale, k


@pytest.mark.parametrize("case", [c for c in CASES if c.B <= 6], ids=lambda c: c.name)
def test_total_loss_with_regularisers(case):
    """the dense regularisers, restated, reproduce the reference's total loss."""
    gold = load_golden(case.name)
    q, v, mask, T = build_inputs(case)
    Tt = torch.tensor(T)
    tok = torch.stack([O.token_sims_for_query(q[i], v, T) for i in range(case.B)])   # (B,B,Nq,Nv)
    if case.kind == "av":
        reg, smooth = O.regularization_av(tok, Tt)
        assert abs(float(smooth) - float(gold["smooth"])) <= 1e-2 * abs(float(gold["smooth"])) + 1e-9
    else:
        reg = O.regularization_tv(tok, 0.80, 0.01)
    tol = 1e-4 if case.dtype == "fp32" else 2e-2
    assert abs(float(reg) - float(gold["reg"])) <= tol * abs(float(gold["reg"])) + 1e-7


def test_autograd_port_matches_closed_form():
    """The CPU-baseline port (materialises token_sims, autograd) and the closed-form oracle agree."""
    q, v, mask = O.make_inputs(5, 12, 24, 32, torch.float32, seed=9, masked=True)
    T = torch.tensor(1.5, requires_grad=True)
    qa, va = q.clone().requires_grad_(), v.clone().requires_grad_()
    loss, clip, _ = O.reference_step_autograd(qa, va, T, mask)
    ref = O.contrastive_step_closed_form(q, v, 1.5, mask)
    assert abs(loss.item() - ref["loss"].item()) < 1e-5
    assert rel_err(qa.grad, ref["dq"]) < 1e-5 and rel_err(va.grad, ref["dv"]) < 1e-5
    assert abs(T.grad.item() - ref["dT"].item()) < 1e-6


def test_retrieval_goldens():
    gold = load_golden("retrieval")
    g = torch.Generator().manual_seed(77)
    row = 0
    for n, (nq, nv, d) in enumerate(gold["agg_shapes"].tolist()):
        qf = torch.randn(nq, d, generator=g)
        vf = torch.randn(nv, d, generator=g)
        if n != 1:
            qf = torch.nn.functional.normalize(qf, dim=1)
            vf = torch.nn.functional.normalize(vf, dim=1)
        for T in gold["agg_T"].tolist():
            ref = gold["agg_vals"][row]
            row += 1
            assert abs(O.aggregate_pair(qf, vf, T, "q2v") - ref[0]) < 1e-6
            assert abs(O.aggregate_pair(qf, vf, T, "v2q") - ref[1]) < 1e-6
            assert abs(O.aggregate_pair(qf, vf, T, "q2v") - ref[2]) < 1e-6
            assert abs(O.aggregate_pair(qf, vf, T, "v2q") - ref[3]) < 1e-6
    rec = O.recall_at_k(gold["recall_sim"])
    assert np.allclose([rec["r1"], rec["r5"], rec["r10"], rec["r20"]], gold["recall_vals"])
    sim = torch.randn(60, 60, generator=g)      # consume the generator like gen_golden did
    f1 = torch.randn(3, 9, 32, generator=g)
    f2 = torch.randn(3, 17, 32, generator=g)
    assert np.allclose(O.similarity_matrix(f1, f2, 1.5).numpy(), gold["simmat"], atol=1e-6)


def test_edge_cases():
    # single query / single image / one token / fully masked row
    q, v, _ = O.make_inputs(1, 1, 3, 8, torch.float32, seed=1)
    out = O.contrastive_step_closed_form(q, v, 1.5)
    assert out["loss"].abs().item() < 1e-12 and out["dq"].abs().max().item() < 1e-12
    q, v, mask = O.make_inputs(3, 5, 4, 8, torch.float32, seed=2, masked=True)
    mask[1] = 0                                   # no valid token: clamp(min=1e-7) keeps it finite
    out = O.maxmean_forward(q, v, 1.5, mask)
    assert torch.isfinite(out["clip"]).all() and out["clip"][1].abs().max().item() == 0.0
