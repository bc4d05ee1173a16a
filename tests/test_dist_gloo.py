"""CPU tier: the row-sharded step (triad_b200/dist.py) over a world_size-2 gloo group gives the
same loss, statistics and gradients as the single-process oracle on the full batch.  Exercises
the host-side sharding + the three all-gathers + the reduce-scatter; the kernels are the
oracle-backed stand-ins of tests/cpu_kernels.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, masked, ret, pipelined=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.cpu_kernels import OracleKernels, PipelinedOracleKernels
        from triad_b200.dist import sharded_contrastive_step, stats_from_all_sums
        B, Nq, Nv, D = 6, 9, 20, 16
        q, v, mask = O.make_inputs(B, Nq, Nv, D, torch.float32, seed=4, masked=masked)
        Bl = B // world
        sl = slice(rank * Bl, (rank + 1) * Bl)
        out = sharded_contrastive_step(q[sl], v[sl], torch.tensor(1.5), mask[sl] if masked else None,
                                       kernels=PipelinedOracleKernels() if pipelined else OracleKernels())
        stats = stats_from_all_sums(out["all_sums"], B, "av")
        ret[rank] = {"loss": out["loss"].item(), "dq": out["dq"], "dv": out["dv"], "dT": out["dT"].item(),
                     "dT_global": out["dT_global"].item(), "stats": stats}
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("pipelined", [False, True])
@pytest.mark.parametrize("masked", [False, True])
def test_two_rank_step_matches_full_batch(masked, pipelined):
    """pipelined=False: one reduce-scatter of the dv partial (kernels without the split entry points);
    pipelined=True: the product path — per-destination reduces overlapped with the next dv chunk and dq."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), masked, ret, pipelined), nprocs=world, join=True)
    B, Nq, Nv, D = 6, 9, 20, 16
    q, v, mask = O.make_inputs(B, Nq, Nv, D, torch.float32, seed=4, masked=masked)
    ref = O.contrastive_step_closed_form(q, v, 1.5, mask)
    st = O.similarity_stats(ref["clip"], "av")
    Bl = B // world
    for r in range(world):
        o = ret[r]
        assert abs(o["loss"] - ref["loss"].item()) < 1e-5
        assert abs(o["dT_global"] - ref["dT"].item()) < 1e-6
        assert torch.allclose(o["dq"].double(), ref["dq"][r * Bl:(r + 1) * Bl], atol=1e-7)
        assert torch.allclose(o["dv"].double(), ref["dv"][r * Bl:(r + 1) * Bl], atol=1e-7)
        for k, val in st.items():
            assert abs(o["stats"][k] - val) < 1e-5, k
    # the temperature is replicated: every rank returns its SHARE, like the gradients that reach replicated encoder
    # weights through dq / dv — a SUM over ranks is the global gradient (and the shares are not all equal)
    assert abs(sum(ret[r]["dT"] for r in range(world)) - ref["dT"].item()) < 1e-6
    g, clip = ref["g"].double(), ref["clip"].double()
    for r in range(world):
        share = (g[r * Bl:(r + 1) * Bl] * clip[r * Bl:(r + 1) * Bl]).sum().item() / 1.5
        assert abs(ret[r]["dT"] - share) < 1e-6


def test_single_process_path():
    from tests.cpu_kernels import OracleKernels
    from triad_b200.dist import sharded_contrastive_step
    q, v, _ = O.make_inputs(4, 5, 7, 8, torch.float32, seed=6)
    out = sharded_contrastive_step(q, v, torch.tensor(1.2), None, kernels=OracleKernels())
    ref = O.contrastive_step_closed_form(q, v, 1.2)
    assert abs(out["loss"].item() - ref["loss"].item()) < 1e-5
    assert torch.allclose(out["dq"].double(), ref["dq"], atol=1e-7)


def _reg_worker(rank, world, port, kind, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.cpu_kernels import OracleKernels
        from triad_b200.dist import sharded_regularizer_step
        B, Nq, Nv, D = 6, 9, 20, 16
        q, v, _ = O.make_inputs(B, Nq, Nv, D, torch.float32, seed=8)
        q, v = q * 3, v * 3                                   # some similarities below zero by a margin
        Bl = B // world
        sl = slice(rank * Bl, (rank + 1) * Bl)
        out = sharded_regularizer_step(q[sl], v[sl], torch.tensor(0.9), kind, kernels=OracleKernels(),
                                       patch_sparsity_threshold=0.02, patch_sparsity_weight=0.5)
        ret[rank] = {k: (val.item() if val.dim() == 0 else val) for k, val in out.items()}
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["av", "tv"])
def test_two_rank_regularisers_match_full_batch(kind):
    """sharded_regularizer_step over 2 ranks == the oracle's restatement of model.py:394-428 / :516-542 on the
    whole batch (value and the gradients w.r.t. every shard and the temperature), fp64 autograd."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_reg_worker, args=(world, _free_port(), kind, ret), nprocs=world, join=True)
    B, Nq, Nv, D = 6, 9, 20, 16
    q, v, _ = O.make_inputs(B, Nq, Nv, D, torch.float32, seed=8)
    q64, v64 = (q * 3).double().requires_grad_(), (v * 3).double().requires_grad_()
    T64 = torch.tensor(0.9, dtype=torch.float64, requires_grad=True)     # T < 1: the calibration term is active
    tok = torch.einsum("iad,jpd->ijap", q64, v64) * T64
    if kind == "av":
        reg, smooth = O.regularization_av(tok, T64)
    else:
        reg = O.regularization_tv(tok, 0.02, 0.5)
    reg.backward()
    Bl = B // world
    for r in range(world):
        o = ret[r]
        assert abs(o["reg"] - reg.item()) < 1e-5 * abs(reg.item())
        assert abs(o["dT_global"] - T64.grad.item()) < 1e-4 * abs(T64.grad.item())
        assert torch.allclose(o["dq"].double(), q64.grad[r * Bl:(r + 1) * Bl], rtol=1e-4, atol=1e-7)
        assert torch.allclose(o["dv"].double(), v64.grad[r * Bl:(r + 1) * Bl], rtol=1e-4, atol=1e-7)
        if kind == "av":
            assert abs(o["smooth"] - smooth.item()) < 1e-5 * abs(smooth.item())
    assert abs(sum(ret[r]["dT"] for r in range(world)) - T64.grad.item()) < 1e-4 * abs(T64.grad.item())


# ---- gallery-sharded retrieval (cfg 5): local top-k + all-gather + merge ------------------------------
def _oracle_topk(q, gallery, temperature, k, direction):
    """CPU stand-in of triad_b200.retrieval.retrieve_topk: scores with the oracle's aggregator (retrieval.py:106-115),
    top-k in the library's order (score descending, ties to the lower id)."""
    s = torch.tensor([O.aggregate_pair(q, gallery[n], temperature, "q2v" if direction == 0 else "v2q")
                      for n in range(gallery.shape[0])], dtype=torch.float32)
    order = torch.argsort(s, descending=True, stable=True)[:k]
    return s[order], order.to(torch.int32)


def _gallery(n_img, Nq, Nv, D):
    g = torch.Generator().manual_seed(21)
    q = torch.nn.functional.normalize(torch.randn(Nq, D, generator=g), dim=1)
    gal = torch.nn.functional.normalize(torch.randn(n_img, Nv, D, generator=g), dim=2)
    if n_img > 9:
        gal[9] = gal[2]                  # exact ties across (and inside) shards: the lower id must win
        gal[5] = gal[2]
    else:
        gal[n_img - 1] = gal[0]
    return q, gal


def _retrieve_worker(rank, world, port, n_img, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from triad_b200.dist import sharded_retrieve_topk
        q, gal = _gallery(n_img, 6, 10, 16)
        base, rem = n_img // world, n_img % world            # uneven shards
        n_loc = base + (1 if rank < rem else 0)
        id0 = rank * base + min(rank, rem)
        s, ids = sharded_retrieve_topk(q, gal[id0:id0 + n_loc], 1.5, k, id0, local_topk=_oracle_topk)
        ret[rank] = (s, ids)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_img,k", [(11, 4), (3, 5)])
def test_two_rank_sharded_retrieval_matches_single_gallery(n_img, k):
    """ids bit-equal to the top-k over the whole gallery, on every rank; k larger than a shard (and than the gallery)
    pads with -inf candidates that never win over real ones."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_retrieve_worker, args=(world, _free_port(), n_img, k, ret), nprocs=world, join=True)
    q, gal = _gallery(n_img, 6, 10, 16)
    want_s, want_i = _oracle_topk(q, gal, 1.5, min(k, n_img), 0)
    for r in range(world):
        s, ids = ret[r]
        n = min(k, n_img)
        assert torch.equal(ids[:n], want_i.to(torch.int64)), (ids, want_i)
        assert torch.equal(s[:n], want_s)
        assert bool(torch.isinf(s[n:]).all())
