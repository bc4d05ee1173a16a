"""Oracle-backed CPU stand-ins for the five kernel entry points, injected into
triad_b200.dist by the gloo tests so the collective plumbing can be exercised without a GPU.
TEST INFRASTRUCTURE: the product binding (dist.CudaKernels) is CUDA-only."""
import torch

from oracle import oracle as O


class OracleKernels:
    def row_scale(self, mask, Bq, Nq, device):
        return O.row_scale_from_mask(mask, Bq, Nq).reshape(-1)

    def maxmean_fwd(self, q, v, scale, T):
        Bq, Nq, _ = q.shape
        Bv = v.shape[0]
        idx = torch.empty(Bq, Bv, Nq, dtype=torch.int64)
        rowmax = torch.empty(Bq, Bv, Nq)
        for i in range(Bq):
            s = O.token_sims_for_query(q[i], v, T)
            m, ix = torch.max(s, dim=2)
            idx[i], rowmax[i] = ix, m.float()
        clip = (rowmax * scale.view(Bq, 1, Nq)).sum(dim=2)
        pad = (Nq + 15) // 16 * 16                       # library layout: [Bv][Bq][nq_pad]
        lib = torch.zeros(Bv, Bq, pad, dtype=torch.uint8 if v.shape[1] <= 256 else torch.int32)
        lib[:, :, :Nq] = idx.permute(1, 0, 2).to(lib.dtype)
        return clip, lib.reshape(Bv, Bq * pad)

    def infonce_partial(self, clip_rows, B, row0):
        row_lse = torch.logsumexp(clip_rows, dim=1)
        m = clip_rows.max(dim=0).values
        s = torch.exp(clip_rows - m[None]).sum(dim=0)
        return row_lse, torch.stack([m, s])

    def infonce_finish(self, clip_rows, B, row0, row_lse, col_parts):
        rows = clip_rows.shape[0]
        m = col_parts[:, 0].max(dim=0).values
        s = (col_parts[:, 1] * torch.exp(col_parts[:, 0] - m[None])).sum(dim=0)
        col_lse = m + torch.log(s)
        c = clip_rows.double()
        eye = torch.zeros(rows, B, dtype=torch.bool)
        eye[torch.arange(rows), row0 + torch.arange(rows)] = True
        g = (torch.exp(c - row_lse.double()[:, None]) + torch.exp(c - col_lse.double()[None, :]) - 2 * eye) / (2 * B)
        d = c[eye]
        off = c[~eye]
        sums = torch.tensor([
            ((row_lse.double() - d) + (col_lse.double()[row0:row0 + rows] - d)).sum(),
            d.sum(), (d * d).sum(), off.sum(), (off * off).sum(), off.max(), (g * c).sum(), 0.0], dtype=torch.float64)
        return g.float(), sums

    def maxmean_bwd(self, q, v, idx, g, clip, scale, T):
        Bq, Nq, D = q.shape
        Bv, Nv, _ = v.shape
        pad = (Nq + 15) // 16 * 16
        idx_ref = idx.to(torch.int64).view(Bv, Bq, pad)[:, :, :Nq].permute(1, 0, 2)
        dq, dv, _ = O.maxmean_backward(q, v, idx_ref, g, float(T), scale.view(Bq, Nq), clip)
        return dq.to(q.dtype), dv.float()


def _nonneg_cpu(q, v, T, lo, numel):
    """fp64 autograd of sum clamp(T<q,v>, lo, 0)^2 / numel (the dense term of model.py:411-412 / :525-526)."""
    q64, v64 = q.double().requires_grad_(), v.double().requires_grad_()
    T64 = torch.tensor(float(T), dtype=torch.float64, requires_grad=True)
    tok = torch.einsum("iad,jpd->ijap", q64, v64) * T64
    s2 = tok.clamp(min=lo, max=0).pow(2).sum()
    (s2 / numel).backward()
    return s2.detach(), q64.grad.to(q.dtype), v64.grad.float(), T64.grad


OracleKernels.nonneg = lambda self, q, v, T, lo, numel: _nonneg_cpu(q, v, T, lo, numel)


def _pospair_cpu(q, v, T, kind, threshold):
    """fp64 autograd of the positive-pair term over these pairs (model.py:394-408 / :528-541), unweighted."""
    q64, v64 = q.double().requires_grad_(), v.double().requires_grad_()
    T64 = torch.tensor(float(T), dtype=torch.float64, requires_grad=True)
    diag = torch.bmm(q64, v64.transpose(1, 2)) * T64
    if kind == "av":
        term = ((diag[:, 1:] - diag[:, :-1]) ** 2).mean()
    else:
        probs = torch.softmax(diag, dim=-1)
        term = (torch.relu(probs.sum(dim=1) / probs.shape[1] - threshold) ** 2).mean()
    term.backward()
    return term.detach(), q64.grad.to(q.dtype), v64.grad.to(v.dtype), T64.grad


OracleKernels.pospair = lambda self, q, v, T, kind, threshold: _pospair_cpu(q, v, T, kind, threshold)


class PipelinedOracleKernels(OracleKernels):
    """Adds the split dv / dq entry points, so the sharded step takes its pipelined (per-destination reduce) path."""

    def maxmean_bwd_dv(self, q, v, idx, g, scale, T):
        Bq, Nq, D = q.shape
        clip = torch.zeros(Bq, v.shape[0])
        return self.maxmean_bwd(q, v, idx, g, clip, scale, T)[1]

    def maxmean_bwd_dq(self, q, v, idx, g, scale, T):
        Bq = q.shape[0]
        clip = torch.zeros(Bq, v.shape[0])
        return self.maxmean_bwd(q, v, idx, g, clip, scale, T)[0]
