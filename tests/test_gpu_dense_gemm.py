"""GPU tier: triad_dense_grad_gemm (the dense regulariser's two backward GEMMs, hand-written tcgen05 with MN-major
operands) against fp32 matmuls of the same bf16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [
    # M (rows of N), Kc (columns of N), D
    (256, 256, 512),
    (300, 1536, 512),          # row / K tails
    (1000, 520, 256),          # one N half, Kc % 64 != 0
    (77, 200, 64),             # tiny: zero-filled B columns, partial tiles everywhere
    (1280, 4096, 128),
    (2500, 3072, 512),
]


@pytest.mark.parametrize("shape", SHAPES, ids=["x".join(map(str, s)) for s in SHAPES])
@pytest.mark.parametrize("mode", [0, 1])
def test_dense_grad_gemm_vs_fp32(shape, mode):
    from triad_b200 import ops
    M, Kc, D = shape
    g = torch.Generator().manual_seed(M + Kc + D + mode)
    N = (torch.randn(M, Kc, generator=g) * 0.05).bfloat16().cuda()
    x = torch.randn(Kc if mode == 0 else M, D, generator=g).bfloat16().cuda()
    out = ops.dense_grad_gemm(N, x, mode)
    ref = (N.float() @ x.float()) if mode == 0 else (N.float().t() @ x.float())
    torch.cuda.synchronize()
    assert out.shape == ref.shape
    err = ((out.float() - ref).norm() / ref.norm()).item()
    assert err < 4e-3, err                                  # one bf16 rounding of the output
    worst = ((out.float() - ref).abs() / (ref.abs() + 1e-2 * ref.abs().max())).max().item()
    assert worst < 2e-2, worst


def test_dense_grad_gemm_padded_pitch_and_determinism():
    """N with a row pitch larger than its width (a column slice of a bigger buffer); two runs are bit-identical."""
    from triad_b200 import ops
    M, Kc, D = 700, 1000, 512
    g = torch.Generator().manual_seed(9)
    big = (torch.randn(M, Kc + 24, generator=g) * 0.05).bfloat16().cuda()
    N = big[:, :Kc]
    for mode in (0, 1):
        x = torch.randn(Kc if mode == 0 else M, D, generator=g).bfloat16().cuda()
        a = ops.dense_grad_gemm(N, x, mode)
        b = ops.dense_grad_gemm(N, x, mode)
        ref = (N.float() @ x.float()) if mode == 0 else (N.float().t() @ x.float())
        assert torch.equal(a, b)
        assert ((a.float() - ref).norm() / ref.norm()).item() < 4e-3
