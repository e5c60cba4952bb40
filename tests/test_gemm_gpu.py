"""tcgen05 GEMM core vs torch on the GPU: bit-exact on integer-valued operands (fp32 accumulation is
exact there, so any descriptor/swizzle/tile bug shows as a hard mismatch), tolerance on random ones."""
import pytest
import torch

from livecc_b200 import _cabi as A

pytestmark = pytest.mark.gpu

SHAPES = [
    (128, 128, 64), (128, 256, 128), (256, 512, 256), (100, 136, 72), (281, 4608, 3584),
    (1024, 1280, 1176), (1024, 3840, 1280), (3072, 5120, 1280), (33, 64, 3584), (130, 264, 200),
]


def _ints(shape, lo, hi, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randint(lo, hi + 1, shape, device="cuda", generator=g).to(torch.bfloat16)


@pytest.mark.parametrize("block_n", [0, 32, 48, 64, 80, 112, 128, 144, 176, 208, 224, 240, 256, -64, -96, -128, -160, -224, -256])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_exact_integer_operands(ctx, M, N, K, block_n):
    a = _ints((M, K), -2, 2, 1)
    b = _ints((N, K), -2, 2, 2)
    ref = (a.float() @ b.float().T).to(torch.bfloat16)
    out = ctx.gemm(a, b, block_n=block_n)
    torch.cuda.synchronize()
    assert torch.equal(out, ref), f"mismatch: {(out.float() - ref.float()).abs().max().item()}"


@pytest.mark.parametrize("M,N,K", [(281, 4608, 3584), (1024, 1280, 5120), (200, 136, 72)])
def test_gemm_bias_residual_exact(ctx, M, N, K):
    a = _ints((M, K), -2, 2, 3)
    b = _ints((N, K), -1, 1, 4)
    bias = _ints((N,), -8, 8, 5)
    res = _ints((M, N), -8, 8, 6)
    acc = a.float() @ b.float().T
    out = ctx.gemm(a, b, bias=bias, epilogue=A.EPI_BIAS)
    assert torch.equal(out, (acc + bias.float()).to(torch.bfloat16))
    out = ctx.gemm(a, b, residual=res, epilogue=A.EPI_RESIDUAL)
    assert torch.equal(out, acc.to(torch.bfloat16) + res)
    out = ctx.gemm(a, b, bias=bias, residual=res, epilogue=A.EPI_BIAS_RESIDUAL)
    assert torch.equal(out, (acc + bias.float()).to(torch.bfloat16) + res)
    # in-place residual (out aliases residual), the way the engine uses it
    res2 = res.clone()
    ctx.gemm(a, b, out=res2, residual=res2, epilogue=A.EPI_RESIDUAL)
    assert torch.equal(res2, acc.to(torch.bfloat16) + res)


def _close(out, ref, ulps=2):
    out, ref = out.float(), ref.float()
    tol = ulps * 2.0 ** -8 * ref.abs().clamp_min(1e-2)
    bad = (out - ref).abs() > tol
    return bad.float().mean().item()


@pytest.mark.parametrize("M,N,K", [(1024, 5120, 1280), (256, 3584, 5120), (77, 320, 136)])
def test_gemm_activation_epilogues(ctx, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(7)
    a = torch.randn((M, K), device="cuda", generator=g).to(torch.bfloat16)
    b = (torch.randn((N, K), device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    bias = (torch.randn((N,), device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    lin = (a.float() @ b.float().T + bias.float()).to(torch.bfloat16)
    out = ctx.gemm(a, b, bias=bias, epilogue=A.EPI_BIAS_QUICKGELU)
    ref = lin * torch.sigmoid(1.702 * lin)
    assert _close(out, ref) < 1e-3
    out = ctx.gemm(a, b, bias=bias, epilogue=A.EPI_BIAS_GELU)
    assert _close(out, torch.nn.functional.gelu(lin)) < 1e-3
    out = ctx.gemm(a, b, bias=bias, epilogue=A.EPI_BIAS)
    assert _close(out, lin, ulps=1) < 5e-3  # accumulation-order flips at the bf16 rounding boundary
    assert _close(out, lin, ulps=2) == 0


@pytest.mark.parametrize("M,I,K", [(281, 18944, 3584), (64, 2432, 896), (300, 96, 64)])
def test_gemm_swiglu(ctx, M, I, K):
    from livecc_b200.checkpoint import interleave_gate_up

    g = torch.Generator(device="cuda").manual_seed(8)
    a = torch.randn((M, K), device="cuda", generator=g).to(torch.bfloat16)
    wg = (torch.randn((I, K), device="cuda", generator=g) * 0.03).to(torch.bfloat16)
    wu = (torch.randn((I, K), device="cuda", generator=g) * 0.03).to(torch.bfloat16)
    gate = (a.float() @ wg.float().T).to(torch.bfloat16)
    up = (a.float() @ wu.float().T).to(torch.bfloat16)
    ref = torch.nn.functional.silu(gate) * up
    out = ctx.gemm(a, interleave_gate_up(wg, wu), epilogue=A.EPI_SWIGLU)
    assert out.shape == (M, I)
    assert _close(out, ref) < 1e-3


@pytest.mark.parametrize("block_n", [32, 96, 160, 224, -64, -160, -256])
def test_gemm_swiglu_every_tile_width(ctx, block_n):
    """The N tile is a run-time value (multiples of 32 for the SwiGLU epilogue): every width gives the same bits as 256.
    Negative widths select the CTA-pair kernel (cta_group::2, 256 x |block_n| tiles)."""
    from livecc_b200.checkpoint import interleave_gate_up

    g = torch.Generator(device="cuda").manual_seed(18)
    M, I, K = 281, 1184, 512
    a = torch.randn((M, K), device="cuda", generator=g).to(torch.bfloat16)
    w = interleave_gate_up((torch.randn((I, K), device="cuda", generator=g) * 0.03).to(torch.bfloat16),
                           (torch.randn((I, K), device="cuda", generator=g) * 0.03).to(torch.bfloat16))
    assert torch.equal(ctx.gemm(a, w, epilogue=A.EPI_SWIGLU, block_n=block_n), ctx.gemm(a, w, epilogue=A.EPI_SWIGLU, block_n=256))


@pytest.mark.parametrize("block_n", [48, 80, 144, 240, -96, -192, -256])
def test_gemm_epilogues_every_tile_width(ctx, block_n):
    """Bias / activation / residual epilogues with a half-used last 32-column chunk (block_n % 32 == 16) and a ragged N."""
    g = torch.Generator(device="cuda").manual_seed(19)
    M, N, K = 300, 1096, 320
    a = torch.randn((M, K), device="cuda", generator=g).to(torch.bfloat16)
    b = (torch.randn((N, K), device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    bias = (torch.randn((N,), device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    res = torch.randn((M, N), device="cuda", generator=g).to(torch.bfloat16)
    for epi, kw in [(A.EPI_BIAS, dict(bias=bias)), (A.EPI_BIAS_QUICKGELU, dict(bias=bias)), (A.EPI_BIAS_GELU, dict(bias=bias)),
                    (A.EPI_RESIDUAL, dict(residual=res)), (A.EPI_BIAS_RESIDUAL, dict(bias=bias, residual=res))]:
        assert torch.equal(ctx.gemm(a, b, epilogue=epi, block_n=block_n, **kw), ctx.gemm(a, b, epilogue=epi, block_n=256, **kw)), epi


def test_gemm_strided_views(ctx):
    """A and C may be column slices of wider buffers (leading dimension > width)."""
    M, N, K = 256, 384, 320
    big_a = _ints((M, K + 64), -2, 2, 9)
    a = big_a[:, 32:32 + K]
    if a.data_ptr() % 16:  # TMA needs a 16-byte aligned base
        pytest.skip("unaligned view")
    b = _ints((N, K), -2, 2, 10)
    big_c = torch.zeros((M, N + 128), dtype=torch.bfloat16, device="cuda")
    ctx.gemm(a, b, out=big_c[:, 64:64 + N])
    ref = (a.float() @ b.float().T).to(torch.bfloat16)
    assert torch.equal(big_c[:, 64:64 + N], ref)
    assert big_c[:, :64].abs().sum().item() == 0 and big_c[:, 64 + N:].abs().sum().item() == 0


@pytest.mark.parametrize("M,N,K", [(281, 3584, 18944), (281, 3584, 3584), (281, 4608, 3584), (64, 256, 2048), (384, 136, 4096)])
def test_gemm_splitk_exact_integer_operands(ctx, M, N, K):
    """split-K (fp32 partial tiles + reduce kernel with the fused epilogue): integer operands keep every partial sum
    exact, so the result must be bit-identical to the unsplit reference for all four supported epilogues. By default
    the dispatch splits only K >= 8192 (the first shape); LIVECC_B200_GEMM_SPLITK=1 forces it for all of them (both
    modes were run on B200 in round 2, profiles/r02_parity_7b.md)."""
    a = _ints((M, K), -2, 2, 11)
    b = _ints((N, K), -1, 1, 12)
    bias = _ints((N,), -8, 8, 13)
    res = _ints((M, N), -8, 8, 14)
    ws = torch.empty((8 * M * N,), dtype=torch.float32, device="cuda")
    acc = a.float() @ b.float().T
    assert torch.equal(ctx.gemm(a, b, splitk_ws=ws), acc.to(torch.bfloat16))
    assert torch.equal(ctx.gemm(a, b, bias=bias, epilogue=A.EPI_BIAS, splitk_ws=ws), (acc + bias.float()).to(torch.bfloat16))
    out = res.clone()  # in place: residual aliases the output, as in the decoder
    ctx.gemm(a, b, out=out, residual=out, epilogue=A.EPI_RESIDUAL, splitk_ws=ws)
    assert torch.equal(out, (acc.to(torch.bfloat16).float() + res.float()).to(torch.bfloat16))
    out = ctx.gemm(a, b, bias=bias, residual=res, epilogue=A.EPI_BIAS_RESIDUAL, splitk_ws=ws)
    assert torch.equal(out, ((acc + bias.float()).to(torch.bfloat16).float() + res.float()).to(torch.bfloat16))
