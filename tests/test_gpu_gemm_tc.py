"""K2 tensor-core GEMM (tcgen05 + TMA) against a plain fp32 reference of the same op on bf16-rounded inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from neural_speech_decoder_b200 import ops

DEV = "cuda"


@pytest.mark.parametrize("M,N,K", [
    (128, 256, 64), (128, 64, 128), (256, 128, 256), (200, 300, 192), (7552, 6144, 2048), (1000, 41 * 8, 72),
    (6144, 1024, 7488), (130, 8200, 520), (64, 48, 64),
])
@pytest.mark.parametrize("out_dtype", ["f32", "bf16"])
def test_gemm_bf16_nt(M, N, K, out_dtype):
    if out_dtype == "bf16" and M * N > 4e6:
        pytest.skip("large case covered in f32")
    g = torch.Generator(device="cpu").manual_seed(M * 31 + N * 7 + K)
    A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    B = (torch.randn(N, K, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    odt = torch.float32 if out_dtype == "f32" else torch.bfloat16
    C = torch.empty((M, N), device=DEV, dtype=odt)
    ops.gemm(False, True, M, N, K, A, K, B, K, C, N, bias=bias)
    ref = A.float() @ B.float().T + bias
    tol = dict(rtol=2e-3, atol=2e-3) if out_dtype == "f32" else dict(rtol=1.6e-2, atol=2e-2)
    torch.testing.assert_close(C.float(), ref, **tol)
    # accumulate form: C <- A B^T + 0.5 C (no bias)
    C0 = torch.randn((M, N), device=DEV).to(odt)
    C1 = C0.clone()
    ops.gemm(False, True, M, N, K, A, K, B, K, C1, N, beta=0.5)
    torch.testing.assert_close(C1.float(), A.float() @ B.float().T + 0.5 * C0.float(), **tol)


@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (64, 64, 128), (200, 304, 192), (6144, 2048, 7552), (3072, 1024, 7488),
                                   (7552, 2048, 48), (41 + 7, 2048, 7552), (136, 72, 328)])
def test_gemm_bf16_mn_major_operands(ta, tb, M, N, K):
    """A stored [K,M] (transa) and/or B stored [K,N] (!transb): the natural layouts of wgrad / dgrad operands."""
    g = torch.Generator(device="cpu").manual_seed(M * 31 + N * 7 + K + 2 * ta + tb)
    A = (torch.randn((K, M) if ta else (M, K), generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    B = (torch.randn((N, K) if tb else (K, N), generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    C = torch.empty((M, N), device=DEV)
    ops.gemm(ta, tb, M, N, K, A, A.shape[1], B, B.shape[1], C, N)
    opA = A.float().T if ta else A.float()
    opB = B.float().T if tb else B.float()
    torch.testing.assert_close(C, opA @ opB, rtol=2e-3, atol=2e-3 * max(1.0, (K / 64) ** 0.5))


def test_gemm_bf16_mn_major_row_offsets():
    """W_hh wgrad form: both operands are row ranges of time-major activations (pointer offsets, K = (T'-1)*B rows)."""
    g = torch.Generator(device="cpu").manual_seed(5)
    Bt, Tp, H = 5, 9, 64
    Mrows = Tp * Bt
    dgh = torch.randn(Mrows, 2 * 3 * H, generator=g).to(torch.bfloat16).to(DEV)
    hs = torch.randn(Mrows, 2 * H, generator=g).to(torch.bfloat16).to(DEV)
    Mh = (Tp - 1) * Bt
    out = torch.empty((3 * H, H), device=DEV)
    d = 1
    ops.gemm(True, False, 3 * H, H, Mh, dgh, 2 * 3 * H, hs, 2 * H, out, H, a_off=0 * 2 * 3 * H + d * 3 * H, b_off=Bt * 2 * H + d * H)
    ref = dgh[0:Mh, d * 3 * H:(d + 1) * 3 * H].float().T @ hs[Bt:Bt + Mh, d * H:(d + 1) * H].float()
    torch.testing.assert_close(out, ref, rtol=2e-3, atol=2e-3)


def test_gemm_bf16_strided_views_and_determinism():
    """Sub-matrix operands (leading dimensions larger than K, offset pointers) as the model uses them."""
    g = torch.Generator(device="cpu").manual_seed(0)
    big_a = (torch.randn(300, 512, generator=g)).to(torch.bfloat16).to(DEV)
    big_b = (torch.randn(400, 512, generator=g)).to(torch.bfloat16).to(DEV)
    M, N, K = 300 - 64, 400, 512 - 128
    C = torch.empty((M, 512), device=DEV)
    ops.gemm(False, True, M, N, K, big_a, 512, big_b, 512, C, 512, a_off=64 * 512 + 64, b_off=128)
    ref = big_a[64:, 64:64 + K].float() @ big_b[:, 128:128 + K].float().T
    torch.testing.assert_close(C[:, :N], ref, rtol=2e-3, atol=2e-3)
    C2 = torch.empty_like(C)
    ops.gemm(False, True, M, N, K, big_a, 512, big_b, 512, C2, 512, a_off=64 * 512 + 64, b_off=128)
    assert torch.equal(C[:, :N], C2[:, :N])


def test_cast_transpose():
    g = torch.Generator(device="cpu").manual_seed(1)
    for R, Cn in [(64, 64), (70, 130), (7552, 96), (3, 5)]:
        x = torch.randn(R, Cn, generator=g).to(DEV)
        d, dT = ops.cast_transpose(x, True, True)
        assert torch.equal(d, x.to(torch.bfloat16)) and torch.equal(dT, x.to(torch.bfloat16).T.contiguous())
        xb = x.to(torch.bfloat16)
        _, dT2 = ops.cast_transpose(xb, False, True)
        assert torch.equal(dT2, xb.T.contiguous())
    big = torch.randn(100, 300, generator=g).to(DEV)
    view = big[:, 20:220]                                    # row stride 300, 200 columns
    d, dT = ops.cast_transpose(view, True, True)
    assert torch.equal(d, view.to(torch.bfloat16)) and torch.equal(dT, view.to(torch.bfloat16).T.contiguous())


@pytest.mark.parametrize("M,N,K", [(3072, 1024, 7488), (384, 256, 200), (300, 192, 96), (64, 1024, 128)])
def test_gemm_x2_equals_two_gemms(M, N, K):
    """nsd_gemm_bf16_x2 (two same-shape problems in one launch, the W_hh weight-gradient form: both operands MN-major, sub-matrix
    views of shared buffers) against two nsd_gemm_bf16 calls and a float64 product."""
    torch.manual_seed(M + N)
    lda, ldb = 2 * M + 24, 2 * N + 16
    A = torch.randn(K + 5, lda, device=DEV).to(torch.bfloat16)       # op(A) = A^T: stored [K, M] (M-major)
    Bm = torch.randn(K + 5, ldb, device=DEV).to(torch.bfloat16)      # op(B): stored [K, N] (N-major)
    a_offs, b_offs = [3 * lda, M + 8 + (-M) % 8], [0, 2 * ldb + N + 16]          # 16-byte aligned sub-matrix origins
    C = torch.zeros(2 * M, N, device=DEV)
    ops.gemm_x2(True, False, M, N, K, A, lda, Bm, ldb, C, N, a_offs, b_offs, [0, M * N])
    ref = torch.zeros_like(C)
    for i in range(2):
        ops.gemm(True, False, M, N, K, A, lda, Bm, ldb, ref, N, a_off=a_offs[i], b_off=b_offs[i], c_off=i * M * N)
    assert torch.allclose(C, ref, rtol=1e-6, atol=1e-6)
    Af, Bf = A.double().reshape(-1), Bm.double().reshape(-1)
    for i in range(2):
        a = torch.as_strided(Af, (K, M), (lda, 1), a_offs[i])
        b = torch.as_strided(Bf, (K, N), (ldb, 1), b_offs[i])
        want = a.T @ b
        got = C[i * M:(i + 1) * M].double()
        assert (got - want).abs().max().item() < 2e-3 * max(1.0, want.abs().max().item())
