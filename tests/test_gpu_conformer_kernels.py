"""Kernel-level parity of the Conformer path's CUDA kernels (csrc/conformer_ew.cu, csrc/conformer_attn.cu) through the C ABI against
fp64 torch operators on the host (the operators the reference's modules dispatch to: F.layer_norm, F.silu / gelu / relu, F.glu,
F.conv1d(groups=D), softmax with a key-padding mask, matmul), at odd sizes: widths that are not multiples of 128, lengths that are
not multiples of the time tile, every tap count, strided / transposed / indexed batched-GEMM operands."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import neural_speech_decoder_b200 as nsd
from neural_speech_decoder_b200 import conformer as CF

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0) if torch.cuda.is_available() else None


def _close(got, ref, tol, what=""):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    err = (got - ref).abs().max().item() / max(1.0, ref.abs().max().item())
    assert err < tol, f"{what}: max err / max|ref| = {err:.3e} >= {tol}"


@pytest.mark.parametrize("M,D,act", [(37, 64, 0), (130, 1024, 1), (9, 100, 2), (70, 2048, 0), (5, 520, 1)])
def test_layernorm_fwd_bwd(M, D, act):
    g = torch.Generator().manual_seed(M + D)
    x = torch.randn(M, D, generator=g, dtype=torch.float64) * 2 + 0.5
    gamma, beta = torch.randn(D, generator=g, dtype=torch.float64), torch.randn(D, generator=g, dtype=torch.float64)
    dy = torch.randn(M, D, generator=g, dtype=torch.float64)
    xr, gr, br = (t.clone().requires_grad_(True) for t in (x, gamma, beta))
    y = F.layer_norm(xr, (D,), gr, br, 1e-5)
    y = [y, F.silu(y), F.gelu(y)][act]
    y.backward(dy)
    xc, gc, bc = (t.float().to(DEV).requires_grad_(True) for t in (x, gamma, beta))
    out = CF._LayerNorm.apply(xc, gc, bc, 1e-5, act, 0.0, 0, True, True)
    _close(out[0], y, 2e-6, "y f32")
    _close(out[1], y, 6e-3, "y bf16")
    out[0].backward(dy.float().to(DEV))
    _close(xc.grad, xr.grad, 2e-5, "dx"); _close(gc.grad, gr.grad, 2e-5, "dgamma"); _close(bc.grad, br.grad, 2e-5, "dbeta")


@pytest.mark.parametrize("act", [1, 2, 3])
def test_activation_and_glu(act):
    g = torch.Generator().manual_seed(act)
    x = torch.randn(33, 72, generator=g, dtype=torch.float64) * 3
    dy = torch.randn(33, 72, generator=g, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    y = {1: F.silu, 2: F.gelu, 3: F.relu}[act](xr)
    y.backward(dy)
    xc = x.float().to(DEV).requires_grad_(True)
    yc = CF._Act.apply(xc, act, 0.0, 0, torch.float32)
    yc.backward(dy.float().to(DEV))
    _close(yc, y, 2e-6, "act"); _close(xc.grad, xr.grad, 2e-6, "act grad")
    u = torch.randn(21, 2 * 40, generator=g, dtype=torch.float64)
    dg = torch.randn(21, 40, generator=g, dtype=torch.float64)
    ur = u.clone().requires_grad_(True)
    F.glu(ur, dim=-1).backward(dg)
    uc = u.float().to(DEV).requires_grad_(True)
    gc = CF._GLU.apply(uc)
    gc.backward(dg.float().to(DEV))
    _close(gc, F.glu(u, dim=-1), 2e-6, "glu"); _close(uc.grad, ur.grad, 2e-6, "glu grad")


@pytest.mark.parametrize("B,T,D,k", [(3, 37, 40, 31), (2, 16, 128, 7), (4, 5, 33, 5), (1, 118, 1024, 31), (2, 50, 260, 1)])
def test_depthwise_conv_fwd_bwd(B, T, D, k):
    g = torch.Generator().manual_seed(T + k)
    x = torch.randn(B, T, D, generator=g, dtype=torch.float64)
    w = torch.randn(D, 1, k, generator=g, dtype=torch.float64)
    b = torch.randn(D, generator=g, dtype=torch.float64)
    dy = torch.randn(B, T, D, generator=g, dtype=torch.float64)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    y = F.conv1d(xr.transpose(1, 2), wr, br, padding=k // 2, groups=D).transpose(1, 2)
    y.backward(dy)
    xc, wc, bc = (t.float().to(DEV).requires_grad_(True) for t in (x.reshape(B * T, D), w, b))
    yc = CF._DwConv.apply(xc, wc, bc, B, T)
    yc.backward(dy.reshape(B * T, D).float().to(DEV))
    _close(yc.view(B, T, D), y, 3e-6, "dwconv"); _close(xc.grad.view(B, T, D), xr.grad, 3e-6, "dx")
    _close(wc.grad, wr.grad, 1e-5, "dw"); _close(bc.grad, br.grad, 1e-5, "db")


@pytest.mark.parametrize("B,T,N,K,S,sigma", [(3, 84, 32, 16, 4, 2.0), (2, 50, 64, 8, 2, 1.0), (2, 500, 256, 32, 4, 2.0), (2, 40, 32, 32, 4, 0.0)])
def test_frontend_day_affine_smoothing_strided_conv(B, T, N, K, S, sigma):
    g = torch.Generator().manual_seed(T + N)
    n_days = 5
    X = torch.randn(B, T, N, generator=g, dtype=torch.float64)
    day = torch.randint(0, n_days, (B,), generator=g)
    day[-1] = day[0]                                               # a repeated day: the index_select backward must add
    W = torch.randn(n_days, N, N, generator=g, dtype=torch.float64) / math.sqrt(N)
    bias = torch.randn(n_days, 1, N, generator=g, dtype=torch.float64)
    tw = torch.randn(N, 1, K, generator=g, dtype=torch.float64) / K
    gauss = None
    if sigma > 0:
        ks = int(sigma * 4) + 1
        t = torch.arange(ks, dtype=torch.float64) - (ks - 1) / 2
        gauss = torch.exp(-t.pow(2) / (2 * sigma ** 2)); gauss = gauss / gauss.sum()
    Wr, br, twr = (t.clone().requires_grad_(True) for t in (W, bias, tw))
    xa = torch.einsum("btd,bdk->btk", X, Wr[day]) + br[day]
    if gauss is not None:
        xa = F.conv1d(xa.transpose(1, 2), gauss.view(1, 1, -1).repeat(N, 1, 1), padding=gauss.numel() // 2, groups=N).transpose(1, 2)
    y = F.conv1d(xa.transpose(1, 2), twr, None, stride=S, groups=N).transpose(1, 2)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    Wc, bc, twc = (t.float().to(DEV).requires_grad_(True) for t in (W, bias, tw))
    yc = CF._Frontend.apply(X.float().to(DEV), day.to(DEV), Wc, bc, Wc.detach(), gauss.float().to(DEV) if gauss is not None else None, twc, S, False)
    yc.backward(dy.reshape(-1, N).float().to(DEV))
    _close(yc.view(y.shape), y, 5e-6, "feats"); _close(Wc.grad, Wr.grad, 2e-5, "d day_weights")
    _close(bc.grad, br.grad, 2e-5, "d day_bias"); _close(twc.grad, twr.grad, 2e-5, "d temporal_conv.weight")


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 5e-6), (torch.bfloat16, 1.5e-2)])
@pytest.mark.parametrize("B,T,H,dh", [(3, 37, 2, 16), (2, 118, 8, 128), (2, 9, 4, 8), (1, 130, 2, 64)])
def test_attention_fwd_bwd_with_key_padding(B, T, H, dh, dtype, tol):
    g = torch.Generator().manual_seed(T + H)
    D = H * dh
    qkv = torch.randn(B * T, 3 * D, generator=g, dtype=torch.float64)
    lens = torch.randint(max(1, T // 2), T + 1, (B,), generator=g).to(torch.int32); lens[0] = T
    do = torch.randn(B * T, D, generator=g, dtype=torch.float64)
    if dtype == torch.bfloat16:                                   # compare on the values the kernel really receives
        qkv, do = qkv.to(dtype).double(), do.to(dtype).double()
    qr = qkv.clone().requires_grad_(True)
    q, k, v = (t.reshape(B, T, H, dh).transpose(1, 2) for t in qr.view(B, T, 3 * D).split(D, dim=-1))
    s = (q / math.sqrt(dh)) @ k.transpose(-1, -2)
    s = s.masked_fill((torch.arange(T)[None, :] >= lens[:, None])[:, None, None, :], float("-inf"))
    o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, D)
    o.backward(do)
    qc = qkv.to(dtype).to(DEV).requires_grad_(True)
    oc = CF._Attention.apply(qc, lens.to(DEV), B, T, H, 0.0, 0)
    oc.backward(do.to(dtype).to(DEV))
    _close(oc, o, tol, "attention out"); _close(qc.grad, qr.grad, 2 * tol, "d qkv")


def test_bgemm_strides_index_bias():
    """nsd_bgemm on every operand layout the model uses and some it does not: row- / column-contiguous A and B, f32 and bf16, a batch
    index on B with an indexed bias, sizes that end inside a tile, 16-byte-misaligned rows (element-wise path)."""
    g = torch.Generator().manual_seed(0)
    for (M, N, K, nb0, nb1, ta, tb, dt, tol) in [(70, 45, 37, 3, 2, False, False, torch.float32, 3e-6), (64, 64, 64, 2, 1, True, True, torch.float32, 3e-6),
                                                 (118, 118, 128, 2, 3, False, True, torch.bfloat16, 8e-3), (118, 128, 118, 2, 2, True, False, torch.bfloat16, 8e-3),
                                                 (33, 130, 72, 4, 1, False, False, torch.bfloat16, 8e-3), (5, 7, 250, 2, 2, True, True, torch.bfloat16, 8e-3)]:
        A = torch.randn(nb0, nb1, *((K, M) if ta else (M, K)), generator=g).to(dt)
        Bm = torch.randn(nb0 + 1, nb1, *((N, K) if tb else (K, N)), generator=g).to(dt)
        idx = torch.randint(0, nb0 + 1, (nb0,), generator=g)
        bias = torch.randn(nb0 + 1, N, generator=g)
        Ad, Bd = A.double(), Bm.double()
        ref = 0.7 * ((Ad.transpose(-1, -2) if ta else Ad) @ (Bd.transpose(-1, -2) if tb else Bd)[idx]) + bias.double()[idx][:, None, None, :]
        Ac, Bc, C = A.to(DEV), Bm.to(DEV), torch.empty(nb0, nb1, M, N, device=DEV)
        a_str = ((1, M) if ta else (K, 1)) + (nb1 * M * K, M * K)
        b_str = ((1, K) if tb else (N, 1)) + (nb1 * K * N, K * N)
        CF._bgemm(Ac, a_str, Bc, b_str, C, (N, nb1 * M * N, M * N), M, N, K, nb0, nb1, alpha=0.7, b_index=idx.to(DEV), bias=bias.to(DEV), bias_b0=N)
        _close(C, ref, tol, f"bgemm M={M} N={N} K={K} ta={ta} tb={tb} {dt}")


def test_softmax_mask_dropout_consistency():
    """Attention-weight dropout: forward and backward regenerate the same mask; kept weights are scaled by 1/(1-p)."""
    B, H, T, p = 2, 3, 40, 0.3
    Tp = 40
    S = torch.randn(B * H * T, Tp, device=DEV)
    lens = torch.tensor([40, 25], dtype=torch.int32, device=DEV)
    P = S.clone()
    Pd = torch.empty_like(P)
    from neural_speech_decoder_b200._lib import call, ptr, stream, F32
    call("nsd_softmax_mask_fwd", ptr(P), ptr(Pd), F32, ptr(lens), B, H, T, Tp, p, 77, stream())
    ref = torch.softmax(S.view(B, H, T, Tp).masked_fill((torch.arange(Tp, device=DEV)[None, :] >= lens[:, None])[:, None, None, :], float("-inf")), -1).view_as(P)
    assert torch.allclose(P, ref, atol=1e-6)
    kept = Pd != 0
    assert torch.allclose(Pd[kept], P[kept] / (1 - p), rtol=1e-6) and abs(kept[ref > 0].float().mean().item() - (1 - p)) < 0.02
    dPd = torch.ones_like(P)
    call("nsd_softmax_mask_bwd", ptr(P), ptr(dPd), B, H, T, Tp, p, 77, stream())
    m = kept.float() / (1 - p)
    ref_dS = P * (m - (m * P).sum(-1, keepdim=True))
    assert torch.allclose(dPd, ref_dS, atol=1e-5)
