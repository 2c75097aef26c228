"""Per-kernel parity: every C-ABI entry point of libnsd_b200.so against the numpy oracle
(oracle/nsd_oracle.py, pinned to the reference by tests/test_oracle_golden.py) on the same seeded inputs.
Runs on the B200 box only (-m gpu)."""
import numpy as np
import pytest
import torch

from conftest import golden_ctor, golden_state, load_golden
from oracle import nsd_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from neural_speech_decoder_b200 import ops
    import neural_speech_decoder_b200 as nsd

DEV = "cuda"


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


# ------------------------------------------------------------------ K1
@pytest.mark.parametrize("B,T,N,K,S,nd,sigma", [
    (3, 40, 16, 8, 2, 3, 2.0),
    (4, 61, 16, 14, 4, 4, 1.5),
    (5, 97, 24, 7, 3, 5, 1.0),          # K, N not multiples of 4: scalar paths
    (2, 32, 256, 32, 4, 24, 2.0),       # T == K: a single frame
    (6, 203, 256, 32, 4, 24, 2.0),      # competition channel count, ragged tail (T-K not a multiple of S)
])
def test_frontend_fwd_bwd(B, T, N, K, S, nd, sigma):
    rng = np.random.default_rng(B * 1000 + T)
    x = rng.standard_normal((B, T, N)).astype(np.float32)
    W = (np.eye(N)[None] + 0.05 * rng.standard_normal((nd, N, N))).astype(np.float32)
    bias = (0.1 * rng.standard_normal((nd, 1, N))).astype(np.float32)
    day = rng.integers(0, nd, size=B).astype(np.int64)
    day[-1] = day[0]                                           # a repeated day exercises the segment reduce
    taps = O.gaussian_taps(sigma)
    p_ref, saved = O.frontend_fwd(x.astype(np.float64), taps.astype(np.float64), W.astype(np.float64),
                                  bias.astype(np.float64), day, K, S)
    patches, ys, z = ops.frontend_fwd(cu(x), cu(day), cu(W), cu(bias), cu(taps), K, S, torch.float32)
    Tp = O.n_frames(T, K, S)
    got = patches.view(Tp, B, N * K).permute(1, 0, 2).cpu().numpy()
    np.testing.assert_allclose(got, p_ref, rtol=1e-5, atol=2e-6)
    used = (Tp - 1) * S + K
    np.testing.assert_allclose(ys.cpu().numpy()[:, :used], saved["ys"][:, :used], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(z.cpu().numpy()[:, :used], saved["z"][:, :used], rtol=1e-5, atol=2e-6)
    # unfold is pure data movement: patches must be bit-identical to an unfold of the kernel's own z
    zk = z.cpu().numpy().copy()
    assert np.array_equal(got, O.unfold(zk, K, S))

    dp = rng.standard_normal((B, Tp, N * K)).astype(np.float32)
    dW_ref, db_ref = O.frontend_bwd(dp.astype(np.float64), saved, W.astype(np.float64), day, T, K, S)
    dp_tm = cu(np.ascontiguousarray(dp.transpose(1, 0, 2)).reshape(Tp * B, N * K))
    dW, db = ops.frontend_bwd(dp_tm, ys, z, cu(day), nd, K, S)
    scale = max(1.0, np.abs(dW_ref).max())
    np.testing.assert_allclose(dW.cpu().numpy(), dW_ref, rtol=2e-4, atol=2e-5 * scale)
    np.testing.assert_allclose(db.cpu().numpy(), db_ref, rtol=2e-4, atol=2e-5 * scale)


def test_frontend_bf16_patches_and_bad_day():
    B, T, N, K, S, nd = 3, 64, 32, 16, 4, 4
    rng = np.random.default_rng(5)
    x = rng.standard_normal((B, T, N)).astype(np.float32)
    W = np.tile(np.eye(N, dtype=np.float32)[None], (nd, 1, 1))
    bias = np.zeros((nd, 1, N), np.float32)
    taps = O.gaussian_taps(2.0)
    day = np.array([0, 1, 2], dtype=np.int64)
    p32, _, _ = ops.frontend_fwd(cu(x), cu(day), cu(W), cu(bias), cu(taps), K, S, torch.float32)
    p16, _, _ = ops.frontend_fwd(cu(x), cu(day), cu(W), cu(bias), cu(taps), K, S, torch.bfloat16)
    assert torch.equal(p16, p32.to(torch.bfloat16))           # same values, rounded to nearest even once
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.frontend_fwd(cu(x), cu(np.array([0, 9, 1], dtype=np.int64)), cu(W), cu(bias), cu(taps), K, S, torch.float32, flag)
    assert int(flag.item()) == 1                               # reference: IndexError from index_select (model.py:89)
    with pytest.raises(RuntimeError):
        ops.frontend_fwd(cu(x[:, :8]), cu(day), cu(W), cu(bias), cu(taps), K, S, torch.float32)   # T < kernelLen


# ------------------------------------------------------------------ K2 (fp32)
@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (37, 41, 19), (128, 128, 16), (300, 123, 260), (513, 96, 1024)])
def test_gemm_f32(ta, tb, M, N, K):
    rng = np.random.default_rng(M + 7 * N + 13 * K)
    A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    Bm = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    opA = A.T if ta else A
    opB = Bm.T if tb else Bm
    ref = opA.astype(np.float64) @ opB.astype(np.float64) + bias + 0.5 * C0
    Cd = cu(C0)
    ops.gemm(ta, tb, M, N, K, cu(A), A.shape[1], cu(Bm), Bm.shape[1], Cd, N, bias=cu(bias), beta=0.5)
    np.testing.assert_allclose(Cd.cpu().numpy(), ref, rtol=1e-4, atol=1e-4 * np.sqrt(K))


def test_colsum_cast_swap():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((333, 70)).astype(np.float32)
    out = torch.empty(50, device=DEV)
    ops.colsum(cu(a), 333, 50, 70, out, a_off=3)
    np.testing.assert_allclose(out.cpu().numpy(), a[:, 3:53].astype(np.float64).sum(0), rtol=1e-5, atol=1e-4)
    t = cu(a)
    assert torch.equal(ops.cast(t, torch.bfloat16), t.to(torch.bfloat16))
    assert torch.equal(ops.cast(t.to(torch.bfloat16), torch.float32), t.to(torch.bfloat16).float())
    x = cu(rng.standard_normal((5, 7, 11)).astype(np.float32))
    assert torch.equal(ops.swap01(x), x.permute(1, 0, 2).contiguous())


# ------------------------------------------------------------------ K3 (fp32)
@pytest.mark.parametrize("B,Tp,H,reverse", [(3, 7, 32, False), (3, 7, 32, True), (5, 1, 20, False), (70, 9, 24, True),
                                            (4, 6, 1024, False)])
def test_gru_f32_fwd_bwd(B, Tp, H, reverse):
    rng = np.random.default_rng(H + Tp)
    In = 12
    x = rng.standard_normal((B, Tp, In))
    w_ih = rng.standard_normal((3 * H, In)) * 0.3
    w_hh = rng.standard_normal((3 * H, H)) / np.sqrt(H)
    b_ih = rng.standard_normal(3 * H) * 0.1
    b_hh = rng.standard_normal(3 * H) * 0.1
    hseq_ref, saved = O.gru_dir_fwd(x, w_ih, w_hh, b_ih, b_hh, reverse)
    gi = (x @ w_ih.T + b_ih)                                           # [B,Tp,3H]
    gi_tm = cu(np.ascontiguousarray(gi.transpose(1, 0, 2)).reshape(Tp * B, 3 * H), torch.float32)
    hseq = torch.empty((Tp * B, H), device=DEV)
    sv = tuple(torch.empty((Tp * B, H), device=DEV) for _ in range(4))
    ops.gru_fwd_f32(gi_tm, 3 * H, 0, cu(w_hh, torch.float32), cu(b_hh, torch.float32), Tp, B, H, reverse, hseq, H, 0, sv)
    got = hseq.view(Tp, B, H).permute(1, 0, 2).cpu().numpy()
    np.testing.assert_allclose(got, hseq_ref, rtol=1e-4, atol=2e-5)

    dh = rng.standard_normal((B, Tp, H))
    dx_ref, dwi_ref, dwh_ref, dbi_ref, dbh_ref = O.gru_dir_bwd(dh, saved, w_ih, w_hh, reverse)
    dh_tm = cu(np.ascontiguousarray(dh.transpose(1, 0, 2)).reshape(Tp * B, H), torch.float32)
    dgi = torch.empty((Tp * B, 3 * H), device=DEV)
    dghn = torch.empty((Tp * B, H), device=DEV)
    ops.gru_bwd_f32(dh_tm, H, 0, hseq, H, 0, sv, cu(w_hh, torch.float32), Tp, B, H, reverse, dgi, 3 * H, 0, dghn)
    # d b_ih = column sums of dgi ; dx = dgi W_ih
    dgi_np = dgi.cpu().numpy().astype(np.float64)
    tol = dict(rtol=2e-3, atol=2e-4 * max(1.0, np.abs(dbi_ref).max()))
    np.testing.assert_allclose(dgi_np.sum(0), dbi_ref, **tol)
    dx = (dgi_np @ w_ih).reshape(Tp, B, In).transpose(1, 0, 2)
    np.testing.assert_allclose(dx, dx_ref, rtol=2e-3, atol=2e-4 * max(1.0, np.abs(dx_ref).max()))
    dgh = np.concatenate([dgi_np[:, :2 * H], dghn.cpu().numpy().astype(np.float64)], axis=1)
    np.testing.assert_allclose(dgh.sum(0), dbh_ref, **tol)
    hprev = saved["hprev"].transpose(1, 0, 2).reshape(Tp * B, H)
    np.testing.assert_allclose(dgh.T @ hprev, dwh_ref, rtol=2e-3, atol=2e-4 * max(1.0, np.abs(dwh_ref).max()))


def test_dropout_statistics_and_backward_mask():
    x = torch.ones(1 << 20, device=DEV)
    y = ops.dropout(x, 0.4, seed=123)
    keep = (y != 0).float().mean().item()
    assert abs(keep - 0.6) < 5e-3
    assert torch.allclose(y[y != 0], torch.tensor(1 / 0.6, device=DEV))
    assert torch.equal(ops.dropout(x, 0.4, seed=123), y)               # same (seed, index) -> same mask
    assert not torch.equal(ops.dropout(x, 0.4, seed=124), y)
    xb = torch.ones(4099, device=DEV, dtype=torch.bfloat16)
    yb = ops.dropout(xb, 0.4, seed=123)
    assert torch.equal(yb != 0, y[:4099] != 0)                         # mask independent of dtype


# ------------------------------------------------------------------ K4 / K5
def _ctc_case(T, B, C, max_tgt, seed, ragged=True):
    rng = np.random.default_rng(seed)
    logits = rng.standard_normal((B, T, C)).astype(np.float32) * 2
    il = rng.integers(max(1, T // 2), T + 1, size=B).astype(np.int32) if ragged else np.full(B, T, np.int32)
    yl = np.minimum(rng.integers(0, max_tgt + 1, size=B), il // 2).astype(np.int32)
    y = np.zeros((B, max(1, max_tgt)), np.int32)
    for b in range(B):
        y[b, :yl[b]] = rng.integers(1, C, size=yl[b])
    return logits, y, il, yl


@pytest.mark.parametrize("T,B,C,max_tgt,seed", [(12, 5, 6, 5, 0), (117, 64, 41, 50, 1), (30, 7, 41, 20, 2), (492, 16, 41, 120, 3)])
@pytest.mark.parametrize("reduction", ["mean", "sum"])
def test_ctc_matches_oracle(T, B, C, max_tgt, seed, reduction):
    logits, y, il, yl = _ctc_case(T, B, C, max_tgt, seed)
    if seed == 0:
        y[0, :3] = [1, 1, 2]; yl[0] = 3; il[0] = 12          # repeated labels
        yl[1] = 0                                            # empty target
        y[2, :5] = 2; yl[2] = 5; il[2] = 8                   # infeasible: needs 9 frames (zero_infinity)
    lp = O.log_softmax(logits.astype(np.float64), axis=2)
    loss_ref, nll_ref, g_ref = O.ctc_loss(np.transpose(lp, (1, 0, 2)), y, il, yl, reduction=reduction)
    # (a) drop-in module on a permuted log-prob view, exactly as the trainer calls it
    lg = cu(logits).requires_grad_(True)
    lp_t = lg.log_softmax(2).permute(1, 0, 2)
    lp_t.retain_grad()
    loss = nsd.CTCLoss(blank=0, reduction=reduction, zero_infinity=True)(lp_t, cu(y), cu(il), cu(yl))
    loss.backward()
    np.testing.assert_allclose(loss.item(), loss_ref, rtol=2e-5, atol=1e-5)
    # fp32 log-space alpha/beta (as torch's CUDA kernel): |alpha| grows ~3 per frame, so the occupancies carry
    # ~1e-3 relative error per 500 frames; north-star tolerance rtol 1e-3 holds at the benchmark's 117 frames
    rtol = 1e-3 if T <= 128 else 1e-2
    atol = (3e-6 if T <= 128 else 1e-4) * (1 if reduction == "mean" else B * 20)
    np.testing.assert_allclose(lp_t.grad.cpu().numpy(), g_ref, rtol=rtol, atol=atol)
    # (b) fused logits entry point: same loss; d/dlogits == log-softmax backward of (a)'s gradient
    lg2 = cu(logits).requires_grad_(True)
    loss2 = nsd.ctc_loss_from_logits(lg2, cu(y), cu(il), cu(yl), reduction=reduction)
    loss2.backward()
    np.testing.assert_allclose(loss2.item(), loss_ref, rtol=2e-5, atol=1e-5)
    dl_ref = g_ref.transpose(1, 0, 2) - np.exp(lp) * g_ref.transpose(1, 0, 2).sum(2, keepdims=True)
    np.testing.assert_allclose(lg2.grad.cpu().numpy(), dl_ref, rtol=rtol, atol=atol)
    np.testing.assert_allclose(lg.grad.cpu().numpy(), dl_ref, rtol=rtol, atol=atol)
    if seed == 0:
        assert np.all(lg2.grad.cpu().numpy()[2] == 0)         # infeasible utterance: zero loss, zero gradient
    # deterministic: bit-identical on a second run
    lg3 = cu(logits).requires_grad_(True)
    loss3 = nsd.ctc_loss_from_logits(lg3, cu(y), cu(il), cu(yl), reduction=reduction)
    loss3.backward()
    assert torch.equal(loss3, loss2) and torch.equal(lg3.grad, lg2.grad)


def test_ctc_none_reduction_and_golden():
    for name in ("small_uni", "small_bi", "comp_uni", "comp_bi"):
        g = load_golden(name + "_f32")
        lp = cu(g["log_probs_tbc"]).requires_grad_(True)
        nll = nsd.CTCLoss(blank=0, reduction="none", zero_infinity=True)(lp, cu(g["y"]), cu(g["out_lens"]), cu(g["y_len"]))
        np.testing.assert_allclose(nll.detach().cpu().numpy(), g["nll"], rtol=2e-5, atol=1e-5)
        lp2 = cu(g["log_probs_tbc"]).requires_grad_(True)
        loss = nsd.CTCLoss(blank=0, reduction="mean", zero_infinity=True)(lp2, cu(g["y"]), cu(g["out_lens"]), cu(g["y_len"]))
        loss.backward()
        np.testing.assert_allclose(loss.item(), g["loss"], rtol=2e-5)
        np.testing.assert_allclose(lp2.grad.cpu().numpy(), g["dlog_probs_tbc"], rtol=1e-3, atol=3e-7)


@pytest.mark.parametrize("T,B,C,seed", [(12, 5, 6, 0), (117, 64, 41, 1), (493, 33, 41, 2), (1, 3, 4, 3)])
def test_greedy_decode_and_edit_distance_bit_exact(T, B, C, seed):
    rng = np.random.default_rng(seed)
    logits = rng.standard_normal((B, T, C)).astype(np.float32)
    logits[:, :, 0] += 1.0                                           # plenty of blanks and repeats
    logits = np.round(logits * 2) / 2                                # exact ties: lowest index must win
    lp = O.log_softmax(logits.astype(np.float32), axis=2).astype(np.float32)
    lens = rng.integers(0, T + 1, size=B).astype(np.int32)
    lens[0] = T
    lp_tbc = cu(lp).permute(1, 0, 2)                                 # non-contiguous view, as in the trainer
    dec, dec_len = nsd.greedy_decode(lp_tbc, cu(lens))
    got = nsd.decoded_to_lists(dec, dec_len)
    ref = O.greedy_decode(np.transpose(lp, (1, 0, 2)), lens)
    assert got == ref
    yl = rng.integers(0, 30, size=B).astype(np.int32)
    y = np.zeros((B, 30), np.int32)
    for b in range(B):
        y[b, :yl[b]] = rng.integers(1, C, size=yl[b])
    dist = nsd.edit_distances(dec, dec_len, cu(y), cu(yl)).cpu().numpy()
    ref_d = [O.edit_distance(y[b, :yl[b]].tolist(), ref[b]) for b in range(B)]
    assert dist.tolist() == ref_d
    d, tot = nsd.phoneme_error_rate(lp_tbc, cu(lens), cu(y), cu(yl))
    assert (d, tot) == O.phoneme_error_rate(ref, y, yl)


def test_log_softmax_kernel():
    rng = np.random.default_rng(4)
    x = (rng.standard_normal((7, 19, 41)) * 5).astype(np.float32)
    got = nsd.ctc.log_softmax_tbc(cu(x))
    assert got.shape == (19, 7, 41) and not got.is_contiguous()
    np.testing.assert_allclose(got.permute(1, 0, 2).cpu().numpy(), O.log_softmax(x.astype(np.float64), 2), rtol=1e-5, atol=2e-6)


# ------------------------------------------------------------------ Adam
def test_fused_adam_matches_torch_adam():
    torch.manual_seed(0)
    shapes = [(300, 77), (5,), (1024, 1024), (3, 1, 9)]
    ps = [torch.randn(s, device=DEV) for s in shapes]
    ref = [p.clone().requires_grad_(True) for p in ps]
    mine = [p.clone().requires_grad_(True) for p in ps]
    o_ref = torch.optim.Adam(ref, lr=0.02, betas=(0.9, 0.999), eps=0.1, weight_decay=1e-5)
    o_mine = nsd.adam.FusedAdam(mine + [torch.nn.Parameter(torch.zeros(3, device=DEV))], lr=0.02, betas=(0.9, 0.999),
                                eps=0.1, weight_decay=1e-5)
    for it in range(5):
        for a, b in zip(ref, mine):
            g = torch.randn_like(a)
            a.grad = g.clone(); b.grad = g.clone()
        o_ref.step(); o_mine.step()
    for a, b in zip(ref, mine):
        torch.testing.assert_close(b, a, rtol=1e-5, atol=1e-6)


def test_transpose_bf16_multi_is_exact():
    """One launch transposes every layer's / direction's W_hh (row-range views of stacked buffers) bit for bit."""
    torch.manual_seed(3)
    for R, Cn, n in [(192, 64, 4), (3072, 1024, 10), (70, 130, 3)]:
        src = torch.randn(n, R + 6, Cn, device=DEV).to(torch.bfloat16)
        dst = torch.zeros(n, Cn, R + 2, device=DEV, dtype=torch.bfloat16)
        ops.transpose_bf16_multi([src[i, 2:2 + R] for i in range(n)], [dst[i, :, :R] for i in range(n)])
        for i in range(n):
            assert torch.equal(dst[i, :, :R], src[i, 2:2 + R].t())
        assert float(dst[:, :, R:].abs().max()) == 0.0


def test_fused_adam_split_step_is_bit_identical():
    """step(only=...) in two calls (what the data-parallel train_step does under the last all-reduce) == one step()."""
    torch.manual_seed(1)
    shapes = [(300, 77), (5,), (512, 640), (3, 1, 9)]
    ps = [torch.randn(s, device=DEV) for s in shapes]
    one = [p.clone().requires_grad_(True) for p in ps]
    two = [p.clone().requires_grad_(True) for p in ps]
    kw = dict(lr=0.02, betas=(0.9, 0.999), eps=0.1, weight_decay=1e-5, grad_scale=0.5)
    o_one, o_two = nsd.adam.FusedAdam(one, **kw), nsd.adam.FusedAdam(two, **kw)
    for it in range(3):
        for a, b in zip(one, two):
            g = torch.randn_like(a)
            a.grad = g.clone(); b.grad = g.clone()
        o_one.step()
        o_two.step(only=two[:2]); o_two.step(only=two[2:])
    for a, b in zip(one, two):
        assert torch.equal(a, b)
        assert o_two.state[b]["step"] == 3


def test_input_noise_moments_and_fused_front_end():
    """trainer:194-201: X += randn*whiteNoiseSD; X += randn([B,1,N])*constantOffsetSD.  Distributional parity (own
    counter-based generator): moments of the two components; the copy fused into K1 must equal K1 of the explicitly
    noised input bit for bit."""
    import neural_speech_decoder_b200 as nsd
    B, T, N = 8, 300, 256
    x = torch.zeros(B, T, N, device=DEV)
    w = ops.input_noise(x, 0.8, 0.0, 11)
    assert abs(w.mean().item()) < 5e-3 and abs(w.std().item() - 0.8) < 5e-3
    kurt = ((w / w.std()) ** 4).mean().item()
    assert abs(kurt - 3.0) < 0.05                                       # Gaussian, not uniform
    assert abs(torch.corrcoef(torch.stack([w[:, :-1].reshape(-1), w[:, 1:].reshape(-1)]))[0, 1].item()) < 5e-3
    o = ops.input_noise(x, 0.0, 0.2, 11)
    assert torch.equal(o, o[:, :1].expand(-1, T, -1).contiguous())      # constant over time
    off = o[:, 0]
    assert abs(off.mean().item()) < 2e-2 and abs(off.std().item() - 0.2) < 2e-2
    assert not torch.equal(ops.input_noise(x, 0.8, 0.2, 11), ops.input_noise(x, 0.8, 0.2, 12))
    both = ops.input_noise(x, 0.8, 0.2, 11)
    assert torch.allclose(both, w + o, atol=1e-6)
    # fused into the front end
    torch.manual_seed(0)
    xr = torch.randn(B, T, N, device=DEV)
    day = torch.randint(0, 4, (B,), device=DEV)
    dw = (torch.eye(N, device=DEV) + 0.05 * torch.randn(4, N, N, device=DEV)).contiguous()
    db = 0.1 * torch.randn(4, 1, N, device=DEV)
    taps = nsd.model.gaussian_kernel_1d(20, 2.0).to(DEV)
    for dt in (torch.float32, torch.bfloat16):
        fused = ops.frontend_fwd(xr, day, dw, db, taps, 32, 4, dt, None, (0.8, 0.2, 77))
        plain = ops.frontend_fwd(ops.input_noise(xr, 0.8, 0.2, 77), day, dw, db, taps, 32, 4, dt)
        for a, b in zip(fused, plain):
            assert torch.equal(a, b)
    # odd channel count takes the scalar path
    xo = torch.randn(3, 50, 30, device=DEV)
    no = ops.input_noise(torch.zeros_like(xo), 1.0, 0.0, 5)
    assert abs(no.std().item() - 1.0) < 0.05


def test_train_step_with_fused_noise_runs():
    import neural_speech_decoder_b200 as nsd
    from neural_speech_decoder_b200.synthetic import make_batch
    kw = dict(neural_dim=32, n_classes=10, hidden_dim=64, layer_dim=2, nDays=4, dropout=0.2, strideLen=4, kernelLen=16,
              gaussianSmoothWidth=2.0, bidirectional=True)
    torch.manual_seed(0)
    m = nsd.GRUDecoder(device=DEV, **kw).to(DEV).train()
    batch = [t.to(DEV) for t in make_batch(4, 72, n_feat=32, n_days=4, n_classes=10, seed=5, min_tgt=2, max_tgt=6, kernel_len=16, stride_len=4)]
    opt, sched = nsd.make_optimizer(m, dict(lrStart=0.02, lrEnd=0.02, nBatch=100, l2_decay=1e-5))
    a = nsd.train_step(m, opt, *batch, white_noise_sd=0.8, constant_offset_sd=0.2).item()
    assert np.isfinite(a)
    m.eval()                                                            # eval: no augmentation even if configured
    X, day = batch[0], batch[4]
    assert torch.equal(m.forward(X, day), m.forward(X, day))


@pytest.mark.parametrize("B,T,N,K,S,nd", [(5, 203, 256, 32, 4, 24), (3, 100, 128, 16, 4, 4), (2, 64, 64, 32, 4, 3), (4, 75, 256, 14, 4, 5)])
def test_frontend_tensor_core_path(B, T, N, K, S, nd):
    """bf16 model path of K1: the day affine (forward) and ys^T dpre (backward) run on mma.sync with bf16 operands and
    fp32 accumulation.  Tolerances are bf16's (stated separately from the fp32 path above): z within 1e-2 absolute of the
    fp64 oracle (|z| < 1), patches == bf16(z) exactly (the unfold stays pure data movement), dW/db within 2 % of max|dW|."""
    rng = np.random.default_rng(N + T)
    x = rng.standard_normal((B, T, N)).astype(np.float32)
    W = (np.eye(N)[None] + 0.05 * rng.standard_normal((nd, N, N))).astype(np.float32)
    bias = (0.1 * rng.standard_normal((nd, 1, N))).astype(np.float32)
    day = rng.integers(0, nd, size=B).astype(np.int64)
    day[-1] = day[0]
    taps = O.gaussian_taps(2.0)
    p_ref, saved = O.frontend_fwd(x.astype(np.float64), taps.astype(np.float64), W.astype(np.float64), bias.astype(np.float64), day, K, S)
    patches, ys, z = ops.frontend_fwd(cu(x), cu(day), cu(W), cu(bias), cu(taps), K, S, torch.bfloat16)
    Tp = O.n_frames(T, K, S)
    used = (Tp - 1) * S + K
    np.testing.assert_allclose(ys.cpu().numpy()[:, :used], saved["ys"][:, :used], rtol=1e-5, atol=2e-6)     # smoothing stays fp32
    zerr = np.abs(z.cpu().numpy()[:, :used] - saved["z"][:, :used]).max()
    print(f"tc front end N={N}: max |z - oracle| = {zerr:.2e}")
    assert zerr < 1e-2
    zk = torch.from_numpy(O.unfold(z.cpu().numpy().copy(), K, S)).to(torch.bfloat16)
    assert torch.equal(patches.view(Tp, B, N * K).permute(1, 0, 2).cpu(), zk)

    dp = rng.standard_normal((B, Tp, N * K)).astype(np.float32)
    dp_bf = torch.from_numpy(np.ascontiguousarray(dp.transpose(1, 0, 2)).reshape(Tp * B, N * K)).to(torch.bfloat16).to(DEV)
    dp_r = dp_bf.float().cpu().numpy().reshape(Tp, B, N * K).transpose(1, 0, 2).astype(np.float64)
    saved_k = dict(saved)
    zk64 = z.cpu().numpy().astype(np.float64)
    saved_k["ys"], saved_k["z"] = ys.cpu().numpy().astype(np.float64), zk64
    saved_k["pre"] = zk64 / (1.0 - np.abs(zk64))               # the backward is checked against the kernel's own forward state
    dW_ref, db_ref = O.frontend_bwd(dp_r, saved_k, W.astype(np.float64), day, T, K, S)
    dW, db = ops.frontend_bwd(dp_bf, ys, z, cu(day), nd, K, S)
    scale = np.abs(dW_ref).max()
    e_w = np.abs(dW.cpu().numpy() - dW_ref).max() / scale
    e_b = np.abs(db.cpu().numpy() - db_ref).max() / max(np.abs(db_ref).max(), 1e-9)
    print(f"tc front end N={N}: dW err / max|dW| = {e_w:.2e}, db err / max|db| = {e_b:.2e}")
    assert e_w < 2e-2 and e_b < 2e-3                            # db is accumulated in fp32 from fp32 dpre
    dW2, db2 = ops.frontend_bwd(dp_bf, ys, z, cu(day), nd, K, S)
    assert torch.equal(dW, dW2) and torch.equal(db, db2)        # deterministic
