"""End-to-end parity of the drop-in GRUDecoder + CTC against fixtures generated from the reference itself
(tests/golden/*.npz) and against the numpy oracle; size-independent properties at the benchmark size.
Runs on the B200 box only (-m gpu)."""
import numpy as np
import pytest
import torch

from conftest import golden_ctor, golden_state, load_golden
from oracle import nsd_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import neural_speech_decoder_b200 as nsd
    from neural_speech_decoder_b200.synthetic import fill_trained_like_, make_batch

DEV = "cuda"
# fp32 mode: north-star tolerance rtol 1e-3 on logits / loss / gradients (atol because logits cross zero)
RTOL, ATOL = 1e-3, 2e-5


def build(g, precision="fp32", dropout=None):
    kw = dict(golden_ctor(g))
    if dropout is not None:
        kw["dropout"] = dropout
    nsd.set_default_precision(precision)
    torch.manual_seed(0)
    m = nsd.GRUDecoder(device=DEV, **kw)
    nsd.set_default_precision("fp32")
    sd = golden_state(g)
    if sd:
        own = m.state_dict()
        for k, v in sd.items():
            own[k].copy_(torch.from_numpy(v))
    else:
        fill_trained_like_(m, seed=7)                     # comp_* fixtures: weights regenerated from the seed
    return m.to(DEV), kw


def run_train_lines(m, g):
    X, day = torch.from_numpy(g["X"]).to(DEV), torch.from_numpy(g["dayIdx"]).to(DEV)
    y, X_len, y_len = (torch.from_numpy(g[k]).to(DEV) for k in ("y", "X_len", "y_len"))
    pred = m.forward(X, day)                                                     # trainer:208
    pred.retain_grad()
    lens = ((X_len - m.kernelLen) / m.strideLen).to(torch.int32)                 # trainer:209
    log_probs = pred.log_softmax(2).permute(1, 0, 2)                             # trainer:210
    loss = nsd.CTCLoss(blank=0, reduction="mean", zero_infinity=True)(log_probs, y, lens, y_len)
    loss = torch.sum(loss)                                                       # trainer:242
    loss.backward()                                                              # trainer:252
    return pred, log_probs, loss, lens


@pytest.mark.parametrize("name", ["small_uni", "small_bi"])
def test_small_models_match_reference_fixture(name):
    g = load_golden(name + "_f32")
    m, kw = build(g)
    m.eval()
    pred, log_probs, loss, lens = run_train_lines(m, g)
    m.check_errors()
    np.testing.assert_allclose(pred.detach().cpu().numpy(), g["logits"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(loss.item(), g["loss"], rtol=RTOL)
    assert np.array_equal(lens.cpu().numpy(), g["out_lens"])
    np.testing.assert_allclose(pred.grad.cpu().numpy(), g["dlogits"], rtol=RTOL, atol=1e-6)
    names = [str(n) for n in g["live_grad_names"]]
    params = dict(m.named_parameters())
    assert sorted(k for k, p in params.items() if p.grad is not None) == names
    for n in names:
        ref = g["grad." + n]
        np.testing.assert_allclose(params[n].grad.cpu().numpy(), ref, rtol=RTOL, atol=1e-5 * max(1.0, np.abs(ref).max()), err_msg=n)
    # decode on the reference's own log-probs: bit-exact sequences
    dec, dec_len = nsd.greedy_decode(torch.from_numpy(g["log_probs_tbc"]).to(DEV), lens)
    got = nsd.decoded_to_lists(dec, dec_len)
    assert got == [g["decoded"][b, :g["decoded_len"][b]].tolist() for b in range(len(got))]


@pytest.mark.parametrize("name", ["comp_uni", "comp_bi"])
def test_competition_shape_matches_reference_fixture(name):
    """256 feats / 24 days / 5x1024 / k32 s4 / 41 classes (BASELINE configs[0] architecture) at B=2, T=80."""
    g = load_golden(name + "_f32")
    m, kw = build(g)
    m.eval()
    pred, log_probs, loss, lens = run_train_lines(m, g)
    np.testing.assert_allclose(pred.detach().cpu().numpy(), g["logits"], rtol=RTOL, atol=1e-4)
    np.testing.assert_allclose(loss.item(), g["loss"], rtol=RTOL)
    params = dict(m.named_parameters())
    for n in [str(n) for n in g["live_grad_names"]]:
        gr = params[n].grad.detach().reshape(-1)
        step = max(1, gr.numel() // 4096)
        ref = g["gsample." + n]
        np.testing.assert_allclose(gr[::step][:4096].cpu().numpy(), ref, rtol=RTOL, atol=2e-5 * max(1.0, np.abs(ref).max()), err_msg=n)
        s_ref = g["gsum." + n]
        np.testing.assert_allclose((gr.double() ** 2).sum().item(), s_ref[1], rtol=2e-3, err_msg=n)
    dec, dec_len = nsd.greedy_decode(log_probs.detach(), lens)
    assert nsd.decoded_to_lists(dec, dec_len) == [g["decoded"][b, :g["decoded_len"][b]].tolist() for b in range(2)]


def test_state_dict_is_reference_compatible():
    g = load_golden("small_bi_f32")
    m, kw = build(g)
    keys = set(m.state_dict().keys())
    want = {"dayWeights", "dayBias", "gaussianSmoother.weight", "fc_decoder_out.weight", "fc_decoder_out.bias"}
    for l in range(kw["layer_dim"]):
        for sfx in ("", "_reverse"):
            for p in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                want.add(f"gru_decoder.{p}_l{l}{sfx}")
    for d in range(kw["nDays"]):
        want |= {f"inpLayer{d}.weight", f"inpLayer{d}.bias"}
    assert keys == want
    np.testing.assert_allclose(m.gaussianSmoother.weight[0, 0].cpu().numpy(), g["taps"], rtol=1e-6)


def test_errors_mirror_reference():
    g = load_golden("small_uni_f32")
    m, kw = build(g)
    X = torch.from_numpy(g["X"]).to(DEV)
    with pytest.raises(RuntimeError):
        m.forward(X[:, :kw["kernelLen"] - 1], torch.zeros(X.shape[0], dtype=torch.int64, device=DEV))   # nn.Unfold: T < K
    with pytest.raises(IndexError):
        m.forward(X, torch.full((X.shape[0],), kw["nDays"], dtype=torch.int64))                         # host indices: eager
    m.forward(X, torch.full((X.shape[0],), kw["nDays"], dtype=torch.int64, device=DEV))               # device indices: deferred
    with pytest.raises(IndexError):
        m.check_errors()
    with pytest.raises(nsd.NsdError):
        m.forward(X.cpu(), torch.zeros(X.shape[0], dtype=torch.int64))                                  # no CPU path
    with pytest.raises(ZeroDivisionError):
        nsd.GRUDecoder(16, 10, 32, 1, gaussianSmoothWidth=0)                                            # augmentations.py:57


def test_dropout_train_mode_runs_and_eval_is_deterministic():
    g = load_golden("small_bi_f32")
    m, kw = build(g, dropout=0.4)
    X, day = torch.from_numpy(g["X"]).to(DEV), torch.from_numpy(g["dayIdx"]).to(DEV)
    m.train()
    a, b = m.forward(X, day), m.forward(X, day)
    assert not torch.equal(a, b)                       # fresh mask every call
    m.eval()
    c, d = m.forward(X, day), m.forward(X, day)
    assert torch.equal(c, d)
    np.testing.assert_allclose(c.detach().cpu().numpy(), g["logits"], rtol=RTOL, atol=ATOL)
    m.train()
    loss = nsd.ctc_loss_from_logits(m.forward(X, day), torch.from_numpy(g["y"]).to(DEV),
                                    torch.from_numpy(g["out_lens"]).to(DEV), torch.from_numpy(g["y_len"]).to(DEV))
    loss.backward()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)


def test_train_step_reduces_loss_and_eval_batch():
    g = load_golden("small_uni_f32")
    m, kw = build(g)
    m.train()
    batch = [torch.from_numpy(g[k]).to(DEV) for k in ("X", "y", "X_len", "y_len", "dayIdx")]
    opt, sched = nsd.make_optimizer(m, dict(lrStart=0.02, lrEnd=0.02, nBatch=100, l2_decay=1e-5))
    losses = [nsd.train_step(m, opt, *batch, scheduler=sched).item() for _ in range(8)]
    assert losses[-1] < losses[0]
    # LossReader: the loss read on a side stream right after the forward equals the returned device scalar, step after step
    reader = nsd.LossReader(DEV)
    for _ in range(3):
        dev_loss = nsd.train_step(m, opt, *batch, scheduler=sched, loss_reader=reader)
        assert reader.item() == dev_loss.item()
    m.eval()
    loss, dist, tot = nsd.eval_batch(m, *batch)
    assert tot == int(g["y_len"].sum()) and 0 <= dist <= tot + int(g["out_lens"].sum())


def test_optimizer_under_the_recurrence_is_bit_identical(monkeypatch):
    """Single GPU, bf16: train_step updates each finished bucket from inside the backward, behind the next BPTT launch (the update runs under
    it).  Losses and every parameter after several steps must equal the classic order (whole update after the backward) bit for bit, with
    and without dropout; and the in-backward path must really have been taken."""
    from neural_speech_decoder_b200 import _lib
    kw = dict(neural_dim=32, n_classes=10, hidden_dim=64, layer_dim=3, nDays=4, dropout=0.3, strideLen=4, kernelLen=16,
              gaussianSmoothWidth=2.0, bidirectional=True)
    X, y, X_len, y_len, day = make_batch(6, 90, n_feat=32, n_days=4, n_classes=10, seed=7, ragged=True, min_tgt=2, max_tgt=6,
                                         kernel_len=16, stride_len=4)
    batch = [t.to(DEV) for t in (X, y, X_len, y_len, day)]
    runs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("NSD_STEP_IN_BACKWARD", flag)
        nsd.set_default_precision("bf16")
        try:
            torch.manual_seed(0)
            m = nsd.GRUDecoder(device=DEV, **kw)
        finally:
            nsd.set_default_precision("fp32")
        fill_trained_like_(m, seed=3)
        m = m.to(DEV)
        m.train()
        opt, sched = nsd.make_optimizer(m, dict(lrStart=0.02, lrEnd=0.01, nBatch=10, l2_decay=1e-5))
        _lib.profile_begin({"nsd_adam_step"})
        losses = [nsd.train_step(m, opt, *batch, scheduler=sched, white_noise_sd=0.3).item() for _ in range(4)]
        calls = sum(n for n, _ in _lib.profile_end().values())
        runs.append((losses, [p.detach().clone() for p in m.parameters()], calls))
    assert runs[0][0] == runs[1][0]
    assert all(torch.equal(a, b) for a, b in zip(runs[0][1], runs[1][1]))
    assert runs[1][2] == 4 and runs[0][2] == 4 * (kw["layer_dim"] + 1)      # fc + layers L-1..1 in the backward, layer 0 + day weights after it


def test_bf16_weight_shadows_follow_the_parameters():
    """bf16 mode keeps bf16 operand copies of the weights across steps; FusedAdam rewrites them in its update kernel.
    Training with the copies maintained by Adam must be bit-identical to training that re-casts every step, and any
    other in-place change of a parameter (load_state_dict, manual edit) must invalidate the copy."""
    from neural_speech_decoder_b200 import _lib
    kw = dict(neural_dim=32, n_classes=10, hidden_dim=64, layer_dim=2, nDays=4, dropout=0.0, strideLen=4, kernelLen=16,
              gaussianSmoothWidth=2.0, bidirectional=True)
    X, y, X_len, y_len, day = make_batch(4, 72, n_feat=32, n_days=4, n_classes=10, seed=5, ragged=True, min_tgt=2, max_tgt=6,
                                         kernel_len=16, stride_len=4)
    batch = [t.to(DEV) for t in (X, y, X_len, y_len, day)]

    def build_bf16():
        nsd.set_default_precision("bf16")
        try:
            torch.manual_seed(0)
            mm = nsd.GRUDecoder(device=DEV, **kw)
        finally:
            nsd.set_default_precision("fp32")
        fill_trained_like_(mm, seed=3)
        return mm.to(DEV)

    runs = []
    for attach in (True, False):
        m = build_bf16()
        m.train()
        opt, sched = nsd.make_optimizer(m, dict(lrStart=0.02, lrEnd=0.02, nBatch=100, l2_decay=1e-5))
        if not attach:
            opt.attach_shadows(None)
        losses = [nsd.train_step(m, opt, *batch, scheduler=sched).item() for _ in range(4)]
        runs.append((losses, [p.detach().clone() for p in m.parameters()], m, opt))
    assert runs[0][0] == runs[1][0]
    assert all(torch.equal(a, b) for a, b in zip(runs[0][1], runs[1][1]))
    # attached: a step does no weight casts (only the W_hh^T transposes of the bf16 copies and the dlogits cast)
    m, opt = runs[0][2], runs[0][3]
    _lib.profile_begin({"nsd_cast_transpose"})
    nsd.train_step(m, opt, *batch)
    with_shadows = sum(n for n, _ in _lib.profile_end().values())
    m2, opt2 = runs[1][2], runs[1][3]
    _lib.profile_begin({"nsd_cast_transpose"})
    nsd.train_step(m2, opt2, *batch)
    without = sum(n for n, _ in _lib.profile_end().values())
    L, D = kw["layer_dim"], 2
    assert without - with_shadows == 2 * L * D + 1                      # W_ih, W_hh per layer-direction and the output layer
    # an in-place edit outside the optimizer invalidates the copy
    m.eval()
    X, day = batch[0], batch[4]
    before = m.forward(X, day).detach().clone()
    with torch.no_grad():
        m.gru_decoder.weight_ih_l0.mul_(0.5)
    after = m.forward(X, day).detach()
    assert not torch.equal(before, after)
    ref = build_bf16()
    ref.load_state_dict(m.state_dict())
    ref.eval()
    assert torch.equal(ref.forward(X, day).detach(), after)


def test_batch_prefetcher_yields_identical_batches_in_order():
    """trainer:185-191 replacement: batches staged on a copy stream arrive complete and in order."""
    host = []
    for i in range(5):
        X, y, X_len, y_len, day = make_batch(3, 64 + 4 * i, n_feat=32, n_days=4, n_classes=10, seed=i, kernel_len=16, stride_len=4)
        host.append(tuple(t.pin_memory() for t in (X, y, X_len, y_len, day)))
    got = []
    for dev_batch in nsd.BatchPrefetcher(iter(host), DEV):
        assert all(t.is_cuda for t in dev_batch)
        got.append(tuple(t.clone() for t in dev_batch))
    torch.cuda.synchronize()
    assert len(got) == 5
    for h, g in zip(host, got):
        assert all(torch.equal(a, b.cpu()) for a, b in zip(h, g))
    with pytest.raises(RuntimeError):
        nsd.BatchPrefetcher(iter(host), "cpu")
