"""Pin the numpy oracle against fixtures produced by the reference itself
(tests/golden/make_golden.py).  CPU only."""
import os
import numpy as np
import pytest

from conftest import golden_ctor, golden_state, load_golden
from oracle import nsd_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SMALL = ["small_uni", "small_bi"]


@pytest.mark.parametrize("name", SMALL)
@pytest.mark.parametrize("tag,tol", [("f64", 1e-11), ("f32", 2e-5)])
def test_stages_match_reference(name, tag, tol):
    g = load_golden(f"{name}_{tag}")
    kw = golden_ctor(g)
    sd = golden_state(g)
    dt = np.float64
    X = g["X"].astype(dt)
    taps = g["taps"].astype(dt)
    np.testing.assert_allclose(O.gaussian_taps(kw["gaussianSmoothWidth"]), g["taps"].astype(np.float32), rtol=2e-6, atol=1e-9)
    ys = O.smooth(X, taps)
    np.testing.assert_allclose(ys, g["smoothed"], rtol=tol, atol=tol)
    z = O.softsign(O.day_affine(ys, sd["dayWeights"].astype(dt), sd["dayBias"].astype(dt), g["dayIdx"]))
    np.testing.assert_allclose(z, g["z"], rtol=tol, atol=tol)
    p = O.unfold(z, kw["kernelLen"], kw["strideLen"])
    np.testing.assert_allclose(p, g["patches"], rtol=tol, atol=tol)
    # unfold is pure data movement: exact on the reference's own z
    assert np.array_equal(O.unfold(g["z"], kw["kernelLen"], kw["strideLen"]), g["patches"])
    ws = [{k: v.astype(dt) for k, v in d.items()} for d in O.split_gru_state(sd, kw["layer_dim"], kw["bidirectional"])]
    hid, _ = O.gru_fwd(g["patches"].astype(dt), ws, kw["bidirectional"])
    np.testing.assert_allclose(hid, g["hid"], rtol=tol * 10, atol=tol * 10)
    logits = O.linear(hid, sd["fc_decoder_out.weight"].astype(dt), sd["fc_decoder_out.bias"].astype(dt))
    np.testing.assert_allclose(logits, g["logits"], rtol=tol * 10, atol=tol * 10)
    lp = O.log_softmax(g["logits"].astype(dt), axis=2)
    np.testing.assert_allclose(np.transpose(lp, (1, 0, 2)), g["log_probs_tbc"], rtol=tol, atol=tol)
    assert np.array_equal(O.out_lens(g["X_len"], kw["kernelLen"], kw["strideLen"]), g["out_lens"])


@pytest.mark.parametrize("name", SMALL + ["comp_uni", "comp_bi"])
def test_ctc_and_decode_match_reference(name):
    g = load_golden(f"{name}_f32")
    loss, nll, grad = O.ctc_loss(g["log_probs_tbc"], g["y"], g["out_lens"], g["y_len"])
    np.testing.assert_allclose(loss, g["loss"], rtol=2e-6)
    np.testing.assert_allclose(nll, g["nll"], rtol=2e-6, atol=1e-5)
    np.testing.assert_allclose(grad, g["dlog_probs_tbc"], rtol=1e-4, atol=2e-7)
    dec = O.greedy_decode(g["log_probs_tbc"], g["out_lens"])
    for b, seq in enumerate(dec):
        assert seq == g["decoded"][b, : g["decoded_len"][b]].tolist()


@pytest.mark.parametrize("name", SMALL)
def test_full_backward_matches_reference(name):
    g = load_golden(f"{name}_f64")
    kw = golden_ctor(g)
    sd = golden_state(g)
    loss, logits, grads = O.train_loss_and_grads(
        sd, g["X"], g["dayIdx"], g["y"], g["X_len"], g["y_len"], kernel_len=kw["kernelLen"],
        stride_len=kw["strideLen"], n_layers=kw["layer_dim"], bidirectional=kw["bidirectional"])
    np.testing.assert_allclose(loss, g["loss"], rtol=1e-10)
    np.testing.assert_allclose(logits, g["logits"], rtol=1e-9, atol=1e-11)
    names = [str(n) for n in g["live_grad_names"]]
    assert sorted(grads.keys()) == names
    for n in names:
        np.testing.assert_allclose(grads[n], g["grad." + n], rtol=1e-7, atol=1e-11, err_msg=n)


def test_ctc_edge_cases_against_torch():
    """empty target, repeated labels, infeasible (zero_infinity), input shorter than T."""
    import torch
    rng = np.random.default_rng(3)
    T, B, C = 12, 5, 6
    lp = O.log_softmax(rng.standard_normal((T, B, C)), axis=2)
    y = np.array([[1, 1, 2, 0, 0, 0], [3, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0], [2, 2, 2, 2, 2, 2], [5, 4, 5, 4, 0, 0]], dtype=np.int32)
    yl = np.array([3, 1, 0, 6, 4], dtype=np.int32)
    il = np.array([12, 7, 5, 9, 12], dtype=np.int32)       # utt 3 infeasible: needs 11 frames, has 9
    loss, nll, grad = O.ctc_loss(lp, y, il, yl)
    t = torch.tensor(lp, requires_grad=True)
    ref = torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)(t, torch.tensor(y), torch.tensor(il), torch.tensor(yl))
    ref.backward()
    np.testing.assert_allclose(loss, ref.item(), rtol=1e-12)
    np.testing.assert_allclose(grad, t.grad.numpy(), rtol=1e-9, atol=1e-14)
    assert nll[3] == 0.0 and np.all(grad[:, 3] == 0)


def test_edit_distance():
    assert O.edit_distance([1, 2, 3], [1, 2, 3]) == 0
    assert O.edit_distance([], [4, 5]) == 2
    assert O.edit_distance([1, 2, 3, 4], [2, 3, 5]) == 2
    assert O.edit_distance([7], []) == 1
    assert O.phoneme_error_rate([[1, 2], [3]], np.array([[1, 2, 0], [4, 5, 6]]), [2, 3]) == (3, 5)


def test_frontend_errors():
    with pytest.raises(ZeroDivisionError):
        O.gaussian_taps(0)
    with pytest.raises(RuntimeError):
        O.unfold(np.zeros((1, 5, 2)), 8, 2)
    with pytest.raises(IndexError):
        O.day_affine(np.zeros((1, 4, 2)), np.zeros((2, 2, 2)), np.zeros((2, 1, 2)), np.array([2]))


@pytest.mark.parametrize("name", SMALL)
def test_torch_port_matches_reference_fixture(name):
    """bench.py's CPU baseline ("port") reproduces the reference's own outputs."""
    import torch
    from oracle import torch_port as P
    g = load_golden(f"{name}_f32")
    kw = golden_ctor(g)
    m = P.PortGRUDecoder(**kw)
    m.load_reference_state(golden_state(g))
    m.eval()
    X, y, X_len, y_len, day = (torch.from_numpy(g[k]) for k in ("X", "y", "X_len", "y_len", "dayIdx"))
    pred = m(X, day)
    np.testing.assert_allclose(pred.detach().numpy(), g["logits"], rtol=1e-5, atol=1e-6)
    loss = P.train_step(m, P.make_adam(m), X, y, X_len, y_len, day)
    np.testing.assert_allclose(loss.item(), g["loss"], rtol=1e-5)


def test_compiled_reference_equals_port():
    """oracle/_ref (the reference's own GRUDecoder, byte-compiled from /root/reference by oracle/build_ref.py) and the
    torch-operator port give the same logits for the same state: bench.py's "reference" and "port" kinds are interchangeable."""
    import torch
    from oracle import torch_port as P
    from oracle.build_ref import load_reference_decoder
    Ref = load_reference_decoder()
    if Ref is None:
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py needs /root/reference)")
    kw = dict(neural_dim=16, n_classes=6, hidden_dim=24, layer_dim=2, nDays=3, dropout=0.0, strideLen=4, kernelLen=8,
              gaussianSmoothWidth=2.0, bidirectional=True)
    torch.manual_seed(3)
    ref = Ref(device="cpu", **kw).eval()
    port = P.PortGRUDecoder(**kw).eval()
    port.load_reference_state(ref.state_dict())
    X, day = torch.randn(3, 50, 16), torch.tensor([0, 2, 1])
    with torch.no_grad():
        a, b = ref(X, day), port(X, day)
    assert torch.allclose(a, b, rtol=0, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------------- Conformer
CONFORMER_CFG = {
    "conformer_small": dict(n_layers=6, n_heads=4, temporal_kernel=16, temporal_stride=4, conv_kernel=7),
    "conformer_shallow": dict(n_layers=2, n_heads=2, temporal_kernel=8, temporal_stride=2, conv_kernel=5),
}


def _load_conformer(name, dtype):
    import torch
    g = np.load(os.path.join(GOLD, name + ".npz"))
    sd = {k[3:]: torch.from_numpy(g[k]).to(dtype) for k in g.files if k.startswith("sd/")}
    return g, sd


@pytest.mark.parametrize("name", sorted(CONFORMER_CFG))
def test_conformer_port_matches_reference_fixture(name):
    """oracle/conformer_port.py (functional restatement over torch operators) against outputs of the imported reference
    (tests/golden/make_golden_conformer.py): fp64 log-probs 1e-12, loss 1e-12, every gradient 1e-9 relative to its tensor's
    largest entry; eval-mode forward; output lengths exact."""
    import torch
    from oracle import conformer_port as CP
    g, sd = _load_conformer(name, torch.float64)
    cfg = CONFORMER_CFG[name]
    X, day = torch.from_numpy(g["X"]).double(), torch.from_numpy(g["day"])
    X_len, y, y_len = torch.from_numpy(g["X_len"]), torch.from_numpy(g["y"]), torch.from_numpy(g["y_len"])
    with torch.no_grad():
        lp_eval, olen, inter = CP.forward(sd, X, day, X_len, training=False, **cfg)
    assert inter is None and olen.dtype == torch.int32 and olen.tolist() == g["out_lens"].tolist()
    np.testing.assert_allclose(lp_eval.numpy(), g["eval_log_probs"], rtol=0, atol=1e-12)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if ("grad/" + k) in g.files}
    lp, olen, inter = CP.forward({**sd, **params}, X, day, X_len, training=True, **cfg)
    np.testing.assert_allclose(lp.detach().numpy(), g["log_probs"], rtol=0, atol=1e-12)
    assert (inter is not None) == ("inter_log_probs" in g.files)
    if inter is not None:
        np.testing.assert_allclose(inter.detach().numpy(), g["inter_log_probs"], rtol=0, atol=1e-12)
    loss = CP.training_loss(lp, inter, y, olen, y_len, label_smoothing=float(g["label_smoothing"]), interctc_weight=float(g["interctc_weight"]))
    np.testing.assert_allclose(loss.item(), float(g["loss"]), rtol=1e-12)
    loss.backward()
    for k, p in params.items():
        ref = g["grad/" + k]
        got = p.grad.numpy() if p.grad is not None else np.zeros_like(ref)
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-9 * max(1.0, np.abs(ref).max()), err_msg=k)
    gn = np.sqrt(sum((p.grad.numpy() ** 2).sum() for p in params.values() if p.grad is not None))
    np.testing.assert_allclose(gn, float(g["grad_norm"]), rtol=1e-10)


def test_conformer_lr_schedule_matches_trainer_lambda():
    """trainer:154-158 warm-up + cosine factor."""
    from oracle import conformer_port as CP
    assert CP.lr_factor(0, 1000, 15000) == 1 / 1000 and CP.lr_factor(999, 1000, 15000) == 1.0
    assert abs(CP.lr_factor(1000, 1000, 15000) - 1.0) < 1e-15 and abs(CP.lr_factor(8000, 1000, 15000) - 0.5) < 1e-12
    assert abs(CP.lr_factor(15000, 1000, 15000)) < 1e-15
