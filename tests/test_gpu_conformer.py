"""Conformer path (BASELINE configs[2]) on the CUDA kernels against (a) fixtures generated from the imported reference
(tests/golden/conformer_*.npz) and (b) the oracle port (oracle/conformer_port.py) on the same seeded inputs at the
competition architecture.  fp32 mode: log-probs / loss / every gradient within rtol 1e-3; bf16 mode: stated tolerances."""
import math
import os

import numpy as np
import pytest
import torch

import neural_speech_decoder_b200 as nsd
from neural_speech_decoder_b200 import conformer as CF
from neural_speech_decoder_b200._lib import call, ptr, stream

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0) if torch.cuda.is_available() else None
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CFG = {
    "conformer_small": dict(n_channels=32, n_classes=11, n_days=3, frontend_dim=64, latent_dim=64, autoencoder_hidden_dim=32,
                            transformer_layers=6, transformer_heads=4, transformer_ff_dim=128, transformer_dropout=0.0, temporal_kernel=16,
                            temporal_stride=4, gaussian_smooth_width=2.0, conformer_conv_kernel=7, use_spec_augment=False, drop_path_prob=0.0),
    "conformer_shallow": dict(n_channels=32, n_classes=11, n_days=3, frontend_dim=64, latent_dim=64, autoencoder_hidden_dim=32,
                              transformer_layers=2, transformer_heads=2, transformer_ff_dim=96, transformer_dropout=0.0, temporal_kernel=8,
                              temporal_stride=2, gaussian_smooth_width=1.0, conformer_conv_kernel=5, use_spec_augment=False, drop_path_prob=0.0),
}


def _load(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    return g, {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}


def _build(name, precision, sd):
    m = nsd.NeuralTransformerCTCModel(device="cuda", precision=precision, **CFG[name])
    own = m.state_dict()
    sd = dict(sd)
    sd["pos_enc.pe"] = own["pos_enc.pe"]                       # the fixture keeps the first 256 rows of the 5000-row buffer
    assert torch.equal(own["pos_enc.pe"][:, :256].cpu(), torch.from_numpy(np.load(os.path.join(GOLD, name + ".npz"))["sd/pos_enc.pe"]))
    m.load_state_dict(sd, strict=True)
    m.output[3].p = 0.0                                        # the fixture switched the head's hard-wired Dropout(0.3) off
    return m.to(DEV)


@pytest.mark.parametrize("name", sorted(CFG))
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_conformer_matches_reference_fixture(name, precision):
    g, sd = _load(name)
    m = _build(name, precision, sd)
    X, day = torch.from_numpy(g["X"]).to(DEV), torch.from_numpy(g["day"]).to(DEV)
    X_len, y, y_len = torch.from_numpy(g["X_len"]).to(DEV), torch.from_numpy(g["y"]).to(DEV), torch.from_numpy(g["y_len"]).to(DEV)
    atol_lp, rtol_loss, gtol = (2e-6, 1e-6, 5e-6) if precision == "fp32" else (0.02, 5e-4, 0.35)
    m.eval()
    with torch.no_grad():
        lp, olen, inter = m(X, day, X_len)
    assert inter is None and olen.dtype == torch.int32 and olen.tolist() == g["out_lens"].tolist()
    e_eval = np.abs(lp.cpu().numpy() - g["eval_log_probs"]).max()
    m.train()
    lp, olen, inter = m(X, day, X_len)
    assert lp.shape == g["log_probs"].shape and (inter is not None) == ("inter_log_probs" in g.files)
    e_lp = np.abs(lp.detach().cpu().numpy() - g["log_probs"]).max()
    e_in = np.abs(inter.detach().cpu().numpy() - g["inter_log_probs"]).max() if inter is not None else 0.0
    loss = nsd.conformer_loss(lp, inter, y, olen, y_len, float(g["label_smoothing"]), float(g["interctc_weight"]))
    loss.backward()
    e_loss = abs(loss.item() - float(g["loss"])) / abs(float(g["loss"]))
    worst, worst_k, n = 0.0, None, 0
    for k, p in m.named_parameters():
        ref = g["grad/" + k]
        assert p.grad is not None, k
        err = np.linalg.norm(p.grad.cpu().numpy().astype(np.float64) - ref) / max(np.linalg.norm(ref), 1e-12 * ref.size ** 0.5 + 1e-9)
        n += 1
        if err > worst:
            worst, worst_k = err, k
    print(f"{name} [{precision}]: eval log-probs max abs err {e_eval:.2e}, train {e_lp:.2e}, InterCTC {e_in:.2e}, loss rel {e_loss:.2e}, "
          f"worst gradient rel-L2 {worst:.2e} ({worst_k}) over {n} tensors")
    assert e_eval < atol_lp and e_lp < atol_lp and e_in < atol_lp and e_loss < rtol_loss and worst < gtol


def test_same_seed_same_weights_and_adamw_clip_step():
    """FusedAdamW + device-side clip_grad_norm_ against torch.optim.AdamW + torch.nn.utils.clip_grad_norm_ on the same gradients
    (trainer:144-151, 255-259), three steps, with the warm-up / cosine factor applied to the lr."""
    torch.manual_seed(3)
    ps = [torch.randn(s, device=DEV).requires_grad_(True) for s in ((300, 70), (5000,), (17,), (64, 64, 3))]
    qs = [p.detach().clone().cpu().requires_grad_(True) for p in ps]
    ours = nsd.FusedAdamW(ps, lr=4e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=1e-3, max_grad_norm=1.0)
    ref = torch.optim.AdamW(qs, lr=4e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=1e-3)
    for step in range(3):
        f = nsd.lr_lambda(step, 2, 10)
        for grp in ours.param_groups + ref.param_groups:
            grp["lr"] = 4e-4 * f
        for p, q in zip(ps, qs):
            gq = torch.randn(q.shape) * (3.0 if step != 1 else 1e-3)          # step 1: norm below the threshold -> no clipping
            q.grad = gq.clone()
            p.grad = gq.to(DEV)
        tn = torch.nn.utils.clip_grad_norm_(qs, max_norm=1.0)
        ref.step()
        ours.step()
        assert abs(math.sqrt(ours.grad_sqnorm.item()) - tn.item()) / tn.item() < 1e-5
        for p, q in zip(ps, qs):
            np.testing.assert_allclose(p.detach().cpu().numpy(), q.detach().numpy(), rtol=2e-5, atol=2e-7)


def test_dropout_masks_are_consistent_and_distributional():
    """The fused dropout / DropPath masks: keep fraction ~ 1-p, scale 1/(1-p), and the backward applies exactly the forward's mask."""
    M, D, p = 512, 256, 0.3
    x = torch.randn(M, D, device=DEV)
    ones = torch.ones(M, D, device=DEV)
    for seed in (1, 2):
        y = CF._Act.apply(ones, CF.ACT_NONE, p, seed, torch.float32)
        keep = (y != 0).float().mean().item()
        assert abs(keep - (1 - p)) < 0.01 and torch.allclose(y[y != 0], torch.tensor(1 / (1 - p), device=DEV))
        xr = x.clone().requires_grad_(True)
        out = CF._Act.apply(xr, CF.ACT_SILU, p, seed, torch.float32)
        out.backward(torch.ones_like(out))
        assert torch.equal((xr.grad != 0) | (x == 0), (y != 0) | (x == 0))
    r = CF._Residual.apply(torch.zeros(64 * 8, 32, device=DEV), torch.ones(64 * 8, 32, device=DEV), 0.5, 0.0, 0, 0.25, 7, 8 * 32)
    per = r.view(64, -1)
    assert all(len(set(row.tolist())) == 1 for row in per.cpu())                 # whole samples are kept or dropped
    vals = set(per[:, 0].cpu().tolist())
    assert len(vals) == 2 and 0.0 in vals and abs(max(vals) - 0.5 / 0.75) < 1e-6


@pytest.mark.parametrize("precision,B", [("fp32", 2), ("bf16", 8)])
def test_conformer_competition_architecture_vs_port(precision, B):
    """The reference's training configuration (256 channels, 24 days, d=1024, 8 layers x 8 heads, FF 2048, conv k31, k32/s4, 41 classes,
    T=500 -> T'=118) with ragged lengths, regularisers off: log-probs, InterCTC log-probs, loss and every gradient against the oracle
    port run on the host with the SAME weights and inputs."""
    from oracle import conformer_port as CP
    kw = dict(n_channels=256, n_classes=41, n_days=24, transformer_dropout=0.0, use_spec_augment=False, drop_path_prob=0.0)
    torch.manual_seed(0)
    m = nsd.NeuralTransformerCTCModel(device="cuda", precision=precision, **kw)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for p in m.parameters():                                      # off the symmetric initial values
            p.add_(torch.randn(p.shape, generator=g) * (0.02 if p.dim() >= 2 else 0.05) * (p.abs().mean() + 0.1))
    m.output[3].p = 0.0
    T = 500
    X = torch.randn(B, T, 256, generator=g)
    day = torch.randint(0, 24, (B,), generator=g)
    X_len = torch.randint(300, T + 1, (B,), generator=g).to(torch.int32); X_len[0] = T
    y_len = torch.randint(10, 40, (B,), generator=g).to(torch.int32)
    y = torch.zeros(B, int(y_len.max()), dtype=torch.int32)
    for b in range(B):
        y[b, :y_len[b]] = torch.randint(1, 41, (int(y_len[b]),), generator=g).to(torch.int32)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    params = {k: sd[k].clone().requires_grad_(True) for k, _ in m.named_parameters()}
    cfg = dict(n_layers=8, n_heads=8, temporal_kernel=32, temporal_stride=4, conv_kernel=31)
    lp_r, olen_r, inter_r = CP.forward({**sd, **params}, X, day, X_len, training=True, **cfg)
    loss_r = CP.training_loss(lp_r, inter_r, y, olen_r, y_len, label_smoothing=0.1, interctc_weight=0.3)
    loss_r.backward()
    m = m.to(DEV).train()
    lp, olen, inter = m(X.to(DEV), day.to(DEV), X_len.to(DEV))
    loss = nsd.conformer_loss(lp, inter, y.to(DEV), olen, y_len.to(DEV), 0.1, 0.3)
    loss.backward()
    assert olen.tolist() == olen_r.tolist()
    e_lp = (lp.detach().cpu() - lp_r.detach()).abs().max().item()
    e_in = (inter.detach().cpu() - inter_r.detach()).abs().max().item()
    e_loss = abs(loss.item() - loss_r.item()) / abs(loss_r.item())
    worst, worst_k = 0.0, None
    for k, p in m.named_parameters():
        ref = params[k].grad
        err = ((p.grad.cpu() - ref).norm() / ref.norm().clamp_min(1e-12)).item()
        if err > worst:
            worst, worst_k = err, k
    print(f"competition architecture [{precision}] B={B}: log-probs max abs err {e_lp:.2e}, InterCTC {e_in:.2e}, loss rel {e_loss:.2e}, "
          f"worst gradient rel-L2 {worst:.2e} ({worst_k})")
    if precision == "fp32":
        assert e_lp < 1e-5 and e_in < 1e-5 and e_loss < 1e-6 and worst < 5e-5      # measured 2.2e-6 / 3.1e-6 / 1.1e-7 / 1.4e-5
    else:
        assert e_lp < 0.03 and e_in < 0.03 and e_loss < 3e-4 and worst < 0.13       # measured 8.4e-3 / 8.2e-3 / 6.7e-5 / 4.3e-2 (temporal_conv.weight: sums 7.5k frames of bf16-rounded terms)


def test_graphed_step_equals_eager_steps_and_redraws_masks():
    """GraphedConformerStep (the whole training step as one replayed CUDA graph, per-step state in device memory): with the regularisers
    off, three graphed steps leave exactly the parameters three eager steps leave (same kernels, same order, warm-up / cosine lr from
    the device-resident schedule); with dropout on and lr = 0, two replays of the same batch give different losses (fresh masks)."""
    name = "conformer_small"
    g, sd = _load(name)
    X, day = torch.from_numpy(g["X"]).to(DEV), torch.from_numpy(g["day"]).to(DEV)
    X_len, y, y_len = torch.from_numpy(g["X_len"]).to(DEV), torch.from_numpy(g["y"]).to(DEV), torch.from_numpy(g["y_len"]).to(DEV)
    B, T = X.shape[0], X.shape[1]
    kw = dict(lr=4e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=1e-3, max_grad_norm=1.0)
    a, b = _build(name, "bf16", sd), _build(name, "bf16", sd)
    oa, ob = nsd.FusedAdamW(a.parameters(), **kw), nsd.FusedAdamW(b.parameters(), **kw)
    oa.attach_shadows(a._shadows); ob.attach_shadows(b._shadows)
    a.check_day_ids = False
    for step in range(3):                                                  # eager reference
        for grp in oa.param_groups:
            grp["lr"] = 4e-4 * nsd.lr_lambda(step, 2, 10)
        la = nsd.conformer_train_step(a, oa, X, y, X_len, y_len, day, 0.1, 0.3)
    gs = nsd.GraphedConformerStep(b, ob, B, T, int(y.shape[1]), label_smoothing=0.1, interctc_weight=0.3, base_lr=4e-4, warmup_steps=2, total_steps=10)
    for step in range(3):
        lb = gs.step(X, y, X_len, y_len, day)
    torch.cuda.synchronize()
    assert gs.graph is not None and gs.kernels_per_replay > 100 and gs.steps_done == 3
    assert abs(la.item() - lb.item()) < 1e-6 * abs(la.item())
    for (k, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.equal(p, q), k
    gs.close()
    # fresh masks on every replay
    cfg = dict(CFG[name]); cfg.update(transformer_dropout=0.3, drop_path_prob=0.1, use_spec_augment=True, spec_augment_freq_mask=20, spec_augment_time_mask=4)
    c = nsd.NeuralTransformerCTCModel(device="cuda", precision="bf16", **cfg).to(DEV)
    oc = nsd.FusedAdamW(c.parameters(), lr=0.0, weight_decay=0.0, max_grad_norm=1.0)
    gc = nsd.GraphedConformerStep(c, oc, B, T, int(y.shape[1]), base_lr=0.0, white_noise_sd=0.8, constant_offset_sd=0.2)
    losses = [gc.step(X, y, X_len, y_len, day).item() for _ in range(4)]
    gc.close()
    assert len(set(round(v, 6) for v in losses)) == 4, losses
