"""Generate tests/golden/conformer_*.npz from the REFERENCE itself (build container only: needs /root/reference).

    python tests/golden/make_golden_conformer.py

Imports the unmodified ``neural_decoder.transformer_ctc.NeuralTransformerCTCModel`` on CPU, perturbs its seeded
initialisation so that no parameter sits at a symmetric value (identity day weights, unit LayerNorm gains, zero biases),
runs the reference forward in train mode with every stochastic regulariser at probability 0 (dropout, DropPath, SpecAugment:
deterministic) and in eval mode, then the trainer's loss lines (neural_decoder_trainer.py:137-141, 212-249, restated here
because the trainer module needs ``edit_distance`` / ``hydra`` / ``wandb``, which are not installed) and backward.  Stores the
state dict, inputs, log-probs, InterCTC log-probs, output lengths, loss and every gradient.
"""
from __future__ import annotations

import math
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")
warnings.filterwarnings("ignore")

from neural_decoder.transformer_ctc import NeuralTransformerCTCModel as RefModel  # noqa: E402

CONFIGS = {
    # name: (ctor kwargs, B, T, label_smoothing, interctc_weight)
    "conformer_small": (dict(n_channels=32, n_classes=11, n_days=3, frontend_dim=64, latent_dim=64, autoencoder_hidden_dim=32,
                             transformer_layers=6, transformer_heads=4, transformer_ff_dim=128, transformer_dropout=0.0,
                             temporal_kernel=16, temporal_stride=4, gaussian_smooth_width=2.0, conformer_conv_kernel=7,
                             use_spec_augment=False, drop_path_prob=0.0), 3, 84, 0.1, 0.3),
    # no InterCTC (< 6 layers), plain mean-reduced CTC (label_smoothing 0): the other branch of trainer:137-141, 235-249
    "conformer_shallow": (dict(n_channels=32, n_classes=11, n_days=3, frontend_dim=64, latent_dim=64, autoencoder_hidden_dim=32,
                               transformer_layers=2, transformer_heads=2, transformer_ff_dim=96, transformer_dropout=0.0,
                               temporal_kernel=8, temporal_stride=2, gaussian_smooth_width=1.0, conformer_conv_kernel=5,
                               use_spec_augment=False, drop_path_prob=0.0), 4, 50, 0.0, 0.3),
}


def make_inputs(B, T, n_ch, n_days, n_classes, kernel, stride, seed):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(B, T, n_ch, generator=g)
    day = torch.randint(0, n_days, (B,), generator=g)
    X_len = torch.randint(max(kernel + 4 * stride, T // 2), T + 1, (B,), generator=g).to(torch.int32)
    X_len[0] = T
    frames = ((X_len - kernel) / stride).to(torch.int32)
    y_len = torch.clamp(torch.randint(2, 9, (B,), generator=g).to(torch.int32), max=(frames // 2).clamp(min=1))
    y = torch.zeros(B, int(y_len.max()), dtype=torch.int32)
    for b in range(B):
        y[b, :y_len[b]] = torch.randint(1, n_classes, (int(y_len[b]),), generator=g).to(torch.int32)
    for b in range(B):
        X[b, X_len[b]:] = 0                                         # the dataset pads with zeros (neural_decoder_trainer.py:26-37)
    return X, y, X_len, y_len, day


def perturb_(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            scale = 0.05 if p.dim() >= 2 else 0.1
            p.add_(torch.randn(p.shape, generator=g) * scale * (p.abs().mean() + 0.1))


def run(name, kw, B, T, ls, iw, dtype):
    torch.manual_seed(0)
    model = RefModel(device="cpu", **kw)
    perturb_(model, 7)
    if dtype == torch.float64:
        model = model.double()
    X, y, X_len, y_len, day = make_inputs(B, T, kw["n_channels"], kw["n_days"], kw["n_classes"], kw["temporal_kernel"],
                                          kw["temporal_stride"], 11)
    Xd = X.to(dtype)
    out = {"X": X.numpy(), "y": y.numpy(), "X_len": X_len.numpy(), "y_len": y_len.numpy(), "day": day.numpy()}
    model.eval()
    with torch.no_grad():
        lp_eval, olen_eval, inter_eval = model(Xd, day, X_len)
    assert inter_eval is None
    out["eval_log_probs"] = lp_eval.numpy()
    model.train()
    model.output[3].p = 0.0     # the deep head's Dropout(0.3) is hard-wired (transformer_ctc.py:413): switched off on this instance, like the rest
    log_probs, out_lens, inter = model(Xd, day, X_len)                                      # trainer:206
    n_classes = kw["n_classes"]
    loss_ctc = torch.nn.CTCLoss(blank=0, reduction="none" if ls > 0 else "mean", zero_infinity=True)   # trainer:137-141
    loss = loss_ctc(log_probs, y, out_lens, y_len)                                          # trainer:213-218
    inter_loss = None
    if inter is not None:                                                                   # trainer:222-232
        inter_loss = loss_ctc(inter, y, out_lens, y_len)
        inter_loss = torch.mean(inter_loss) if ls > 0 else torch.sum(inter_loss)
    if ls > 0:                                                                              # trainer:235-243
        ctc_loss = torch.mean(loss)
        uniform = torch.full_like(log_probs, -math.log(n_classes))
        kl = torch.nn.functional.kl_div(log_probs, uniform, reduction="batchmean", log_target=True)
        main = (1 - ls) * ctc_loss + ls * kl
    else:
        main = torch.sum(loss)
    total = (1.0 - iw) * main + iw * inter_loss if inter is not None else main              # trainer:246-249
    model.zero_grad()
    total.backward()                                                                        # trainer:252-253
    out.update({"log_probs": log_probs.detach().numpy(), "out_lens": out_lens.numpy(), "loss": np.asarray(total.item()),
                "label_smoothing": np.asarray(ls), "interctc_weight": np.asarray(iw)})
    if inter is not None:
        out["inter_log_probs"] = inter.detach().numpy()
    gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in model.parameters() if p.grad is not None))
    out["grad_norm"] = np.asarray(gn.item())
    if dtype == torch.float32:                                       # the fp32 run only contributes its outputs (what fp32 arithmetic gives)
        return {"f32/" + k: out[k] for k in ("log_probs", "eval_log_probs", "loss", "grad_norm")}
    for k, v in model.state_dict().items():                          # weights were built in fp32 and widened: fp32 storage is exact
        v = v.detach()
        if k == "pos_enc.pe":
            v = v[:, :256]                                            # the rows a test can reach (the buffer has 5000)
        assert torch.equal(v.float().double(), v.double()), k
        out["sd/" + k] = v.float().numpy()
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        out["grad/" + k] = p.grad.numpy()
    return out


if __name__ == "__main__":
    for name, (kw, B, T, ls, iw) in CONFIGS.items():
        out = run(name, kw, B, T, ls, iw, torch.float64)
        out.update(run(name, kw, B, T, ls, iw, torch.float32))
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{path}: loss {out['loss']:.10f} (fp32 run {out['f32/loss']:.8f}), |grad| {out['grad_norm']:.8f}, frames {out['log_probs'].shape[0]}, "
              f"out_lens {out['out_lens'].tolist()}, {os.path.getsize(path) / 1e6:.2f} MB")
