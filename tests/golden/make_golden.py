"""Generate the golden fixtures under tests/golden/ from the REFERENCE itself.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

It imports the unmodified ``neural_decoder.model.GRUDecoder`` from
``/root/reference/src`` on CPU, fills it with machine-independent weights
(``neural_speech_decoder_b200.synthetic``), runs the reference's forward and the
trainer's loss / backward / greedy-decode lines (neural_decoder_trainer.py:209-218,
242, 251-252, 313-320, restated verbatim here because the trainer module needs
``edit_distance``/``hydra``, which are not installed) and stores inputs,
per-stage intermediates (captured with forward hooks on the reference's own
sub-modules), outputs and gradients as .npz.
"""
from __future__ import annotations

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

from neural_decoder.model import GRUDecoder as RefGRUDecoder  # noqa: E402

from neural_speech_decoder_b200.synthetic import fill_trained_like_, make_batch  # noqa: E402

CONFIGS = {
    # name: (ctor kwargs, B, T, ragged, store_weights)
    "small_uni": (dict(neural_dim=16, n_classes=10, hidden_dim=32, layer_dim=2, nDays=3, dropout=0.0,
                       strideLen=2, kernelLen=8, gaussianSmoothWidth=2.0, bidirectional=False), 3, 40, True, True),
    "small_bi": (dict(neural_dim=16, n_classes=10, hidden_dim=32, layer_dim=3, nDays=4, dropout=0.0,
                      strideLen=4, kernelLen=14, gaussianSmoothWidth=1.5, bidirectional=True), 4, 61, True, True),
    # competition shape (256 feats, 24 days, 41 classes, 5x1024, k32/s4) at a tiny batch; weights are
    # regenerated from the seed by the tests, only outputs and gradient digests are stored.
    "comp_uni": (dict(neural_dim=256, n_classes=40, hidden_dim=1024, layer_dim=5, nDays=24, dropout=0.0,
                      strideLen=4, kernelLen=32, gaussianSmoothWidth=2.0, bidirectional=False), 2, 80, True, False),
    "comp_bi": (dict(neural_dim=256, n_classes=40, hidden_dim=1024, layer_dim=5, nDays=24, dropout=0.0,
                     strideLen=4, kernelLen=32, gaussianSmoothWidth=2.0, bidirectional=True), 2, 80, True, False),
}


def run(name, kw, B, T, ragged, store_weights, dtype):
    torch.manual_seed(0)
    model = RefGRUDecoder(device="cpu", **kw)
    fill_trained_like_(model, seed=7)
    if dtype == torch.float64:
        model = model.double()
        torch.set_default_dtype(torch.float64)       # h0 is created with the default dtype (model.py:105-117)
    model.eval()
    X, y, X_len, y_len, day = make_batch(B, T, n_feat=kw["neural_dim"], n_days=kw["nDays"],
                                         n_classes=kw["n_classes"], seed=11, ragged=ragged,
                                         min_tgt=2, max_tgt=12, kernel_len=kw["kernelLen"],
                                         stride_len=kw["strideLen"])
    Xd = X.to(dtype)

    stages = {}
    hooks = [
        model.gaussianSmoother.register_forward_hook(lambda m, i, o: stages.__setitem__("smoothed_bnt", o.detach())),
        model.inputLayerNonlinearity.register_forward_hook(lambda m, i, o: stages.__setitem__("z", o.detach())),
        model.unfolder.register_forward_hook(lambda m, i, o: stages.__setitem__("unfold_raw", o.detach())),
        model.gru_decoder.register_forward_hook(lambda m, i, o: stages.__setitem__("hid", o[0].detach())),
    ]
    pred = model.forward(Xd, day)                                                       # trainer:208
    for h in hooks:
        h.remove()
    pred.retain_grad()
    out_lens = ((X_len - model.kernelLen) / model.strideLen).to(torch.int32)           # trainer:209
    log_probs = pred.log_softmax(2).permute(1, 0, 2)                                    # trainer:210
    log_probs.retain_grad()
    loss_ctc = torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)          # trainer:141
    loss = loss_ctc(log_probs, y, out_lens, y_len)                                      # trainer:213-218
    loss = torch.sum(loss)                                                              # trainer:242
    loss.backward()                                                                     # trainer:252
    nll = torch.nn.CTCLoss(blank=0, reduction="none", zero_infinity=True)(
        log_probs.detach(), y, out_lens, y_len)

    decoded = []
    with torch.no_grad():                                                               # trainer:313-320
        lp = log_probs.detach()
        for i in range(lp.shape[1]):
            d = torch.argmax(lp[0:out_lens[i], i, :], dim=-1)
            d = torch.unique_consecutive(d, dim=-1)
            d = d.cpu().numpy()
            decoded.append(np.array([v for v in d if v != 0], dtype=np.int64))
    dec_pad = np.zeros((B, max(1, max(len(d) for d in decoded))), dtype=np.int64)
    dec_len = np.zeros(B, dtype=np.int64)
    for i, d in enumerate(decoded):
        dec_pad[i, :len(d)] = d
        dec_len[i] = len(d)

    f = lambda t: t.detach().cpu().numpy()
    out = {
        "X": f(X), "y": f(y), "X_len": f(X_len), "y_len": f(y_len), "dayIdx": f(day),
        "out_lens": f(out_lens), "logits": f(pred), "log_probs_tbc": f(log_probs),
        "loss": np.array(loss.item()), "nll": f(nll),
        "dlogits": f(pred.grad), "dlog_probs_tbc": f(log_probs.grad),
        "decoded": dec_pad, "decoded_len": dec_len,
        "smoothed": f(stages["smoothed_bnt"].permute(0, 2, 1)),      # back to [B,T,N]
        "z": f(stages["z"]),
        "patches": f(stages["unfold_raw"].permute(0, 2, 1)),         # [B,T',N*K]   (model.py:96-101)
        "hid": f(stages["hid"]),
        "taps": f(model.gaussianSmoother.weight[0, 0]),
    }
    live = {k: p for k, p in model.named_parameters() if p.grad is not None}
    if store_weights:
        for k, v in model.state_dict().items():
            if not k.startswith("inpLayer"):
                out["sd." + k] = f(v)
        for k, p in live.items():
            out["grad." + k] = f(p.grad)
    else:
        # digests: a strided sample plus sum and sum of squares of every gradient
        for k, p in live.items():
            g = p.grad.detach().reshape(-1)
            step = max(1, g.numel() // 4096)
            out["gsample." + k] = f(g[::step][:4096])
            out["gsum." + k] = np.array([g.double().sum().item(), (g.double() ** 2).sum().item()])
        out["hid"] = out["hid"][:, :, ::16]                          # keep the file small
        del out["patches"]
    out["live_grad_names"] = np.array(sorted(live.keys()))
    torch.set_default_dtype(torch.float32)
    return out


def main():
    for name, (kw, B, T, ragged, sw) in CONFIGS.items():
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            if not sw and tag == "f64":
                continue
            out = run(name, kw, B, T, ragged, sw, dtype)
            out["ctor"] = np.array(repr(kw))
            path = os.path.join(HERE, f"{name}_{tag}.npz")
            np.savez_compressed(path, **out)
            print(path, os.path.getsize(path) // 1024, "KiB", "loss", float(out["loss"]))


if __name__ == "__main__":
    main()
