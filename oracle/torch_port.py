"""CPU baseline port of the reference hot path, built from the SAME third-party torch operators the reference
calls.  TEST / BENCH INFRASTRUCTURE ONLY -- the product package never imports this.

The reference (EdwardoSunny/Neural-Speech-Decoder) is pure Python; its arithmetic is torch's
(``F.conv1d``, ``torch.einsum``, ``nn.Softsign``, ``nn.Unfold``, ``nn.GRU``, ``nn.Linear``, ``log_softmax``,
``nn.CTCLoss``, ``torch.optim.Adam``).  /root/reference does not exist on the GPU box and a Python reference
cannot travel, so ``bench.py``'s ``cpu_baseline`` leg and ``--impl reference`` arm time this restatement
("kind": "port") on the box's host cores.  It follows, stage by stage:
    src/neural_decoder/augmentations.py:41-69, 83-91     Gaussian taps, depthwise conv1d padding="same"
    src/neural_decoder/model.py:36-81                    module construction and initialisation
    src/neural_decoder/model.py:83-123                   forward
    src/neural_decoder/neural_decoder_trainer.py:139-141, 163-169, 208-218, 242, 251-259   loss / Adam step
Pinned: tests/test_oracle_golden.py::test_torch_port_matches_reference_fixture checks its logits, loss and
gradients against tests/golden/*.npz (outputs of the reference itself).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn


class PortGRUDecoder(nn.Module):
    def __init__(self, neural_dim, n_classes, hidden_dim, layer_dim, nDays=24, dropout=0, device="cpu",
                 strideLen=4, kernelLen=14, gaussianSmoothWidth=0, bidirectional=False):
        super().__init__()
        self.kernelLen, self.strideLen, self.neural_dim = kernelLen, strideLen, neural_dim
        self.layer_dim, self.hidden_dim, self.bidirectional = layer_dim, hidden_dim, bidirectional
        # augmentations.py:50-69
        grid = torch.arange(20, dtype=torch.float32)
        mean = (20 - 1) / 2
        k = 1 / (gaussianSmoothWidth * math.sqrt(2 * math.pi)) * torch.exp(-(((grid - mean) / gaussianSmoothWidth) ** 2) / 2)
        k = k / torch.sum(k)
        self.register_buffer("smooth_w", k.view(1, 1, -1).repeat(neural_dim, 1, 1))
        # model.py:43-47
        self.dayWeights = nn.Parameter(torch.eye(neural_dim).repeat(nDays, 1, 1))
        self.dayBias = nn.Parameter(torch.zeros(nDays, 1, neural_dim))
        # model.py:50-63
        self.gru = nn.GRU(neural_dim * kernelLen, hidden_dim, layer_dim, batch_first=True, dropout=dropout,
                          bidirectional=bidirectional)
        for name, p in self.gru.named_parameters():
            if "weight_hh" in name:
                nn.init.orthogonal_(p)
            if "weight_ih" in name:
                nn.init.xavier_uniform_(p)
        # model.py:76-81
        self.fc = nn.Linear(hidden_dim * (2 if bidirectional else 1), n_classes + 1)

    def load_reference_state(self, sd):
        """Accepts a GRUDecoder state dict (reference key names)."""
        own = self.state_dict()
        with torch.no_grad():
            for k, v in sd.items():
                v = torch.as_tensor(v)
                if k.startswith("gru_decoder."):
                    own["gru." + k[len("gru_decoder."):]].copy_(v)
                elif k.startswith("fc_decoder_out."):
                    own["fc." + k[len("fc_decoder_out."):]].copy_(v)
                elif k in ("dayWeights", "dayBias"):
                    own[k].copy_(v)
                elif k == "gaussianSmoother.weight":
                    own["smooth_w"].copy_(v)

    def forward(self, x, dayIdx):
        x = x.permute(0, 2, 1)                                               # model.py:84
        x = F.conv1d(x, self.smooth_w, groups=self.neural_dim, padding="same")   # augmentations.py:91
        x = x.permute(0, 2, 1)                                               # model.py:86
        w = torch.index_select(self.dayWeights, 0, dayIdx)                   # model.py:89
        x = torch.einsum("btd,bdk->btk", x, w) + torch.index_select(self.dayBias, 0, dayIdx)   # model.py:90-92
        x = F.softsign(x)                                                    # model.py:93
        x = F.unfold(x.permute(0, 2, 1).unsqueeze(3), (self.kernelLen, 1), stride=self.strideLen).permute(0, 2, 1)  # :96-101
        D = 2 if self.bidirectional else 1
        h0 = torch.zeros(self.layer_dim * D, x.size(0), self.hidden_dim, dtype=x.dtype, device=x.device)   # model.py:104-117
        hid, _ = self.gru(x, h0)                                             # model.py:119
        return self.fc(hid)                                                  # model.py:122


def make_adam(model, lr=0.02, l2=1e-5):
    return torch.optim.Adam(model.parameters(), lr=lr, betas=(0.9, 0.999), eps=0.1, weight_decay=l2)   # trainer:163-169


def train_step(model, opt, X, y, X_len, y_len, dayIdx, white_noise_sd: float = 0.0, constant_offset_sd: float = 0.0):
    if white_noise_sd > 0:                                                           # trainer:194-196
        X = X + torch.randn(X.shape, device=X.device) * white_noise_sd
    if constant_offset_sd > 0:                                                       # trainer:198-201
        X = X + torch.randn([X.shape[0], 1, X.shape[2]], device=X.device) * constant_offset_sd
    pred = model(X, dayIdx)                                                          # trainer:208
    out_lens = ((X_len - model.kernelLen) / model.strideLen).to(torch.int32)         # trainer:209
    log_probs = pred.log_softmax(2).permute(1, 0, 2)                                 # trainer:210
    loss = nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)(log_probs, y, out_lens, y_len)   # trainer:141, 213-218
    loss = torch.sum(loss)                                                           # trainer:242
    opt.zero_grad()                                                                  # trainer:251
    loss.backward()                                                                  # trainer:252
    opt.step()                                                                       # trainer:259
    return loss
