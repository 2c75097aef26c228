"""Synthetic inputs and machine-independent parameter fills.

The benchmark and the parity fixtures need (a) batches of the competition shape
(256 features, 20 ms bins, 24 days, 41 classes) and (b)
weights that are bit-identical on every machine without shipping 135 M floats.
Both come from numpy's PCG64 streams, which are reproducible across platforms.
Nothing here touches the GPU.
"""
from __future__ import annotations

import zlib
from typing import Dict

import numpy as np
import torch


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.default_rng([seed, zlib.crc32(name.encode())])


def trained_like_state(shapes: Dict[str, tuple], seed: int = 0) -> Dict[str, np.ndarray]:
    """Deterministic float32 values for every GRUDecoder tensor in ``shapes``
    (name -> shape), scaled so that activations and logits look like a trained
    model's rather than the near-uniform logits of a fresh init (SURVEY.md
    section 7, "hard parts").  The Gaussian taps buffer is left to the caller."""
    out = {}
    for name, shape in shapes.items():
        g = _rng(seed, name)
        n = g.standard_normal(shape, dtype=np.float32)
        if name == "dayWeights":
            eye = np.eye(shape[1], dtype=np.float32)[None]
            v = eye + np.float32(0.05) * n
        elif name == "dayBias":
            v = np.float32(0.1) * n
        elif "weight_ih" in name:
            v = n * np.float32(1.5 * np.sqrt(2.0 / (shape[0] / 3 + shape[1])))
        elif "weight_hh" in name:
            v = n * np.float32(1.0 / np.sqrt(shape[1]))
        elif name.startswith("gru_decoder.bias"):
            v = np.float32(0.1) * n
        elif name == "fc_decoder_out.weight":
            v = n * np.float32(3.0 / np.sqrt(shape[1]))
        elif name == "fc_decoder_out.bias":
            v = np.float32(0.1) * n
        elif name.startswith("inpLayer"):
            v = np.float32(0.01) * n          # dead parameters (model.py:66-73)
        else:
            continue
        out[name] = np.ascontiguousarray(v, dtype=np.float32)
    return out


@torch.no_grad()
def fill_trained_like_(module: torch.nn.Module, seed: int = 0) -> None:
    """In-place ``trained_like_state`` for any module with GRUDecoder's names."""
    sd = module.state_dict()
    vals = trained_like_state({k: tuple(v.shape) for k, v in sd.items()}, seed)
    for k, v in vals.items():
        sd[k].copy_(torch.from_numpy(v))


def make_batch(B: int, T: int, n_feat: int = 256, n_days: int = 24, n_classes: int = 40,
               seed: int = 1, ragged: bool = False, min_tgt: int = 10, max_tgt: int = 50,
               kernel_len: int = 32, stride_len: int = 4):
    """Synthetic batch in the trainer's collate format (trainer:26-37):
    X f32[B,T,N] (zero-padded past X_len), y i32[B,maxlen] zero-padded,
    X_len i32[B], y_len i32[B], dayIdx i64[B].  Labels are 1..n_classes
    (0 is blank/pad).  Target lengths are capped so every utterance is
    CTC-feasible for its number of output frames."""
    g = np.random.default_rng([seed, B, T])
    X = g.standard_normal((B, T, n_feat), dtype=np.float32)
    if ragged:
        x_len = g.integers(max(kernel_len + stride_len, T // 2), T + 1, size=B).astype(np.int32)
        x_len[0] = T
        for b in range(B):
            X[b, x_len[b]:] = 0.0
    else:
        x_len = np.full(B, T, dtype=np.int32)
    frames = np.trunc((x_len.astype(np.float32) - kernel_len) / stride_len).astype(np.int64)
    hi = np.maximum(1, np.minimum(max_tgt, frames // 2))
    lo = np.minimum(min_tgt, hi)
    y_len = (lo + (g.random(B) * (hi - lo + 1)).astype(np.int64)).clip(lo, hi).astype(np.int32)
    y = np.zeros((B, int(y_len.max())), dtype=np.int32)
    for b in range(B):
        y[b, :y_len[b]] = g.integers(1, n_classes + 1, size=int(y_len[b]))
    day = g.integers(0, n_days, size=B).astype(np.int64)
    return (torch.from_numpy(X), torch.from_numpy(y), torch.from_numpy(x_len),
            torch.from_numpy(y_len), torch.from_numpy(day))
