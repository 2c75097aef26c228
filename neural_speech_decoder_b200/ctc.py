"""CTC loss, greedy decode and phoneme error rate on the sm_100a kernels (K4 / K5).

``CTCLoss`` mirrors ``torch.nn.CTCLoss`` as the reference trainer uses it
(neural_decoder_trainer.py:139-141, 213-218): call signature
``(log_probs[T',B,C], targets i32[B,S], input_lengths[B], target_lengths[B])``,
``blank``, ``reduction`` in {"mean","sum","none"} and ``zero_infinity``.  The
log-probs may be any strided view (the trainer passes ``.permute(1,0,2)`` of a
``[B,T',C]`` tensor); lengths and targets stay on the device, there is no host
synchronisation.

``ctc_loss_from_logits`` is the fused form used by ``train_step``: log-softmax,
alpha/beta, loss and d loss / d logits in one launch (trainer:210-218, 242, 252).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
from torch import nn

from . import ops
from ._lib import NsdError, require_cuda


def _as_i32(t: torch.Tensor, dev) -> torch.Tensor:
    return t.to(device=dev, dtype=torch.int32).contiguous()


def _prep_targets(targets: torch.Tensor, dev) -> torch.Tensor:
    if targets.dim() != 2:
        raise NsdError("CTCLoss (B200): targets must be 2-D padded [B,S] as the trainer's collate produces them")
    t = _as_i32(targets, dev)
    if t.shape[1] == 0:
        t = torch.zeros((t.shape[0], 1), device=dev, dtype=torch.int32)
    return t


class _CtcFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, act, targets, in_lens, tgt_lens, blank, reduction, zero_infinity, is_logits, tbc):
        """act: log-probs [T,B,C] (any strides) when tbc else logits [B,T,C]."""
        require_cuda(act, "log_probs")
        if not zero_infinity:
            raise NsdError("CTCLoss (B200): only zero_infinity=True (the reference's setting) is implemented")
        if act.dtype != torch.float32:
            act = act.float()
        if torch.empty_like(act).stride() != act.stride():      # the gradient is written with act's own strides
            act = act.contiguous()
        if tbc:
            T, B, C = act.shape
            st, sb, sc = act.stride()
        else:
            B, T, C = act.shape
            sb, st, sc = act.stride()
        want_grad = bool(ctx.needs_input_grad[0])
        mean = reduction == "mean"
        loss, nll, grad = ops.ctc_loss_raw(act, st, sb, sc, is_logits, targets, in_lens, tgt_lens, T, B, C, blank,
                                           mean, want_grad)
        ctx.reduction, ctx.tbc = reduction, tbc
        ctx.save_for_backward(grad)
        if reduction == "none":
            return nll
        return loss

    @staticmethod
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        if grad is None:
            return (None,) * 9
        if ctx.reduction == "none":
            # grad holds d nll[b] / d act; scale each utterance by its upstream gradient
            g = grad * (gout.view(1, -1, 1) if ctx.tbc else gout.view(-1, 1, 1))
        else:
            g = grad * gout
        return (g,) + (None,) * 8


def _ctc_apply(act, targets, in_lens, tgt_lens, blank, reduction, zero_infinity, is_logits, tbc):
    require_cuda(act, "log_probs")
    dev = act.device
    targets = _prep_targets(targets, dev)
    in_lens = _as_i32(in_lens, dev)
    tgt_lens = _as_i32(tgt_lens, dev)
    if reduction not in ("mean", "sum", "none"):
        raise ValueError(reduction)
    with torch.cuda.device(dev):            # raw-pointer launches must target the tensors' device, not the caller's current one
        return _CtcFunction.apply(act, targets, in_lens, tgt_lens, blank, reduction, zero_infinity, is_logits, tbc)


class CTCLoss(nn.Module):
    """Drop-in for ``torch.nn.CTCLoss`` on CUDA tensors (trainer:139-141)."""

    def __init__(self, blank: int = 0, reduction: str = "mean", zero_infinity: bool = False):
        super().__init__()
        self.blank, self.reduction, self.zero_infinity = blank, reduction, zero_infinity

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        return _ctc_apply(log_probs, targets, input_lengths, target_lengths, self.blank, self.reduction,
                          self.zero_infinity, False, True)


def ctc_loss_from_logits(logits, targets, input_lengths, target_lengths, blank=0, reduction="mean"):
    """log_softmax(2) + CTCLoss(blank, reduction, zero_infinity=True) on logits [B,T',C]; one launch, and the
    backward is the already-computed d loss / d logits (trainer:210, 213-218, 252)."""
    return _ctc_apply(logits, targets, input_lengths, target_lengths, blank, reduction, True, True, False)


def log_softmax_tbc(logits: torch.Tensor) -> torch.Tensor:
    """``logits.log_softmax(2).permute(1, 0, 2)`` (trainer:210, 301): logits [B,T',C] -> log-probs as a
    [T',B,C] view of a contiguous [B,T',C] buffer, exactly the layout the reference hands to CTCLoss."""
    require_cuda(logits, "logits")
    with torch.cuda.device(logits.device):
        return ops.log_softmax(logits.float().contiguous()).permute(1, 0, 2)


def out_lens(X_len: torch.Tensor, kernel_len: int, stride_len: int) -> torch.Tensor:
    """((X_len - kernelLen) / strideLen).to(int32)  -- trainer:209, 300 (true division, truncation)."""
    return ((X_len - kernel_len) / stride_len).to(torch.int32)


def greedy_decode(log_probs: torch.Tensor, lens: torch.Tensor, blank: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """argmax -> unique_consecutive -> drop blank for every utterance of ``log_probs`` [T',B,C] (any strides)
    in ONE launch (trainer:313-320 does it per utterance with a device->host sync each).
    Returns (decoded int64 [B,T'], decoded_len int32 [B]) on the device."""
    require_cuda(log_probs, "log_probs")
    if log_probs.dtype != torch.float32:
        log_probs = log_probs.float()
    T, B, C = log_probs.shape
    st, sb, sc = log_probs.stride()
    with torch.cuda.device(log_probs.device):
        return ops.greedy_decode_raw(log_probs, st, sb, sc, _as_i32(lens, log_probs.device), T, B, C, blank)


def decoded_to_lists(dec: torch.Tensor, dec_len: torch.Tensor) -> List[List[int]]:
    d, l = dec.cpu().numpy(), dec_len.cpu().numpy()
    return [d[i, :l[i]].tolist() for i in range(d.shape[0])]


def edit_distances(dec, dec_len, targets, target_lengths) -> torch.Tensor:
    """Levenshtein distance per utterance on the device (SequenceMatcher.distance(), trainer:322-330)."""
    dev = dec.device
    with torch.cuda.device(dev):
        return ops.edit_distance_raw(dec.contiguous(), _as_i32(dec_len, dev), _prep_targets(targets, dev),
                                     _as_i32(target_lengths, dev))


def phoneme_error_rate(log_probs, lens, targets, target_lengths, blank: int = 0) -> Tuple[int, int]:
    """(sum of edit distances, sum of true lengths) for a batch -- `cer` numerator/denominator (trainer:332-333).
    One device->host read for the whole batch."""
    dec, dec_len = greedy_decode(log_probs, lens, blank)
    dist = edit_distances(dec, dec_len, targets, target_lengths)
    both = torch.stack([dist.sum(), _as_i32(target_lengths, dist.device).sum()]).cpu()
    return int(both[0]), int(both[1])
