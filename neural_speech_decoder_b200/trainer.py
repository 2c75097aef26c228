"""Hot-loop lines of the reference trainer, on the B200 kernels.

Only what sits on the north-star path is mirrored here (neural_decoder_trainer.py):
  make_optimizer  :163-175   Adam(lr, betas .9/.999, eps=0.1, weight_decay=l2) + LinearLR
  train_step      :208-218, 242, 251-260   forward -> out_lens -> log-softmax + CTC -> backward -> step
  eval_batch      :299-333   forward -> CTC loss -> greedy decode -> edit distance
Data loading, wandb, checkpointing stay the reference's business; its ``trainModel`` runs unchanged with
``GRUDecoder`` / ``CTCLoss`` swapped in.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ctc as _ctc
from .adam import FusedAdam


def _step_in_backward() -> bool:
    import os
    return os.environ.get("NSD_STEP_IN_BACKWARD", "1") != "0"


def make_optimizer(model: torch.nn.Module, args: dict):
    """trainer:163-175 (the non-AdamW branch, which the GRU configs use)."""
    opt = FusedAdam(model.parameters(), lr=args["lrStart"], betas=(0.9, 0.999), eps=0.1,
                    weight_decay=args.get("l2_decay", 0.0))
    if getattr(model, "_shadows", None) is not None:
        opt.attach_shadows(model._shadows)
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=args["lrEnd"] / args["lrStart"],
                                              total_iters=args["nBatch"])
    return opt, sched


class LossReader:
    """Host read-back of a step's loss that does not drain the backward behind it.  ``loss.item()`` on the compute stream waits for the
    whole step (the stream is in order) and the GPU then idles until the host has enqueued the next step's first kernels; the loss itself
    is final after the forward + CTC kernel, ~40 % into the step.  ``train_step(..., loss_reader=r)`` copies it to a pinned float on a side
    stream as soon as it exists; ``r.item()`` waits for that copy only, so the host enqueues step i+1 while step i's backward still runs."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.buf = torch.zeros(1, dtype=torch.float32).pin_memory()
        self.stream = torch.cuda.Stream(self.device)
        self.done = torch.cuda.Event()
        self._ready = torch.cuda.Event()

    def capture(self, loss: torch.Tensor) -> None:
        self._ready.record(torch.cuda.current_stream(self.device))
        self.stream.wait_event(self._ready)
        with torch.cuda.stream(self.stream):
            self.buf.copy_(loss.detach().reshape(1), non_blocking=True)
            self.done.record(self.stream)
        loss.record_stream(self.stream)

    def item(self) -> float:
        self.done.synchronize()
        return float(self.buf[0])


def train_step(model, optimizer, X, y, X_len, y_len, dayIdx, scheduler=None, grad_sync=None,
               white_noise_sd: float = 0.0, constant_offset_sd: float = 0.0, loss_reader: Optional[LossReader] = None) -> torch.Tensor:
    """One optimisation step; returns the (device) loss scalar.  ``grad_sync`` is the data-parallel hook
    (parallel.GradSync): it all-reduces gradients bucket by bucket while the backward is still running.
    ``white_noise_sd`` / ``constant_offset_sd`` (args["whiteNoiseSD"], args["constantOffsetSD"]): the in-loop
    augmentation of trainer:194-201, generated inside the front-end kernel instead of in extra passes over X.
    ``loss_reader`` (LossReader): host read-back of the loss that does not wait for the backward."""
    model.grad_sync = grad_sync
    # the buckets are all-reduced with SUM: the 1/world average is folded into the Adam kernel.  Adam with eps=0.1 and
    # L2-in-gradient is NOT scale-invariant, so the scale must follow the sync object on every step.
    want_scale = grad_sync.grad_scale if grad_sync is not None else 1.0
    if isinstance(optimizer, FusedAdam):
        optimizer.grad_scale = want_scale
    elif grad_sync is not None and grad_sync.world > 1:
        raise RuntimeError("train_step(grad_sync=...) sums gradients over ranks; use FusedAdam (make_optimizer), which folds the "
                           "1/world_size average into its update, or divide the gradients yourself")
    model.input_noise = (white_noise_sd, constant_offset_sd) if (white_noise_sd or constant_offset_sd) else None
    # bf16 path: the update of each finished gradient bucket is issued from inside the backward, directly behind a BPTT launch, and runs
    # under it on the SMs the recurrence leaves free (model_tc.decoder_backward_tc): behind the next layer's launch on one GPU, one layer
    # later (after the bucket's all-reduce) when data parallel.  What has no recurrence left to hide under (layer 0, the day weights; data
    # parallel also layer 1) is updated after the backward as usual.  Same arithmetic per parameter.
    stepped = []
    model.step_hook = None
    if (isinstance(optimizer, FusedAdam) and getattr(model, "precision", None) == "bf16" and _step_in_backward()
            and len(optimizer.param_groups) == 1):
        def _hook(ps):
            optimizer.step(only=ps, under_recurrence=True)
            stepped.extend(ps)
        model.step_hook = _hook
    pred = model.forward(X, dayIdx)                                         # trainer:208
    lens = _ctc.out_lens(X_len, model.kernelLen, model.strideLen)           # trainer:209
    loss = _ctc.ctc_loss_from_logits(pred, y, lens, y_len, blank=0, reduction="mean")   # trainer:210-218, 242
    if loss_reader is not None:
        loss_reader.capture(loss)
    optimizer.zero_grad(set_to_none=True)                                   # trainer:251
    if grad_sync is not None:
        grad_sync.begin()
    try:
        loss.backward()                                                     # trainer:252
    finally:
        model.step_hook = None
    if grad_sync is not None and grad_sync.world > 1 and isinstance(optimizer, FusedAdam):
        # the last big bucket (layer 0) is still being all-reduced when the backward's kernels are done: update the parameters of
        # every finished bucket under it, then the rest (same arithmetic per parameter; the step is just issued in two launches)
        pending = grad_sync.finish_early()
        done = {id(p) for p in stepped}                       # buckets already updated from inside the backward
        if pending:
            live = [p for g in optimizer.param_groups for p in g["params"] if p.grad is not None and id(p) not in done]
            late = [p for p in live if p.grad.untyped_storage().data_ptr() in pending]
            late_ids = {id(p) for p in late}
            optimizer.step(only=[p for p in live if id(p) not in late_ids])
            grad_sync.finish()
            optimizer.step(only=late)
        else:
            grad_sync.finish()
            optimizer.step(only=[p for g in optimizer.param_groups for p in g["params"] if p.grad is not None and id(p) not in done])
    elif stepped:
        done = {id(p) for p in stepped}
        optimizer.step(only=[p for g in optimizer.param_groups for p in g["params"] if p.grad is not None and id(p) not in done])
    else:
        if grad_sync is not None:
            grad_sync.finish()
        optimizer.step()                                                    # trainer:259
    if scheduler is not None:
        scheduler.step()                                                    # trainer:260
    return loss.detach()


@torch.no_grad()
def eval_batch(model, X, y, X_len, y_len, dayIdx) -> Tuple[torch.Tensor, int, int]:
    """trainer:299-333 for one batch: (ctc loss, sum edit distance, sum true length)."""
    logits = model.forward(X, dayIdx)                                       # trainer:299
    lens = _ctc.out_lens(X_len, model.kernelLen, model.strideLen)           # trainer:300
    pred = _ctc.log_softmax_tbc(logits)                                     # trainer:301  [T',B,C] view
    loss = _ctc.CTCLoss(blank=0, reduction="mean", zero_infinity=True)(pred, y, lens, y_len)   # trainer:303-309
    # decode from the log-probs, as the reference does: log-softmax can create ties the raw logits lack
    dist, tot = _ctc.phoneme_error_rate(pred, lens, y, y_len)               # trainer:313-333
    return loss, dist, tot
