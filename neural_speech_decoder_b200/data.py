"""Host -> device staging of training batches (the reference's hot loop does five synchronous ``.to(device)`` copies
per step, neural_decoder_trainer.py:185-191).  ``BatchPrefetcher`` keeps one batch in flight on a copy stream so the
PCIe transfer of step i+1 runs under the kernels of step i; the consumer's stream waits on the copy's event only."""
from __future__ import annotations

from typing import Iterable, Iterator, Sequence, Tuple

import torch


class BatchPrefetcher:
    """Iterate device batches from an iterable of host batches (tuples of tensors, ideally pinned).

        for X, y, X_len, y_len, dayIdx in BatchPrefetcher(loader, device):
            loss = train_step(model, opt, X, y, X_len, y_len, dayIdx)
    """

    def __init__(self, batches: Iterable[Sequence[torch.Tensor]], device, depth: int = 1):
        self.it: Iterator = iter(batches)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("BatchPrefetcher stages batches onto a CUDA device")
        self.copy_stream = torch.cuda.Stream(self.device)
        self.depth = max(1, depth)
        self.queue = []

    def _issue(self) -> bool:
        try:
            host = next(self.it)
        except StopIteration:
            return False
        with torch.cuda.stream(self.copy_stream):
            dev = tuple(t.to(self.device, non_blocking=True) for t in host)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.queue.append((dev, ev))
        return True

    def __iter__(self):
        return self

    def __next__(self) -> Tuple[torch.Tensor, ...]:
        while len(self.queue) <= self.depth and self._issue():
            pass
        if not self.queue:
            raise StopIteration
        dev, ev = self.queue.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev:
            t.record_stream(cur)          # allocated on the copy stream, consumed on the compute stream
        return dev
