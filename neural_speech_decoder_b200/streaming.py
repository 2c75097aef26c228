"""Stateful streaming inference for the unidirectional GRUDecoder (BASELINE configs[3]).

The reference has no streaming API: ``forward`` always starts from h0 = 0 and drops the final state
(model.py:104-119).  ``StreamingDecoder`` feeds the SAME kernels incrementally and reproduces the offline logits
bit for bit: it carries the fp32 hidden state of every layer between calls and keeps the raw bins a future frame still
needs.  Frame j covers bins [4j, 4j+32) (kernel 32 / stride 4, model.py:37-39) of the smoothed signal, and the 20-tap
Gaussian is padded 9 left / 10 right (augmentations.py:91), so frame j can be emitted once bin 4j+41 has arrived: a
fixed look-ahead of 10 bins (200 ms).  ``finish()`` flushes the frames the offline model would still produce by zero
padding past the end of the utterance, exactly as the offline smoothing does.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import ops
from ._lib import NsdError


def complete_frames(n_bins: int, kernel_len: int, stride_len: int, lookahead: int) -> int:
    """Number of output frames that can be emitted once ``n_bins`` bins have arrived: frame j reads smoothed bins
    [S*j, S*j+K) and the smoothing reads ``lookahead`` raw bins past each of them."""
    if n_bins < kernel_len + lookahead:
        return 0
    return (n_bins - kernel_len - lookahead) // stride_len + 1


def window_for(j0: int, stride_len: int, halo: int):
    """(first raw bin, number of leading frames to discard) of the window that recomputes frames from j0 on with
    ``halo`` bins (a multiple of the stride, >= the smoothing's left reach) of true history in front."""
    r0 = max(0, stride_len * j0 - halo)
    return r0, (stride_len * j0 - r0) // stride_len


class StreamingDecoder:
    def __init__(self, model, batch_size: int, dayIdx: torch.Tensor):
        if model.bidirectional:
            raise NsdError("streaming needs the unidirectional GRUDecoder (the reverse direction reads the future)")
        if model.precision != "bf16":
            raise NsdError("StreamingDecoder runs the bf16 tensor-core path (set_default_precision('bf16'))")
        self.m = model
        self.B = int(batch_size)
        p = next(model.parameters())
        self.dev = p.device
        self.day = dayIdx.to(device=self.dev, dtype=torch.int64).contiguous()
        if self.day.shape != (self.B,):
            raise RuntimeError(f"dayIdx must be [{self.B}]")
        self.K, self.S, self.N = model.kernelLen, model.strideLen, model.neural_dim
        taps = model.gaussianSmoother.weight[0, 0]
        self.left = (taps.numel() - 1) // 2                     # 9
        self.right = taps.numel() - 1 - self.left               # 10
        self.halo = -(-self.left // self.S) * self.S            # left context re-read per call, a multiple of the stride (12)
        self.reset()

    def reset(self) -> None:
        m = self.m
        self.h: List[torch.Tensor] = [torch.zeros(self.B, m.hidden_dim, device=self.dev) for _ in range(m.layer_dim)]
        self.hist = torch.empty(self.B, 0, self.N, device=self.dev)
        self.hist_start = 0            # absolute bin index of hist[:, 0]
        self.n_bins = 0                # bins received so far
        self.next_frame = 0            # first frame not yet emitted

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _emit(self, j0: int, j1: int, r1: int) -> torch.Tensor:
        """Frames j0..j1 (inclusive) from bins [r0, r1): r0 re-reads ``halo`` bins of left context so that the frames
        we keep see the true smoothing history; frames computed from the window's padded edges are discarded."""
        m, B, K, S = self.m, self.B, self.K, self.S
        r0, skip = window_for(j0, S, self.halo)
        x = self.hist[:, r0 - self.hist_start:r1 - self.hist_start].contiguous()
        k = j1 - j0 + 1
        taps = m.gaussianSmoother.weight[0, 0].contiguous()
        patches, _, _ = ops.frontend_fwd(x, self.day, m.dayWeights.detach().contiguous(), m.dayBias.detach().contiguous(), taps,
                                         K, S, torch.bfloat16, m._err_flag if m._err_flag is not None else None)
        inp = patches[skip * B:(skip + k) * B]                  # time-major rows: a contiguous block of frames
        H, L = m.hidden_dim, m.layer_dim
        M = k * B
        gw = m._gru_weights()
        for l in range(L):
            w_ih, w_hh, b_ih, b_hh = gw[4 * l:4 * l + 4]
            w_ih_bf = m._shadows.stacked(("ih", l), [w_ih])
            w_hh_bf = m._shadows.stacked(("hh", l), [w_hh])
            gi = torch.empty((M, 3 * H), device=self.dev, dtype=torch.float32)
            in_l = inp.shape[1]
            ops.gemm(False, True, M, 3 * H, in_l, inp, in_l, w_ih_bf, in_l, gi, 3 * H, bias=b_ih.detach().contiguous())
            hseq, hseq_bf, _ = ops.gru_fwd_bf16(gi, w_hh_bf, b_hh.detach().contiguous(), k, B, H, 1, False, False, h0=self.h[l])
            self.h[l] = hseq[(k - 1) * B:].clone()              # fp32 state carried to the next call
            inp = hseq_bf
        fc = m.fc_decoder_out
        C = fc.weight.shape[0]
        fc_bf = m._shadows.stacked(("fc", 0), [fc.weight])
        logits_tm = torch.empty((M, C), device=self.dev, dtype=torch.float32)
        ops.gemm(False, True, M, C, H, inp, H, fc_bf, H, logits_tm, C, bias=fc.bias.detach().contiguous())
        return ops.swap01(logits_tm.view(k, B, C))

    def _trim(self) -> None:
        keep_from = max(0, self.S * self.next_frame - self.halo)
        if keep_from > self.hist_start:
            self.hist = self.hist[:, keep_from - self.hist_start:].contiguous()
            self.hist_start = keep_from

    @torch.no_grad()
    def push(self, bins: torch.Tensor) -> Optional[torch.Tensor]:
        """bins [B, n, N] (n >= 0 new 20 ms bins) -> logits [B, f, C] of the f frames that became complete (None if f == 0)."""
        if bins.dim() != 3 or bins.shape[0] != self.B or bins.shape[2] != self.N:
            raise RuntimeError(f"bins must be [{self.B}, n, {self.N}], got {tuple(bins.shape)}")
        if self.m._err_flag is None:
            self.m._err_flag = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.hist = torch.cat([self.hist, bins.to(self.dev, torch.float32)], dim=1)
        self.n_bins += bins.shape[1]
        # frame j needs smoothed bins up to S*j+K-1, i.e. raw bins up to S*j+K-1+right
        j1 = complete_frames(self.n_bins, self.K, self.S, self.right) - 1
        if j1 < self.next_frame:
            return None
        with torch.cuda.device(self.dev):
            out = self._emit(self.next_frame, j1, self.S * j1 + self.K + self.right)
        self.next_frame = j1 + 1
        self._trim()
        return out

    @torch.no_grad()
    def finish(self) -> Optional[torch.Tensor]:
        """End of utterance: the frames the offline forward still produces (it zero-pads the smoothing past the last bin)."""
        if self.n_bins < self.K:
            if self.next_frame == 0 and self.n_bins > 0:
                raise RuntimeError(f"utterance shorter than kernelLen={self.K} bins")     # the reference raises too (nn.Unfold)
            return None
        j1 = (self.n_bins - self.K) // self.S
        if j1 < self.next_frame:
            return None
        with torch.cuda.device(self.dev):
            out = self._emit(self.next_frame, j1, self.n_bins)
        self.next_frame = j1 + 1
        self._trim()
        return out
