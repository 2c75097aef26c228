"""Debug helper (not a test): NSD_GRU_TRACE=1 python tests/trace_gru.py -> per-step event timing of the K3 kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_speech_decoder_b200 import ops
B, Tp, H, D = 64, 20, 1024, 2
M = Tp * B
torch.manual_seed(0)
gi = torch.randn(M, D * 3 * H, device="cuda")
w = (torch.randn(D * 3 * H, H, device="cuda") / 32).to(torch.bfloat16)
b = torch.zeros(D * 3 * H, device="cuda")
for _ in range(2):
    hseq, hbf, sv = ops.gru_fwd_bf16(gi, w, b, Tp, B, H, D, False, True)
    torch.cuda.synchronize()
wT = torch.cat([w[d * 3 * H:(d + 1) * 3 * H].T.contiguous() for d in range(D)], 0)
dh = torch.randn(M, D * H, device="cuda")
for _ in range(2):
    ops.gru_bwd_bf16(dh, hseq, sv, wT, Tp, B, H, D, False)
    torch.cuda.synchronize()
