"""Debug helper: CUDA-event time of the K3 kernels at the benchmark shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_speech_decoder_b200 import ops
B, Tp, H, D = int(os.environ.get("B", 64)), int(os.environ.get("TP", 118)), 1024, 2
M = Tp * B
torch.manual_seed(0)
gi = torch.randn(M, D * 3 * H, device="cuda")
w = (torch.randn(D * 3 * H, H, device="cuda") / 32).to(torch.bfloat16)
b = torch.zeros(D * 3 * H, device="cuda")
wT = torch.cat([w[d * 3 * H:(d + 1) * 3 * H].T.contiguous() for d in range(D)], 0)
dh = torch.randn(M, D * H, device="cuda")
hseq, hbf, sv = ops.gru_fwd_bf16(gi, w, b, Tp, B, H, D, False, True)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
tf = t(lambda: ops.gru_fwd_bf16(gi, w, b, Tp, B, H, D, False, True))
tb = t(lambda: ops.gru_bwd_bf16(dh, hseq, sv, wT, Tp, B, H, D, False))
print(f"B={B} Tp={Tp}: fwd {tf*1e3:.1f} us ({tf*1e3/Tp:.2f} us/step)  bwd {tb*1e3:.1f} us ({tb*1e3/Tp:.2f} us/step)")
