"""Debug helper: CUDA-event time of the K2 GEMM shapes of the step (fp32 vs bf16 output)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neural_speech_decoder_b200 import ops
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
M = 7552
for (N, K, ta, tb, name) in [(6144, 2048, False, True, "fwd l1-4"), (6144, 8192, False, True, "fwd l0"), (2048, 6144, False, False, "dgrad l1-4"),
                             (8192, 6144, False, False, "dgrad l0")]:
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    Bm = (torch.randn(N, K, device="cuda") if tb else torch.randn(K, N, device="cuda")).to(torch.bfloat16)
    for dt in (torch.float32, torch.bfloat16):
        C = torch.empty(M, N, device="cuda", dtype=dt)
        ms = t(lambda: ops.gemm(ta, tb, M, N, K, A, K, Bm, Bm.shape[1], C, N))
        print(f"{name:12s} M={M} N={N} K={K} out={str(dt)[6:]:8s} {ms*1e3:7.1f} us  {2*M*N*K/ms/1e9:7.1f} TFLOP/s")
# wgrad shapes (reduction over the 7552 rows)
for (Mo, No, name) in [(6144, 2048, "wgrad ih l1-4"), (3072, 1024, "wgrad hh")]:
    A = torch.randn(M, Mo, device="cuda").to(torch.bfloat16)
    Bm = torch.randn(M, No, device="cuda").to(torch.bfloat16)
    C = torch.empty(Mo, No, device="cuda")
    ms = t(lambda: ops.gemm(True, False, Mo, No, M, A, Mo, Bm, No, C, No))
    print(f"{name:12s} M={Mo} N={No} K={M} out=float32  {ms*1e3:7.1f} us  {2*M*Mo*No/ms/1e9:7.1f} TFLOP/s")
