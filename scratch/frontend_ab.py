"""Debug helper: the fast front-end forward (frontend_fwd_fast_kernel) against the generic kernel it replaces at the benchmark shape:
bit-identical outputs (patches, ys, z) over a set of shapes, and CUDA-event times.  Run: python scratch/frontend_ab.py"""
import os, subprocess, sys, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CASES = [(64, 500, 32, 4, (0.8, 0.2, 7)), (64, 500, 32, 4, None), (3, 97, 32, 4, (0.5, 0.0, 3)), (5, 213, 16, 8, (0.0, 0.3, 1)), (2, 64, 64, 4, None),
         (7, 333, 8, 4, (0.8, 0.2, 5)), (256, 2000, 32, 4, (0.8, 0.2, 9))]
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from neural_speech_decoder_b200 import ops
    for B, T, K, S, noise in CASES:
        torch.manual_seed(B * 1000 + T)
        x = torch.randn(B, T, 256, device="cuda")
        day = torch.randint(0, 4, (B,), device="cuda")
        W = torch.randn(4, 256, 256, device="cuda") / 16
        bb = torch.randn(4, 256, device="cuda") * 0.1
        g = torch.arange(-9.5, 10.5, device="cuda")
        taps = torch.exp(-0.5 * (g / 2) ** 2); taps = (taps / taps.sum()).contiguous()
        f = lambda: ops.frontend_fwd(x, day, W, bb, taps, K, S, torch.bfloat16, None, noise)
        pch, ys, z = f()
        torch.cuda.synchronize()
        h = hashlib.sha1(pch.view(torch.int16).cpu().numpy().tobytes() + ys.cpu().numpy().tobytes() + z.cpu().numpy().tobytes()).hexdigest()[:16]
        for _ in range(3): f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): f()
        e1.record(); torch.cuda.synchronize()
        print(f"B={B} T={T} K={K} S={S} noise={noise}: sha1 {h}  {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
else:
    outs = {}
    for tag, env in (("fast", "0"), ("generic", "1")):
        r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, NSD_FRONTEND_GENERIC=env), capture_output=True, text=True)
        outs[tag] = [l for l in r.stdout.splitlines() if "sha1" in l]
        print(f"--- {tag}\n" + "\n".join(outs[tag]) + ("\n" + r.stderr[-2000:] if r.returncode else ""))
    same = [a.split("sha1")[1].split()[0] == b.split("sha1")[1].split()[0] for a, b in zip(outs["fast"], outs["generic"])]
    print("bit-identical:", same, "ALL" if all(same) and len(same) == len(CASES) else "MISMATCH")
