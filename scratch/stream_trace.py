"""NSD_STREAM_TRACE=1 python scratch/stream_trace.py [B] -> phase boundaries (ns, globaltimer of block 0) of one nsd_stream_push launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NSD_STREAM_TRACE", "1")
import torch
import neural_speech_decoder_b200 as nsd
from neural_speech_decoder_b200.synthetic import make_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0)
nsd.set_default_precision("bf16")
torch.manual_seed(0)
kw = dict(neural_dim=256, n_classes=40, hidden_dim=1024, layer_dim=5, nDays=24, dropout=0.0, strideLen=4, kernelLen=32, gaussianSmoothWidth=2.0)
m = nsd.GRUDecoder(device="cuda", bidirectional=False, **kw).to(dev).eval()
X, y, X_len, y_len, day = make_batch(B, 400, seed=2)
X = X.pin_memory()
sd = nsd.StreamingDecoder(m, B, day, use_graph=False)
L, H = 5, 1024
rows = []
for pos in range(0, 400, 4):
    sd.push_decode(X[:, pos:pos + 4])
    if sd._steady and pos > 200:
        off = (4 * L * 3 * H * B + 255) // 256 * 256
        t = sd._ws[off:off + 8 * (4 + 2 * L)].view(torch.int64).cpu().tolist()
        rows.append([t[i] - t[0] for i in range(len(t))])
import statistics
med = [statistics.median(r[i] for r in rows) for i in range(len(rows[0]))]
names = ["start", "phase0 done", "sync0"] + sum([[f"layer{l} done", f"sync{l + 1}"] for l in range(L)], []) + ["logits+argmax done"]
prev = 0
for n, v in zip(names, med):
    print(f"B={B} {n:22s} {v / 1000:8.2f} us  (+{(v - prev) / 1000:6.2f})")
    prev = v
