"""StreamingDecoder (stateful incremental inference; BASELINE configs[3]) against the offline forward of
the same module: identical kernels, carried fp32 state, 10-bin look-ahead -> logits must be BIT-IDENTICAL to
GRUDecoder.forward on the whole utterance, whatever the chunking."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import neural_speech_decoder_b200 as nsd
    from neural_speech_decoder_b200.synthetic import fill_trained_like_

DEV = "cuda"


def build(**kw):
    nsd.set_default_precision("bf16")
    try:
        torch.manual_seed(0)
        m = nsd.GRUDecoder(device=DEV, **kw)
    finally:
        nsd.set_default_precision("fp32")
    fill_trained_like_(m, seed=5)
    return m.to(DEV).eval()


@pytest.mark.parametrize("kw,B,T,chunks", [
    (dict(neural_dim=64, n_classes=10, hidden_dim=128, layer_dim=2, nDays=4, strideLen=4, kernelLen=32, gaussianSmoothWidth=2.0), 3, 171, [1, 7, 50, 3, 4, 4, 64, 38]),
    (dict(neural_dim=64, n_classes=10, hidden_dim=128, layer_dim=2, nDays=4, strideLen=4, kernelLen=32, gaussianSmoothWidth=2.0), 1, 96, [4] * 24),
    (dict(neural_dim=256, n_classes=40, hidden_dim=1024, layer_dim=5, nDays=24, strideLen=4, kernelLen=32, gaussianSmoothWidth=2.0), 2, 203, [60, 4, 4, 4, 31, 100]),
    (dict(neural_dim=32, n_classes=10, hidden_dim=64, layer_dim=2, nDays=4, strideLen=2, kernelLen=14, gaussianSmoothWidth=1.5), 2, 77, [20, 1, 1, 30, 25]),
])
def test_streaming_equals_offline(kw, B, T, chunks):
    assert sum(chunks) == T
    m = build(**kw)
    g = torch.Generator().manual_seed(T)
    X = torch.randn(B, T, kw["neural_dim"], generator=g).to(DEV)
    day = torch.randint(0, kw["nDays"], (B,), generator=g).to(DEV)
    with torch.no_grad():
        ref = m.forward(X, day)                                          # [B, T', C]
    sd = nsd.StreamingDecoder(m, B, day)
    outs, pos, emitted_after = [], 0, []
    for n in chunks:
        o = sd.push(X[:, pos:pos + n])
        pos += n
        if o is not None:
            outs.append(o)
        emitted_after.append(sd.next_frame)
        # causality: only frames whose 10-bin look-ahead has arrived may have been emitted
        assert sd.next_frame <= max(0, (pos - kw["kernelLen"] - 10) // kw["strideLen"] + 1)
    o = sd.finish()
    if o is not None:
        outs.append(o)
    got = torch.cat(outs, dim=1)
    assert got.shape == ref.shape
    assert torch.equal(got, ref)
    # a second utterance after reset() starts from the zero state again
    sd.reset()
    o1 = sd.push(X)
    o2 = sd.finish()
    assert torch.equal(torch.cat([t for t in (o1, o2) if t is not None], dim=1), ref)


def test_streaming_rejects_unsupported_models():
    kw = dict(neural_dim=32, n_classes=10, hidden_dim=64, layer_dim=1, nDays=2, strideLen=4, kernelLen=16, gaussianSmoothWidth=2.0)
    m = build(bidirectional=True, **kw)
    with pytest.raises(nsd.NsdError):
        nsd.StreamingDecoder(m, 1, torch.zeros(1, dtype=torch.int64))
