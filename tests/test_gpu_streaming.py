"""StreamingDecoder (stateful incremental inference; BASELINE configs[3]).
Exact form: identical kernels, carried fp32 state, 10-bin look-ahead -> logits BIT-IDENTICAL to GRUDecoder.forward on the
whole utterance, whatever the chunking.  Fast form (batch <= 8, one stride per push: nsd_stream_push, CUDA graph): against
an oracle built from the reference's operators with the state carried between calls, and against the offline forward."""
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
    sd = nsd.StreamingDecoder(m, B, day, fast=False)                      # the exact form: same kernels as the offline forward
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


def port_streaming_oracle(port, X, day, frames_per_call):
    """Oracle for the streaming path built from the REFERENCE's operators (oracle/torch_port.py): the front end of the whole
    utterance (smoothing -> day affine -> softsign -> unfold, model.py:84-101), then ``nn.GRU`` fed ``frames_per_call`` frames at
    a time with the final state h_n of one call carried into the next as h0 -- what a stateful version of model.py:104-119
    computes -- and the output layer.  Returns logits [B, T', C]."""
    import torch.nn.functional as F
    with torch.no_grad():
        x = X.permute(0, 2, 1)
        x = F.conv1d(x, port.smooth_w, groups=port.neural_dim, padding="same").permute(0, 2, 1)
        w = torch.index_select(port.dayWeights, 0, day)
        x = F.softsign(torch.einsum("btd,bdk->btk", x, w) + torch.index_select(port.dayBias, 0, day))
        x = F.unfold(x.permute(0, 2, 1).unsqueeze(3), (port.kernelLen, 1), stride=port.strideLen).permute(0, 2, 1)
        h = torch.zeros(port.layer_dim, x.size(0), port.hidden_dim, dtype=x.dtype)
        outs = []
        for j in range(0, x.size(1), frames_per_call):
            hid, h = port.gru(x[:, j:j + frames_per_call], h)            # carried h_n
            outs.append(port.fc(hid))
        return torch.cat(outs, dim=1)


def test_streaming_vs_reference_operators_with_carried_state():
    """StreamingDecoder at the competition architecture (unidirectional 5x1024), 4 bins per push, against the chunked
    reference-operator oracle above (CPU, fp64).  bf16 tolerance as stated in tests/test_gpu_fullshape.py (BF16_TOL)."""
    from oracle import torch_port as P
    kw = dict(neural_dim=256, n_classes=40, hidden_dim=1024, layer_dim=5, nDays=24, strideLen=4, kernelLen=32, gaussianSmoothWidth=2.0)
    B, T = 2, 260
    m = build(**kw)
    port = P.PortGRUDecoder(bidirectional=False, dropout=0.0, **kw)
    port.load_reference_state({k: v.detach().cpu() for k, v in m.state_dict().items()})
    port = port.double().eval()
    g = torch.Generator().manual_seed(77)
    X = torch.randn(B, T, 256, generator=g)
    day = torch.randint(0, 24, (B,), generator=g)
    ref = port_streaming_oracle(port, X.double(), day, frames_per_call=1).numpy()
    # the chunked oracle equals the reference's offline forward (h0 = 0 once): carrying h_n is exact
    with torch.no_grad():
        np.testing.assert_allclose(ref, port(X.double(), day).numpy(), rtol=0, atol=1e-10)
    sd = nsd.StreamingDecoder(m, B, day.to(DEV))
    outs = []
    for pos in range(0, T, 4):
        o = sd.push(X[:, pos:pos + 4].to(DEV))
        if o is not None:
            outs.append(o)
    o = sd.finish()
    if o is not None:
        outs.append(o)
    got = torch.cat(outs, dim=1).cpu().numpy().astype(np.float64)
    assert got.shape == ref.shape
    err = np.abs(got - ref).max()
    agree = (got.argmax(-1) == ref.argmax(-1)).mean()
    print(f"streaming vs chunked reference operators: logits max abs err {err:.3e} (|ref| max {np.abs(ref).max():.2f}), argmax agreement {agree:.4f}")
    top2 = np.sort(ref, axis=-1)[..., -2:]                    # greedy decisions may differ only at near ties of the oracle
    flips = got.argmax(-1) != ref.argmax(-1)
    assert not (flips & ((top2[..., 1] - top2[..., 0]) > 2 * err)).any()
    assert err < 0.12 and agree >= 0.95


@pytest.mark.parametrize("B,use_graph", [(1, True), (3, True), (8, False)])
def test_streaming_fast_form(B, use_graph):
    """The single-launch push (nsd_stream_push: front end + stack + logits + greedy id) engaged by steady 4-bin pushes, with and without CUDA-graph replay:
    vs the chunked reference-operator oracle (bf16 tolerance), vs the offline forward of the module (same bf16 arithmetic,
    different fp32 summation order -> a tight bound), greedy ids == argmax of the returned logits, and a mid-stream irregular
    push falls back to the exact form and re-engages."""
    from oracle import torch_port as P
    kw = dict(neural_dim=256, n_classes=40, hidden_dim=1024, layer_dim=5, nDays=24, strideLen=4, kernelLen=32, gaussianSmoothWidth=2.0)
    T = 232
    m = build(**kw)
    port = P.PortGRUDecoder(bidirectional=False, dropout=0.0, **kw)
    port.load_reference_state({k: v.detach().cpu() for k, v in m.state_dict().items()})
    port = port.double().eval()
    g = torch.Generator().manual_seed(5 + B)
    X = torch.randn(B, T, 256, generator=g)
    day = torch.randint(0, 24, (B,), generator=g)
    ref = port_streaming_oracle(port, X.double(), day, frames_per_call=1).numpy()
    with torch.no_grad():
        off = m.forward(X.to(DEV), day.to(DEV)).cpu().numpy().astype(np.float64)
    sd = nsd.StreamingDecoder(m, B, day.to(DEV), use_graph=use_graph)
    assert sd.fast
    outs, fast_frames, pos = [], 0, 0
    sizes = [4] * 30 + [6] + [4] * 100                     # one irregular push in the middle
    for n in sizes:
        if pos >= T:
            break
        n = min(n, T - pos)
        o = sd.push(X[:, pos:pos + n].to(DEV))
        pos += n
        if o is not None:
            outs.append(o)
            if sd._steady:
                fast_frames += o.shape[1]
                assert torch.equal(sd.last_ids.long(), o[:, -1].argmax(-1))
    o = sd.finish()
    if o is not None:
        outs.append(o)
    got = torch.cat(outs, dim=1).cpu().numpy().astype(np.float64)
    assert got.shape == ref.shape and fast_frames >= 30
    if use_graph:
        assert sd._graph is not None, getattr(sd, "_graph_error", "no graph captured")
    e_ref, e_off = np.abs(got - ref).max(), np.abs(got - off).max()
    agree = (got.argmax(-1) == ref.argmax(-1)).mean()
    print(f"fast streaming B={B} graph={use_graph}: {fast_frames} frames through the step kernel; vs oracle {e_ref:.3e}, vs offline forward {e_off:.3e}, argmax agreement {agree:.4f}")
    # greedy decisions may differ from the fp64 oracle only at near ties (top-2 margin of the oracle below twice the logit error)
    top2 = np.sort(ref, axis=-1)[..., -2:]
    flips = got.argmax(-1) != ref.argmax(-1)
    assert not (flips & ((top2[..., 1] - top2[..., 0]) > 2 * e_ref)).any()
    assert e_ref < 0.12 and agree >= 0.95 and e_off < 0.06


def test_streaming_push_decode_host_path():
    """push_decode (pinned host bins in, greedy ids back through pinned memory, one graph replay + one synchronisation per
    push) returns exactly the argmax of the logits that push() returns for the same stream."""
    kw = dict(neural_dim=256, n_classes=40, hidden_dim=1024, layer_dim=5, nDays=24, strideLen=4, kernelLen=32, gaussianSmoothWidth=2.0)
    B, T = 2, 160
    m = build(**kw)
    g = torch.Generator().manual_seed(11)
    X = torch.randn(B, T, 256, generator=g)
    day = torch.randint(0, 24, (B,), generator=g).to(DEV)
    a, b = nsd.StreamingDecoder(m, B, day), nsd.StreamingDecoder(m, B, day)
    Xp = X.pin_memory()
    n_fast = 0
    for pos in range(0, T, 4):
        o = a.push(X[:, pos:pos + 4].to(DEV))
        ids = b.push_decode(Xp[:, pos:pos + 4])
        assert (o is None) == (ids is None)
        if o is not None:
            assert ids.dtype == torch.int32 and not ids.is_cuda
            assert torch.equal(ids.long(), o.argmax(-1).cpu())
            n_fast += int(b._steady)
    assert n_fast >= 20 and b._graph is not None
