"""Data-parallel host logic on CPU (gloo, world size 2): GradSync buckets + the exactness of "per-rank mean loss, summed
gradients, 1/N folded into the optimizer" against one rank stepping on the concatenated batch.
The model here is the torch-operator port (the kernels need a GPU); what is under test is the sharding/reduction scheme."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neural_speech_decoder_b200.parallel import GradSync
from neural_speech_decoder_b200.synthetic import make_batch

KW = dict(neural_dim=16, n_classes=10, hidden_dim=32, layer_dim=2, nDays=3, dropout=0.0, strideLen=4, kernelLen=14,
          gaussianSmoothWidth=2.0, bidirectional=True)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _grads(batch):
    from oracle import torch_port as P
    torch.manual_seed(0)
    m = P.PortGRUDecoder(**KW)
    X, y, X_len, y_len, day = batch
    pred = m(X, day)
    lens = ((X_len - m.kernelLen) / m.strideLen).to(torch.int32)
    loss = torch.nn.CTCLoss(blank=0, reduction="mean", zero_infinity=True)(pred.log_softmax(2).permute(1, 0, 2), y, lens, y_len)
    loss.backward()
    return [p.grad.clone() for p in m.parameters()], loss.item()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    full = make_batch(4, 61, n_feat=16, n_days=3, n_classes=10, seed=3, ragged=False, min_tgt=2, max_tgt=5, kernel_len=14, stride_len=4)
    shard = [t[rank * 2:(rank + 1) * 2] for t in full]
    shard[1] = shard[1].contiguous()
    grads, _ = _grads(shard)
    gs = GradSync(world)
    assert gs.grad_scale == 0.5
    gs.begin()
    # two buckets, as the backward produces them: flat buffers whose views are the per-parameter gradients
    half = len(grads) // 2
    flats = []
    for chunk in (grads[:half], grads[half:]):
        flat = torch.cat([g.reshape(-1) for g in chunk])
        gs.bucket_ready(flat)
        flats.append((flat, chunk))
    # the split finish the trainer uses: everything before the last bucket of >= tail_bytes is waited for, the tail stays in flight
    pending = gs.finish_early(tail_bytes=flats[1][0].numel() * 4)
    assert pending == {flats[1][0].untyped_storage().data_ptr()} and len(gs.handles) == 1
    gs.finish()
    assert gs.handles == []
    assert gs.bytes == sum(f.numel() * 4 for f, _ in flats)
    if rank == 0:
        red = []
        for flat, chunk in flats:
            off = 0
            for g in chunk:
                red.append((flat[off:off + g.numel()] * gs.grad_scale).view_as(g).clone())
                off += g.numel()
        ref, _ = _grads(full)
        out.put([float((a - b).abs().max() / (b.abs().max() + 1e-12)) for a, b in zip(red, ref)])
    dist.destroy_process_group()


def test_two_rank_step_equals_single_rank_on_concatenated_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    errs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert max(errs) < 1e-5, errs


def test_gradsync_single_rank_is_a_noop():
    gs = GradSync(1)
    gs.begin()
    t = torch.ones(8)
    gs.bucket_ready(t)
    gs.finish()
    assert gs.bytes == 0 and torch.equal(t, torch.ones(8)) and gs.grad_scale == 1.0
