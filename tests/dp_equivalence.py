"""Data-parallel equivalence on real GPUs: an N-rank step through the CUDA kernels with the
bucketed NCCL all-reduce overlapped with the backward must equal ONE rank stepping on the concatenated batch.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tests/dp_equivalence.py        (tests/test_gpu_dp.py wraps it)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    backend = os.environ.get("NSD_DP_BACKEND", "nccl")     # "gloo": both ranks may share ONE GPU (NCCL refuses duplicate devices)
    local = local % torch.cuda.device_count()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group(backend)
    import neural_speech_decoder_b200 as nsd
    from neural_speech_decoder_b200.parallel import GradSync
    from neural_speech_decoder_b200.synthetic import fill_trained_like_, make_batch
    kw = dict(neural_dim=256, n_classes=40, hidden_dim=256, layer_dim=3, nDays=6, dropout=0.0, strideLen=4, kernelLen=32,
              gaussianSmoothWidth=2.0, bidirectional=True)
    per = 6
    full = make_batch(per * world, 120, n_days=6, seed=9, ragged=True, min_tgt=3, max_tgt=12)
    ok = True
    for precision, tol in (("bf16", 2e-3), ("fp32", 2e-5)):
        nsd.set_default_precision(precision)

        def fresh():
            torch.manual_seed(0)
            m = nsd.GRUDecoder(device="cuda", **kw)
            fill_trained_like_(m, seed=4)
            return m.to(dev).train()

        args = dict(lrStart=0.02, lrEnd=0.02, nBatch=100, l2_decay=1e-5)
        m = fresh()
        opt, sched = nsd.make_optimizer(m, args)
        gs = GradSync(world)                                    # train_step folds gs.grad_scale into the optimizer itself
        shard = [t[rank * per:(rank + 1) * per].contiguous().to(dev) for t in full]
        loss = nsd.train_step(m, opt, *shard, scheduler=sched, grad_sync=gs)
        assert opt.grad_scale == gs.grad_scale
        losses = [torch.zeros_like(loss) for _ in range(world)]
        dist.all_gather(losses, loss)
        torch.cuda.synchronize()
        if rank == 0:
            ref = fresh()
            ropt, rsched = nsd.make_optimizer(ref, args)
            rloss = nsd.train_step(ref, ropt, *[t.to(dev) for t in full], scheduler=rsched)
            worst = 0.0
            for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
                if p.grad is None:
                    assert q.grad is None, n
                    continue
                g = p.grad * gs.grad_scale                      # all-reduced sum -> mean over ranks
                rel = ((g - q.grad).norm() / q.grad.norm().clamp_min(1e-20)).item()
                worst = max(worst, rel)
                dp = (p.detach() - q.detach()).abs().max().item()
                if rel > tol or dp > 1e-5:
                    ok = False
                    print(f"[{precision}] MISMATCH {n}: grad rel L2 {rel:.2e}, param max diff {dp:.2e}")
            mean_loss = torch.stack(losses).mean().item()
            print(f"[{precision}] world {world}: worst grad rel L2 err {worst:.2e} (tol {tol:.0e}); "
                  f"mean of rank losses {mean_loss:.6f} vs single-rank loss {rloss.item():.6f}; all-reduced bytes {gs.bytes}")
            if abs(mean_loss - rloss.item()) > 1e-4 * max(1.0, abs(rloss.item())):
                ok = False
        dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DP_EQUIVALENCE_OK" if ok else "DP_EQUIVALENCE_FAILED")
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
