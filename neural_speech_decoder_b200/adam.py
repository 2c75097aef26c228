"""Adam with the reference's semantics (torch.optim.Adam: L2 folded into the gradient, bias-corrected,
``eps`` added to sqrt(v_hat); neural_decoder_trainer.py:163-169, 259) as one multi-tensor CUDA launch."""
from __future__ import annotations

import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.grad_scale = grad_scale          # data parallel: 1/world_size folds the gradient average in
        self.shadows = None                   # model_tc.Bf16Shadows: bf16 operand copies refreshed by the update kernel

    def attach_shadows(self, shadows) -> None:
        """Let the update kernel rewrite the model's bf16 weight copies in the same pass (saves the per-step casts)."""
        self.shadows = shadows

    @torch.no_grad()
    def step(self, closure=None, only=None, under_recurrence=False):
        """``only``: optional collection of parameters -- update just those (each parameter keeps its own step count, so a step
        may be split into several calls; the data-parallel train_step updates the buckets whose all-reduce has finished while
        the last one is still on the wire).  ``under_recurrence``: the call comes from inside the backward, directly behind the launch of a
        recurrence kernel that does not produce these parameters' gradients: the update runs under that kernel (ops.adam_step)."""
        loss = None
        only_ids = None if only is None else {id(p) for p in only}
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps, gs, ms, vs = [], [], [], []
            step = None
            touched = []
            for p in group["params"]:
                if p.grad is None:            # the reference's dead inpLayer* parameters never get one
                    continue
                if only_ids is not None and id(p) not in only_ids:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("FusedAdam (B200) handles CUDA float32 parameters only")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                touched.append(p)
                step = st["step"] if step is None else step
                if st["step"] != step:        # parameters that joined later: separate launch group
                    with torch.cuda.device(p.device):
                        ops.adam_step([p], [p.grad.contiguous()], [st["exp_avg"]], [st["exp_avg_sq"]], group["lr"],
                                      *group["betas"], group["eps"], group["weight_decay"], st["step"], self.grad_scale)
                    continue
                ps.append(p); gs.append(p.grad if p.grad.is_contiguous() else p.grad.contiguous())
                ms.append(st["exp_avg"]); vs.append(st["exp_avg_sq"])
            sh = [self.shadows.slice_for(p) for p in ps] if self.shadows is not None else None
            if ps:
                if any(p.device != ps[0].device for p in ps):
                    raise RuntimeError("FusedAdam (B200): the parameters of one group must live on one device")
                with torch.cuda.device(ps[0].device):
                    ops.adam_step(ps, gs, ms, vs, group["lr"], *group["betas"], group["eps"], group["weight_decay"], step,
                                  self.grad_scale, sh, under_recurrence=under_recurrence and len(ps) <= 48)
            # the kernels write through raw pointers: tell autograd (and the shadow cache) that the parameters changed
            if touched:
                torch.autograd.graph.increment_version(touched)
            if sh is not None:
                for p, s_ in zip(ps, sh):
                    if s_ is not None:
                        self.shadows.mark_fresh(p)
        return loss
