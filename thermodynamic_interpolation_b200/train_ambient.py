"""Training driver in the shape of mdqm9/train_ambient.py:96-148, with the whole step on the device:

    trainer = Trainer(model, interpolant, lr=1e-4, weight_decay=0, max_grad_norm=1.0)
    loss = trainer.step(batch0, batch1)        # loss + gradients, [grad all-reduce], clipping, Adam - all in libtib.so
    trainer.sync_to_model()                    # write the flat weights back into model.state_dict()

Weights, gradients and Adam moments are flat device vectors in the C ABI's packing order, so a step is
tib_train_loss_grad -> (data parallel: one NCCL all-reduce of the gradient vector, averaged over ranks as
DistributedDataParallel does) -> tib_adam_step, with no host synchronisation; the loss comes back as a device scalar.
The reference's own loop (optim.zero_grad / loss.backward / clip_grad_norm_ / optim.step) also works unchanged through
ambient/losses.py::StandardVelocityLoss."""
from __future__ import annotations

from typing import Optional

import torch

from .ambient.losses import draw_times
from .train import TrainEngine, flatten, packed_parameters


class Trainer:
    def __init__(self, model, interpolant, t_distr: str = "uniform", lr: float = 1e-4, weight_decay: float = 0.0,
                 betas=(0.9, 0.999), eps: float = 1e-8, max_grad_norm: float = 1.0, process_group=None, data_parallel: bool = False):
        self.model = model
        self.engine = TrainEngine(model.hyper, model.device)
        self.kind, self.a = interpolant.kind, interpolant.a_value
        if self.kind not in ("brownian", "sin2"):
            raise NotImplementedError("native training supports gamma 'brownian' and 'sin2'")
        self.t_distr = t_distr
        self.lr, self.weight_decay, self.betas, self.eps, self.max_grad_norm = lr, weight_decay, betas, eps, max_grad_norm
        self.weights = flatten(packed_parameters(model)).contiguous()
        self.m = torch.zeros_like(self.weights)
        self.v = torch.zeros_like(self.weights)
        self.step_no = 0
        self.data_parallel = data_parallel
        self.group = process_group
        self.last_grad: Optional[torch.Tensor] = None
        self.last_grad_sqnorm: Optional[torch.Tensor] = None
        if data_parallel:
            import torch.distributed as dist
            dist.broadcast(self.weights, src=0, group=self.group)      # replicas start from rank 0's weights

    def loss_and_grad(self, batch0, batch1, t=None, z=None):
        tb = self.engine.prepare(batch0, batch1)
        if t is None:
            t = draw_times(tb.n_atoms, self.t_distr)
        if z is None:
            z = torch.randn(tb.x0.shape)
        loss, grad, _ = self.engine.loss_and_grad(self.weights, tb, t, z, gamma=self.kind, a=self.a)
        return loss, grad

    def apply(self, grad: torch.Tensor):
        if self.data_parallel:
            import torch.distributed as dist
            dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=self.group)
            grad.div_(dist.get_world_size(self.group))
        self.step_no += 1
        self.last_grad = grad
        self.last_grad_sqnorm = self.engine.adam_step(self.weights, grad, self.m, self.v, self.step_no, lr=self.lr, betas=self.betas,
                                                      eps=self.eps, weight_decay=self.weight_decay, max_grad_norm=self.max_grad_norm)

    def step(self, batch0, batch1, t=None, z=None) -> torch.Tensor:
        """One optimisation step; returns the loss as a device scalar (fp64, shape [1]) without synchronising."""
        loss, grad = self.loss_and_grad(batch0, batch1, t, z)
        self.apply(grad)
        return loss

    def sync_to_model(self):
        """Copies the flat weights back into the model's parameters (state_dict order and keys untouched)."""
        off = 0
        with torch.no_grad():
            for p in packed_parameters(self.model):
                n = p.numel()
                p.copy_(self.weights[off:off + n].view(p.shape))
                off += n
        self.engine.status()
        return self.model
