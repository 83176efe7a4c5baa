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
                 betas=(0.9, 0.999), eps: float = 1e-8, max_grad_norm: float = 1.0, process_group=None, data_parallel: bool = False,
                 cuda_graph: bool = False):
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
        # cuda_graph=True: the ~450 launches of tib_train_loss_grad (both streams) are captured once per batch shape and
        # replayed; inputs are copied into the captured buffers (the optimiser step stays outside: its bias corrections
        # are host arguments that change every step)
        self.cuda_graph = cuda_graph
        self._graphs = {}
        self.last_grad: Optional[torch.Tensor] = None
        self.last_grad_sqnorm: Optional[torch.Tensor] = None
        if data_parallel:
            import torch.distributed as dist
            dist.broadcast(self.weights, src=0, group=self.group)      # replicas start from rank 0's weights

    def loss_and_grad(self, batch0, batch1, t=None, z=None, prepared=None):
        tb = prepared if prepared is not None else self.engine.prepare(batch0, batch1)
        if t is None:
            t = draw_times(tb.n_atoms, self.t_distr)
        if z is None:
            z = torch.randn(tb.x0.shape)
        if self.cuda_graph:
            return self._replay(tb, t, z)
        loss, grad, _ = self.engine.loss_and_grad(self.weights, tb, t, z, gamma=self.kind, a=self.a)
        return loss, grad

    _STATIC = ("mol_ptr", "edge_ptr", "atom_id", "edge_type", "temp0", "temp1")

    def _replay(self, tb, t, z):
        dev = self.engine.device
        t = t.to(dev, torch.float32).reshape(-1)
        z = z.to(dev, torch.float32)
        key = (tb.pb.n_mol, tb.pb.n_nodes, tb.pb.n_edges)
        entry = self._graphs.get(key)
        if entry is None:
            st_t, st_z = t.clone(), z.clone()
            self.engine.loss_and_grad(self.weights, tb, st_t, st_z, gamma=self.kind, a=self.a)      # warm: workspace, streams, events
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                loss, grad, _ = self.engine.loss_and_grad(self.weights, tb, st_t, st_z, gamma=self.kind, a=self.a)
            entry = self._graphs[key] = (graph, tb, st_t, st_z, loss, grad)
        graph, st_tb, st_t, st_z, loss, grad = entry
        if tb is not st_tb:
            for name in self._STATIC:
                getattr(st_tb.pb, name).copy_(getattr(tb.pb, name))
            st_tb.x0.copy_(tb.x0)
            st_tb.x1.copy_(tb.x1)
        st_t.copy_(t)
        st_z.copy_(z)
        graph.replay()
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
        return loss.clone() if self.cuda_graph else loss

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


def trainer(config, make_loaders, model=None, device: Optional[str] = None, data_parallel: bool = False, cuda_graph: bool = True,
            verbose: bool = True) -> dict:
    """The epoch loop of the reference's `trainer(config)` (mdqm9/train_ambient.py:22-181) around the native step.

    `config` carries the reference's fields (n_features, score_layers, temp_length, a, gamma, t_distr, learning_rate,
    weight_decay, n_epochs, seed, model_save_path, model_save_name); `make_loaders(epoch) -> (loader0, loader1)` yields the
    two batch streams the reference rebuilds every epoch (train_ambient.py:100-117; the RDKit / trajectory datasets behind
    them are out of scope).  Per batch: loss + gradients + clip_grad_norm_(1) + Adam in libtib.so, a non-finite loss skips
    the update (train_ambient.py:136-142); per epoch: the mean training loss drives ReduceLROnPlateau(factor 0.5, patience 10),
    a second pass evaluates the final weights, and two state_dicts are written exactly where the reference writes them -
    `<name>_<epoch>_weights.pt` and `<name>_best<epoch>_weights.pt` (the weights at the batch with the lowest loss of the
    epoch).  Returns the per-epoch losses."""
    import os

    import numpy as np

    from .ambient.interpolants import LinearInterpolant
    from .ambient.models.cpainn import cPaiNN

    out_dir = os.path.join(config.model_save_path, config.model_save_name)
    os.makedirs(out_dir, exist_ok=True)
    np.random.seed(config.seed)
    torch.manual_seed(config.seed)
    if model is None:
        model = cPaiNN(n_features=config.n_features, score_layers=config.score_layers, temp_length=config.temp_length)
    dev = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
    model.to(dev)
    tr = Trainer(model, LinearInterpolant(a=config.a, gamma=config.gamma), t_distr=config.t_distr, lr=config.learning_rate,
                 weight_decay=config.weight_decay, max_grad_norm=1.0, data_parallel=data_parallel, cuda_graph=cuda_graph)
    # torch's own plateau scheduler on a stand-in optimiser: its learning rate is copied into the native step
    knob = torch.optim.SGD([torch.zeros(1, requires_grad=True)], lr=config.learning_rate)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(knob, factor=0.5, patience=10)
    history = dict(train_loss=[], last_model_train_loss=[], epoch_best_loss=[], lr=[])
    for epoch in range(config.n_epochs):
        loader0, loader1 = make_loaders(epoch)
        total, n_batches, best, best_w = 0.0, 0, float("inf"), None
        for batch0, batch1 in zip(loader0, loader1):
            before = tr.weights.clone()
            loss, grad = tr.loss_and_grad(batch0, batch1)
            value = float(loss)                       # the reference reads loss.item() every batch too
            if value < best:                          # "epoch best model": the weights that produced the lowest loss
                best, best_w = value, before
            if not np.isfinite(value):
                if verbose:
                    print("NaN loss")
                continue
            tr.apply(grad)
            total += value
            n_batches += 1
        last = 0.0
        loader0, loader1 = make_loaders(epoch)
        for batch0, batch1 in zip(loader0, loader1):   # the final weights of the epoch on the training stream
            last += float(tr.loss_and_grad(batch0, batch1)[0])
        n_eval = max(n_batches, 1)
        total, last = total / n_eval, last / n_eval
        sched.step(total)
        tr.lr = knob.param_groups[0]["lr"]
        history["train_loss"].append(total); history["last_model_train_loss"].append(last)
        history["epoch_best_loss"].append(best); history["lr"].append(tr.lr)
        if verbose:
            print(f"Epoch {epoch + 1}/{config.n_epochs} - Train Loss: {total:.4f} - Last Train Loss: {last:.4f} - Epoch Best Loss: {best:.4f}")
        tr.sync_to_model()
        if not data_parallel or torch.distributed.get_rank() == 0:
            torch.save(model.state_dict(), os.path.join(out_dir, f"{config.model_save_name}_{epoch}_weights.pt"))
            if best_w is not None:
                final_w = tr.weights
                tr.weights = best_w
                tr.sync_to_model()
                torch.save(model.state_dict(), os.path.join(out_dir, f"{config.model_save_name}_best{epoch}_weights.pt"))
                tr.weights = final_w
                tr.sync_to_model()
    return history
