"""Gradient hub: the reference's distributed-learning server on collectives.

``AgentServer`` (bots/bot-0.5/AgentServer.cpp) is a TCP hub: clients push the gradients of their
``AgentModel`` ('G', AgentClient.hpp:72-82), the server sums them in client order, divides by the
number of contributing clients, takes ONE ``torch::optim::AdamW`` step with ``AdamWOptions(lr)`` on
its own copy of the model (``aggregate_and_update``, :465-511), and hands every client the *update
vector* ``theta_new - theta_old`` (``compute_update_vector`` :513-524; 'U', AgentClient.hpp:84-103),
which the client adds to its parameters.

With one process per GPU that exchange is one all-reduce: every rank flattens its gradients into
one bucket (plus a "did I contribute" count), the bucket is summed over the ranks (NCCL on the
GPUs, riding NVLink; gloo in the CPU tests), and every rank then takes the identical AdamW step on
its own replica -- the server's model and optimiser state exist on every rank, bit-identical,
instead of on a hub, and the update vector never has to travel.  ``GradientHub.step`` returns it
anyway, because it is what the reference's clients consume.

AdamW defaults are libtorch's ``AdamWOptions``: betas (0.9, 0.999), eps 1e-8, weight_decay 1e-2,
amsgrad off; lr 1e-3 (AgentServer.cpp:614).  Pinned against libtorch's own AdamW run around the
reference's AgentModel: oracle/ref_harness/hub_oracle.cpp -> tests/golden/hub_golden.bin.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradientHub:
    def __init__(self, parameters, lr=1e-3, group=None):
        self.params = [p for p in parameters]
        if not self.params:
            raise ValueError("GradientHub: no parameters")
        self.group = group
        # foreach=False: one tensor at a time, the order of operations of libtorch's AdamW::step
        self.opt = torch.optim.AdamW(self.params, lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False,
                                     foreach=False)
        self.sizes = [p.numel() for p in self.params]
        dev, dt = self.params[0].device, self.params[0].dtype
        self.bucket = torch.zeros(sum(self.sizes) + 1, dtype=dt, device=dev)  # gradients | contributing clients
        self.version = 0  # update_version, AgentServer.cpp:441

    def _world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    @torch.no_grad()
    def step(self, grads=None, contributing=True):
        """One update round.  ``grads``: this rank's gradients, one tensor per parameter (default: the
        ``.grad`` of the parameters); ``contributing=False``: this rank has none this round
        (a client whose ``gradient_ready`` is false: it does not count in the mean, :470-474).
        Returns the update vector (a list shaped like the parameters), or None when nobody contributed."""
        if contributing:
            grads = [p.grad for p in self.params] if grads is None else list(grads)
            if len(grads) != len(self.params) or any(g is None for g in grads):
                raise ValueError("GradientHub.step: one gradient per parameter is needed")
            off = 0
            for g, n in zip(grads, self.sizes):
                self.bucket[off:off + n].copy_(g.reshape(-1))
                off += n
            self.bucket[-1] = 1
        else:
            self.bucket.zero_()
        if self._world() > 1:
            dist.all_reduce(self.bucket, op=dist.ReduceOp.SUM, group=self.group)
        n = float(self.bucket[-1].item())
        if n == 0:
            return None  # "No clients with gradients!", :476-479
        theta_old = [p.detach().clone() for p in self.params]
        self.opt.zero_grad(set_to_none=True)
        off = 0
        for p, sz in zip(self.params, self.sizes):
            p.grad = (self.bucket[off:off + sz] / n).view_as(p).clone()  # sum_grad / contributing_clients, :489
            off += sz
        self.opt.step()
        self.version += 1
        return [p.detach() - o for p, o in zip(self.params, theta_old)]

    @staticmethod
    @torch.no_grad()
    def apply_update(parameters, update):
        """What a reference client does with the update vector it fetched: theta += update."""
        for p, u in zip(parameters, update):
            p.add_(u)
