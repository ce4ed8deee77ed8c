"""Optimizer step of the item table.

The reference hands every parameter to one ``torch.optim.Adam(betas=(0.9, 0.99), eps=1e-09)``
(train_auto.py:354-364; wrapped by ``ScheduledOptim`` for pre-training,
Transformer/transformer.py:15-51).  For the ``[V1, d]`` item table that is ~10 elementwise
kernels per step, each streaming the whole table; :class:`FusedTableAdam` applies the same update
in one kernel, directly from the dense gradient the deterministic table backward (K2) produced.

Usage next to the reference's optimizer::

    table = model.src_emb.weight
    opt_rest = optim.Adam([p for p in model.parameters() if p is not table], lr=lr, betas=(0.9, 0.99), eps=1e-9)
    opt_table = FusedTableAdam(table, lr=lr)           # same hyper-parameters
    ...
    opt_rest.step(); opt_table.step()                  # uses table.grad, or pass grad=hp.dW

Sparse-gradient form ("without materialising a dense grad"): with
``TrainHotPath(..., dense_grad=False)`` the table backward writes only the rows of the step's
lookups into ``hp.dW`` (no zero fill) and stamps them in ``hp.row_epoch``;
``opt_table.step_rows(hp.dW, hp.row_epoch, hp.epoch)`` then applies dense-equivalent Adam
(untouched rows: g = 0) bit-identically to the dense call.

``param_groups`` mirrors torch's layout far enough for ``ScheduledOptim`` to set ``['lr']``.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops

Tensor = torch.Tensor


class FusedTableAdam:
    def __init__(self, weight: Tensor, lr: float = 1e-3, betas=(0.9, 0.99), eps: float = 1e-9):
        if not weight.is_cuda or weight.dtype != torch.float32 or not weight.is_contiguous():
            raise RuntimeError("recguru: FusedTableAdam needs a contiguous fp32 CUDA table (no CPU path exists)")
        self.weight = weight
        self.param_groups = [{"params": [weight], "lr": float(lr), "betas": tuple(betas), "eps": float(eps)}]
        self.state = {"step": 0, "exp_avg": torch.zeros_like(weight), "exp_avg_sq": torch.zeros_like(weight)}

    @torch.no_grad()
    def step(self, grad: Optional[Tensor] = None) -> None:
        g = self.weight.grad if grad is None else grad
        if g is None:
            return                                         # torch.optim.Adam skips parameters without a gradient
        grp = self.param_groups[0]
        self.state["step"] += 1
        w = self.weight.data if isinstance(self.weight, torch.nn.Parameter) else self.weight
        ops.raw_table_adam(w, self.state["exp_avg"], self.state["exp_avg_sq"], g.contiguous(), grp["lr"],
                           grp["betas"][0], grp["betas"][1], grp["eps"], self.state["step"])

    @torch.no_grad()
    def step_rows(self, grad_rows: Tensor, row_epoch: Tensor, epoch: int) -> None:
        """The same update from a ROW STORE instead of a dense gradient: ``grad_rows [V1, d]`` is valid
        only where ``row_epoch[row] == epoch`` (``TrainHotPath(dense_grad=False)`` produces the pair);
        all other rows take g = 0 without being read.  Bit-identical to :meth:`step` on the
        zero-filled gradient; saves the dense zero fill and the reads of untouched rows."""
        grp = self.param_groups[0]
        self.state["step"] += 1
        w = self.weight.data if isinstance(self.weight, torch.nn.Parameter) else self.weight
        ops.raw_table_adam_rows(w, self.state["exp_avg"], self.state["exp_avg_sq"], grad_rows, row_epoch, epoch,
                                grp["lr"], grp["betas"][0], grp["betas"][1], grp["eps"], self.state["step"])

    def zero_grad(self, set_to_none: bool = True) -> None:
        if isinstance(self.weight, torch.nn.Parameter) and self.weight.grad is not None:
            if set_to_none:
                self.weight.grad = None
            else:
                self.weight.grad.zero_()

    def state_dict(self) -> dict:
        return {"state": {0: dict(self.state)}, "param_groups": [{k: v for k, v in self.param_groups[0].items() if k != "params"}]}

    def load_state_dict(self, sd: dict) -> None:
        st = sd["state"][0]
        self.state["step"] = int(st["step"])
        self.state["exp_avg"].copy_(st["exp_avg"])
        self.state["exp_avg_sq"].copy_(st["exp_avg_sq"])
        for k, v in sd["param_groups"][0].items():
            self.param_groups[0][k] = v
