"""RMSNorm / fused residual-add RMSNorm on the F1 kernels (reference: layers/layernorm.py).

``forward(x)`` returns a new tensor. ``forward(x, residual)`` returns ``(normed, new_residual)``
like the reference but computes them IN PLACE in the two argument buffers (one pass over HBM
instead of the reference's seven eager kernels); the model code rebinds both names, exactly as
``Qwen3DecoderLayer.forward`` does (models/qwen3.py:154-159).
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


class RMSNorm(nn.Module):

    def __init__(self, hidden_size: int, eps: float = 1e-6) -> None:
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(hidden_size))

    def forward(self, x: torch.Tensor, residual: torch.Tensor | None = None):
        if residual is None:
            return ops.rmsnorm(x.contiguous(), self.weight, self.eps)
        from ..comm import get_comm
        comm = get_comm()
        if comm is not None and comm.pending is not None and comm.pending[0] is x:
            # x is this rank's PARTIAL row-parallel output sitting in the peer-mapped buffer:
            # all-reduce + residual add + norm in one kernel over NVLink (C1 fused)
            return comm.allreduce_add_rmsnorm(None, residual, self.weight, self.eps)
        return ops.add_rmsnorm_(x, residual, self.weight, self.eps)
