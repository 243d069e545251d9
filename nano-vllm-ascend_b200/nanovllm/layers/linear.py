"""Tensor-parallel linear layers (reference: layers/linear.py).

Same constructors and ``weight_loader(param, tensor[, shard_id])`` signatures. GEMMs stay on
cuBLAS (``F.linear``). The one collective of a decoder layer is the SUM all-reduce
of the row-parallel output: NCCL over NVLink/NVSwitch through ``torch.distributed``, issued on
the current stream so it is captured into the decode CUDA graph.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn


from .. import ops
from ..utils.context import get_context

_TP_OVERRIDE: tuple[int, int] | None = None
_TP_COLLECTIVES = True


def set_tp_world(rank: int | None, size: int | None = None, collectives: bool = True) -> None:
    """Pin the tensor-parallel (rank, size) the layers shard by. ``None`` restores the
    reference behaviour (the default process group's rank/size, linear.py:23-24). The runner
    pins (0, 1) for data-parallel replicas that share a process group only for timing.
    ``collectives=False`` (Config.tp_shard_emulation, a profiling aid): build one rank's shard of a
    size-way model in a single process and skip every collective -- the per-rank kernel timeline of
    a TP=8 step on one GPU; the numbers it computes are meaningless."""
    global _TP_OVERRIDE, _TP_COLLECTIVES
    _TP_OVERRIDE = None if rank is None else (int(rank), int(size))
    _TP_COLLECTIVES = bool(collectives)


def tp_collectives() -> bool:
    return _TP_COLLECTIVES


def tp_rank_size() -> tuple[int, int]:
    if _TP_OVERRIDE is not None:
        return _TP_OVERRIDE
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def matmul(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None) -> torch.Tensor:
    """x @ weight.T (+ bias): cuBLAS, or -- when the step context asks for it (Context.own_gemm, the
    two-micro-batch decode step) -- the library's single-CTA tcgen05 kernel, which unlike cuBLAS'
    clustered 2-CTA kernels can run on whatever SMs a resident attention kernel leaves free."""
    if (bias is None and x.dim() == 2 and get_context().own_gemm and x.is_contiguous()
            and weight.shape[0] % 128 == 0 and weight.shape[1] % 64 == 0):
        return ops.linear_umma(x, weight)
    return F.linear(x, weight, bias)


def divide(a: int, b: int) -> int:
    if a % b:
        raise ValueError(f"{a} is not divisible by the tensor-parallel size {b}")
    return a // b


class LinearBase(nn.Module):

    def __init__(self, input_size, output_size, bias=False, tp_dim=None):
        super().__init__()
        self.tp_dim = tp_dim
        self.tp_rank, self.tp_size = tp_rank_size()
        self.weight = nn.Parameter(torch.empty(output_size, input_size))
        self.weight.weight_loader = self.weight_loader
        if bias:
            self.bias = nn.Parameter(torch.empty(output_size))
            self.bias.weight_loader = self.weight_loader
        else:
            self.register_parameter("bias", None)


class ReplicatedLinear(LinearBase):

    def __init__(self, input_size, output_size, bias=False):
        super().__init__(input_size, output_size, bias)

    def weight_loader(self, param, loaded_weight):
        param.data.copy_(loaded_weight)

    def forward(self, x):
        return matmul(x, self.weight, self.bias)


class ColumnParallelLinear(LinearBase):
    """Output features sharded across ranks."""

    def __init__(self, input_size, output_size, bias=False):
        _, tp = tp_rank_size()
        super().__init__(input_size, divide(output_size, tp), bias, 0)

    def weight_loader(self, param, loaded_weight):
        n = param.data.size(self.tp_dim)
        param.data.copy_(loaded_weight.narrow(self.tp_dim, self.tp_rank * n, n))

    def forward(self, x):
        return matmul(x, self.weight, self.bias)


class MergedColumnParallelLinear(ColumnParallelLinear):
    """Several column-parallel matrices stacked (gate|up); each is sharded on its own."""

    def __init__(self, input_size, output_sizes, bias=False):
        self.output_sizes = list(output_sizes)
        super().__init__(input_size, sum(output_sizes), bias)

    def weight_loader(self, param, loaded_weight, loaded_shard_id: int):
        start = sum(self.output_sizes[:loaded_shard_id]) // self.tp_size
        size = self.output_sizes[loaded_shard_id] // self.tp_size
        piece = loaded_weight.chunk(self.tp_size, self.tp_dim)[self.tp_rank]
        param.data.narrow(self.tp_dim, start, size).copy_(piece)


class QKVParallelLinear(ColumnParallelLinear):
    """Fused q|k|v projection; each rank holds H/tp query heads and Hkv/tp kv heads."""

    def __init__(self, hidden_size, head_size, total_num_heads, total_num_kv_heads=None, bias=False):
        _, tp = tp_rank_size()
        total_num_kv_heads = total_num_kv_heads or total_num_heads
        self.head_size = head_size
        self.num_heads = divide(total_num_heads, tp)
        # kv heads: sharded when there are at least `tp` of them, otherwise each one is replicated on
        # tp / total consecutive ranks (one kv head per rank; an extension of the reference)
        self.kv_replicas = 1 if total_num_kv_heads >= tp else divide(tp, total_num_kv_heads)
        self.num_kv_heads = divide(total_num_kv_heads, tp) if self.kv_replicas == 1 else 1
        rows = (self.num_heads + 2 * self.num_kv_heads) * head_size
        LinearBase.__init__(self, hidden_size, rows, bias, 0)
        self.total_num_kv_heads = total_num_kv_heads

    def weight_loader(self, param, loaded_weight, loaded_shard_id: str):
        q_rows = self.num_heads * self.head_size
        kv_rows = self.num_kv_heads * self.head_size
        start, size = {"q": (0, q_rows), "k": (q_rows, kv_rows), "v": (q_rows + kv_rows, kv_rows)}[
            loaded_shard_id]
        if loaded_shard_id == "q" or self.kv_replicas == 1:
            piece = loaded_weight.chunk(self.tp_size, self.tp_dim)[self.tp_rank]
        else:
            piece = loaded_weight.chunk(self.total_num_kv_heads, self.tp_dim)[self.tp_rank // self.kv_replicas]
        param.data.narrow(self.tp_dim, start, size).copy_(piece)


class RowParallelLinear(LinearBase):
    """Input features sharded; partial products are summed across ranks (the layer's only
    collective, reference linear.py:149-153). Bias is added on rank 0 only, before the sum."""

    def __init__(self, input_size, output_size, bias=False):
        _, tp = tp_rank_size()
        super().__init__(divide(input_size, tp), output_size, bias, 1)
        # set by decoder stacks whose next op is a residual-add RMSNorm: the all-reduce is then
        # deferred and fused into that norm over NVLink peer memory (comm.py) for small messages
        self.defer_allreduce = False

    def weight_loader(self, param, loaded_weight):
        if param.dim() <= self.tp_dim:          # bias: replicated, applied on rank 0 only
            param.data.copy_(loaded_weight)
            return
        n = param.data.size(self.tp_dim)
        param.data.copy_(loaded_weight.narrow(self.tp_dim, self.tp_rank * n, n))

    def forward(self, x):
        bias = self.bias if self.tp_rank == 0 else None
        if self.tp_size > 1 and self.defer_allreduce and x.dim() == 2:
            from ..comm import get_comm
            comm = get_comm()
            if comm is not None and comm.fits(x.shape[0], self.weight.shape[0], x.dtype.itemsize):
                buf = comm.partial_buffer(x.shape[0], self.weight.shape[0], x.dtype)
                if bias is None:
                    torch.mm(x, self.weight.t(), out=buf)
                else:
                    torch.addmm(bias, x, self.weight.t(), out=buf)
                return buf          # partial sum; RMSNorm(x, residual) completes the all-reduce
        y = matmul(x, self.weight, bias)
        if self.tp_size > 1 and _TP_COLLECTIVES:
            dist.all_reduce(y)
        return y
