"""Sparse mixture-of-experts MLP on the csrc/moe.cu kernels (reference: ``Qwen3MoeSparseMoeBlock``,
models/qwen3_moe.py:125-186, and its expert ``Qwen3MoeMLP`` :95-122).

Parameter names are the reference's -- ``gate.weight``, ``experts.N.gate_up_proj.weight`` (loaded
from ``gate_proj`` / ``up_proj`` through the packed-module mapping), ``experts.N.down_proj.weight``
-- but the storage is not: every rank keeps the experts it owns stacked in two tensors
(``w13 [n_local, 2*I, h]``, ``w2 [n_local, h, I]``) that the grouped kernels index by expert, and
the per-expert parameters are views into them, so checkpoints and the seeded initialiser fill the
stacks without knowing about them.

Parallel layout (differs from the reference on purpose): the reference shards every expert's
matrices across the tensor-parallel group and all-reduces once per expert inside a Python loop
(qwen3_moe.py:101-112,176-186). Here a rank owns ``E / tp`` experts WHOLE (expert parallelism over
the same group); tokens are already replicated on every rank, so dispatch is a local selection
and the combine is ONE sum all-reduce of the layer output -- the same collective a dense row-
parallel ``down_proj`` issues, fused with the following add+RMSNorm when it fits (comm.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

from .. import ops
from .linear import ReplicatedLinear, tp_collectives, tp_rank_size


class _ExpertMatrix(nn.Module):
    """One expert matrix: ``weight`` is a view into the rank's stacked tensor."""

    def __init__(self, storage: torch.Tensor, output_sizes: list[int] | None):
        super().__init__()
        self.weight = nn.Parameter(storage, requires_grad=False)
        self.weight.weight_loader = self.weight_loader
        self.tp_dim, self.tp_size = None, 1            # never sharded: experts are placed, not cut
        if output_sizes is not None:
            self.output_sizes = output_sizes           # gate | up, routed by shard id like a merged linear

    def weight_loader(self, param, loaded_weight, loaded_shard_id: int | None = None):
        if loaded_shard_id is None:
            param.data.copy_(loaded_weight)
            return
        start = sum(self.output_sizes[:loaded_shard_id])
        param.data.narrow(0, start, self.output_sizes[loaded_shard_id]).copy_(loaded_weight)


class _Expert(nn.Module):

    def __init__(self, w13: torch.Tensor, w2: torch.Tensor, inter: int):
        super().__init__()
        self.gate_up_proj = _ExpertMatrix(w13, [inter, inter])
        self.down_proj = _ExpertMatrix(w2, None)


class SparseMoeBlock(nn.Module):

    def __init__(self, hidden_size: int, moe_intermediate_size: int, num_experts: int, top_k: int):
        super().__init__()
        self.ep_rank, self.ep_size = tp_rank_size()
        if num_experts % self.ep_size:
            raise ValueError(f"{num_experts} experts do not divide over {self.ep_size} ranks")
        if top_k > 8 or num_experts > 256:
            raise ValueError("moe kernels support top_k <= 8 and <= 256 experts")
        self.hidden_size, self.inter, self.num_experts, self.top_k = (
            hidden_size, moe_intermediate_size, num_experts, top_k)
        self.num_local = num_experts // self.ep_size
        self.expert_begin = self.ep_rank * self.num_local
        self.gate = ReplicatedLinear(hidden_size, num_experts, bias=False)
        self.register_buffer("w13", torch.empty(self.num_local, 2 * self.inter, hidden_size), persistent=False)
        self.register_buffer("w2", torch.empty(self.num_local, hidden_size, self.inter), persistent=False)
        # sparse ModuleDict keyed by the GLOBAL expert index, so parameter names match the checkpoint
        self.experts = nn.ModuleDict({
            str(self.expert_begin + i): _Expert(self.w13[i], self.w2[i], self.inter)
            for i in range(self.num_local)})
        self.defer_allreduce = False       # set by the decoder layer: the next op is an add+RMSNorm

    def nonlocal_param_shapes(self, prefix: str) -> dict[str, tuple[int, ...]]:
        """Names/shapes of the expert parameters other ranks hold (the seeded initialiser numbers
        parameters over the FULL name list so that every layout draws the same weights)."""
        out = {}
        for e in range(self.num_experts):
            if not self.expert_begin <= e < self.expert_begin + self.num_local:
                out[f"{prefix}experts.{e}.gate_up_proj.weight"] = (2 * self.inter, self.hidden_size)
                out[f"{prefix}experts.{e}.down_proj.weight"] = (self.hidden_size, self.inter)
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        tokens, hidden = x.shape
        router_logits = F.linear(x, self.gate.weight)
        topk_ids, topk_w = ops.moe_route(router_logits, self.top_k)
        plan = ops.moe_align(topk_ids, self.expert_begin, self.num_local)
        out = None
        if self.ep_size > 1 and self.defer_allreduce:
            from ..comm import get_comm
            comm = get_comm()
            if comm is not None and comm.fits(tokens, hidden, x.dtype.itemsize):
                out = comm.partial_buffer(tokens, hidden, x.dtype)   # RMSNorm completes the all-reduce
                return ops.moe_experts(x, self.w13, self.w2, topk_ids, topk_w, plan, out=out)
        y = ops.moe_experts(x, self.w13, self.w2, topk_ids, topk_w, plan)
        if self.ep_size > 1 and tp_collectives():
            dist.all_reduce(y)
        return y
