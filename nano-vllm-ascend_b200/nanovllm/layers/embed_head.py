"""Vocab-parallel embedding and LM head (reference: layers/embed_head.py)."""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

from ..utils.context import get_context
from .linear import tp_collectives, tp_rank_size


class VocabParallelEmbedding(nn.Module):

    def __init__(self, num_embeddings: int, embedding_dim: int):
        super().__init__()
        self.tp_rank, self.tp_size = tp_rank_size()
        if num_embeddings % self.tp_size:
            raise ValueError("vocab size must divide by the tensor-parallel size")
        self.num_embeddings = num_embeddings
        self.num_embeddings_per_partition = num_embeddings // self.tp_size
        self.vocab_start_idx = self.num_embeddings_per_partition * self.tp_rank
        self.vocab_end_idx = self.vocab_start_idx + self.num_embeddings_per_partition
        self.weight = nn.Parameter(torch.empty(self.num_embeddings_per_partition, embedding_dim))
        self.weight.weight_loader = self.weight_loader

    def weight_loader(self, param, loaded_weight):
        n = param.data.size(0)
        param.data.copy_(loaded_weight.narrow(0, self.tp_rank * n, n))

    def forward(self, x):
        if self.tp_size == 1:
            return F.embedding(x, self.weight)
        mine = (x >= self.vocab_start_idx) & (x < self.vocab_end_idx)
        y = F.embedding(mine * (x - self.vocab_start_idx), self.weight)
        y = mine.unsqueeze(1) * y
        if tp_collectives():
            dist.all_reduce(y)
        return y


class ParallelLMHead(VocabParallelEmbedding):
    """Logits for the last token of every sequence (prefill) / every row (decode); with TP the
    vocab shards are gathered on rank 0, which is the only rank that samples."""

    def __init__(self, num_embeddings: int, embedding_dim: int, bias: bool = False):
        if bias:
            raise ValueError("lm_head bias is not part of the reference path")
        super().__init__(num_embeddings, embedding_dim)

    def local_logits(self, x):
        """This rank's vocabulary shard of the logits, no collective (the sharded sampling tail,
        Sampler.shard_keys, consumes it where it is)."""
        context = get_context()
        if context.is_prefill:
            x = x[(context.cu_seqlens_q[1:] - 1).long()].contiguous()
        return F.linear(x, self.weight)

    def forward(self, x):
        context = get_context()
        if context.is_prefill:
            x = x[(context.cu_seqlens_q[1:] - 1).long()].contiguous()
        logits = F.linear(x, self.weight)
        if self.tp_size > 1 and tp_collectives():
            parts = [torch.empty_like(logits) for _ in range(self.tp_size)] if self.tp_rank == 0 else None
            dist.gather(logits, parts, 0)
            logits = torch.cat(parts, -1) if self.tp_rank == 0 else None
        return logits
