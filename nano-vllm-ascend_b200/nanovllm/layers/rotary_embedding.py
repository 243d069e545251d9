"""NeoX rotary embedding (reference: layers/rotary_embedding.py).

The cos|sin table has the reference's layout ([max_position, 1, D] fp32, cos half then sin
half). ``forward(positions, q, k)`` rotates q and k in place with the K8 kernel (norm and
cache scatter disabled) and returns them; the decoder layers normally call
``ops.qknorm_rope_cache`` directly so that norm + rope + scatter are one launch.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


class RotaryEmbedding(nn.Module):

    def __init__(self, head_size: int, rotary_dim: int, max_position_embeddings: int, base: float):
        super().__init__()
        if rotary_dim != head_size:
            raise ValueError("partial rotary embeddings are not part of the reference path")
        self.head_size = head_size
        inv_freq = 1.0 / (base ** (torch.arange(0, rotary_dim, 2, dtype=torch.float) / rotary_dim))
        angles = torch.outer(torch.arange(max_position_embeddings, dtype=torch.float), inv_freq)
        table = torch.cat((angles.cos(), angles.sin()), dim=-1).unsqueeze_(1)
        self.register_buffer("cos_sin_cache", table.float(), persistent=False)

    def forward(self, positions, query, key):
        D = self.head_size
        for x in (query, key):
            flat = x.view(x.shape[0], -1)
            ops.qknorm_rope_cache(flat, positions, self.cos_sin_cache.view(-1, D), None, None, 0.0,
                                  None, None, None, flat.shape[1] // D, 0, D)
        return query, key


_rope_cache: dict = {}


def get_rope(head_size: int, rotary_dim: int, max_position: int, base: float):
    """One shared table per (shape, base), like the reference's lru_cache'd constructor."""
    key = (head_size, rotary_dim, max_position, float(base))
    if key not in _rope_cache:
        _rope_cache.clear()
        _rope_cache[key] = RotaryEmbedding(head_size, rotary_dim, max_position, base)
    return _rope_cache[key]


def get_rope_llama(head_size: int, rotary_dim: int, max_position: int, base: float):
    return RotaryEmbedding(head_size, rotary_dim, max_position, base)
