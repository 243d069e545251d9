"""Drop-in ``Attention`` operator: paged-KV scatter + prefill / decode attention on sm_100a.

Same surface as the reference's ``nanovllm/layers/attention.py:8-20,37``:
``Attention(num_heads, head_dim, scaling, num_kv_heads)``; ``k_cache`` / ``v_cache`` start
empty and are bound by the runner as ``[num_blocks, block_size, Hkv*D]`` views;
``forward(q, k, v)`` takes 2-D or 3-D q/k/v, reads the global step ``Context`` and returns a
2-D ``[T, H*D]`` tensor. Work is done by ``libb200attn`` (K1/K2 scatter, K4 prefill FA,
K6 paged decode) through the C ABI; there is no eager fallback.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from ..utils.context import get_context


class Attention(nn.Module):

    def __init__(self, num_heads, head_dim, scaling, num_kv_heads):
        super().__init__()
        self.num_heads = int(num_heads)
        self.num_kv_heads = int(num_kv_heads)
        self.head_dim = int(head_dim)
        self.scale = scaling if scaling is not None else 1.0 / (head_dim ** 0.5)
        self.block_size = 0
        self.layer_id = -1
        self.k_cache = torch.tensor([])
        self.v_cache = torch.tensor([])
        # only read when the runner binds e4m3 (uint8) caches: cache = e4m3(x / scale)
        self.k_scale = 1.0
        self.v_scale = 1.0

    @property
    def fp8_cache(self) -> bool:
        return self.k_cache.dtype in (torch.uint8, torch.float8_e4m3fn)

    def _as_heads(self, x, heads):
        return x.view(-1, heads, self.head_dim) if x.dim() == 2 else x

    def store_kvcache(self, k, v, context=None):
        """K1/K2: write this step's K/V rows at context.slot_mapping (flat or (block, offset))."""
        context = context or get_context()
        if self.k_cache.numel() and self.v_cache.numel():
            k, v = self._as_heads(k, self.num_kv_heads), self._as_heads(v, self.num_kv_heads)
            if self.fp8_cache:
                ops.reshape_and_cache_fp8(k, v, self.k_cache, self.v_cache, context.slot_mapping,
                                          self.k_scale, self.v_scale)
            else:
                ops.reshape_and_cache(k, v, self.k_cache, self.v_cache, context.slot_mapping)

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, kv_stored: bool = False):
        """kv_stored=True: the caller already scattered K/V (fused K8 epilogue)."""
        context = get_context()
        self.block_size = context.block_size
        q = self._as_heads(q, self.num_heads)
        k = self._as_heads(k, self.num_kv_heads)
        v = self._as_heads(v, self.num_kv_heads)
        if not kv_stored:
            self.store_kvcache(k, v, context)
        if context.is_prefill:
            if context.paged_prefill:
                if self.fp8_cache:
                    # the tcgen05 kernel multiplies 16-bit operands: gather + de-quantise the cached
                    # tokens into packed K/V first (this step's own rows were just quantised into the
                    # cache, so every key is read at cache precision, like the decode steps that follow)
                    k, v = ops.gather_dequant_fp8(self.k_cache, self.v_cache, context.block_tables,
                                                  context.cu_seqlens_k, context.max_seqlen_k,
                                                  self.num_kv_heads, self.head_dim, self.k_scale,
                                                  self.v_scale, dtype=q.dtype)
                    return ops.prefill_varlen_attn(q, k, v, context.cu_seqlens_q, context.cu_seqlens_k,
                                                   context.max_seqlen_q, context.max_seqlen_k, self.scale,
                                                   self.num_kv_heads)
                return ops.prefill_varlen_attn(q, self.k_cache, self.v_cache, context.cu_seqlens_q,
                                               context.cu_seqlens_k, context.max_seqlen_q,
                                               context.max_seqlen_k, self.scale, self.num_kv_heads,
                                               block_tables=context.block_tables)
            return ops.prefill_varlen_attn(q, k, v, context.cu_seqlens_q, context.cu_seqlens_k,
                                           context.max_seqlen_q, context.max_seqlen_k, self.scale,
                                           self.num_kv_heads)
        if self.fp8_cache:
            return ops.paged_decode_attn_fp8(q, self.k_cache, self.v_cache, context.block_tables,
                                             context.context_lens, self.scale, self.num_kv_heads,
                                             self.k_scale, self.v_scale, num_ctas=context.decode_attn_ctas)
        return ops.paged_decode_attn(q, self.k_cache, self.v_cache, context.block_tables,
                                     context.context_lens, self.scale, self.num_kv_heads,
                                     num_ctas=context.decode_attn_ctas)
