"""Per-step metadata channel between the runner and the attention / LM-head layers.

Field names and meaning are the reference's contract (``nanovllm/utils/context.py:5-37``):
the runner publishes the step's index tensors once, every layer reads them with
``get_context()``; nothing else is passed "around" the model graph.

  is_prefill      prefill step (varlen packed tokens) vs decode step (1 token / sequence)
  cu_seqlens_q/k  int32 [num_seqs+1] cumulative lengths (prefill)
  max_seqlen_q/k  python ints (prefill)
  slot_mapping    int32: flat [T] slots (prefill) or [B,2] (block, offset) pairs (decode),
                  exactly as the reference runner builds them (model_runner.py:278,301,353)
  context_lens    int32 [B] tokens to attend per sequence, 0 for padded graph rows
  block_tables    int32 [B, max_blocks], -1 padded
  is_enforce_eager, real_bs, block_size   as in the reference
One addition: ``paged_prefill`` -- True when the prefill step feeds only the uncached suffix
and attention must read K/V through block_tables (true prefix hit, attention_ori.py:65-66);
the reference runner at this commit always recomputes, so it defaults to False.
Second addition: ``decode_attn_ctas`` -- persistent grid size of the paged decode kernel for this step
(0 = one CTA per SM); the runner's two-micro-batch decode step sets it so that the attention of one
half leaves SMs to the GEMMs of the other (engine/model_runner.py). Third: ``own_gemm`` -- the
linear layers of this step run on the library's single-CTA tcgen05 GEMM (ops.linear_umma) instead of
cuBLAS, whose 2-CTA clustered kernels cannot be placed next to a resident attention kernel.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass
class Context:
    is_prefill: bool = False
    cu_seqlens_q: torch.Tensor | None = None
    cu_seqlens_k: torch.Tensor | None = None
    max_seqlen_q: int = 0
    max_seqlen_k: int = 0
    slot_mapping: torch.Tensor | None = None
    context_lens: torch.Tensor | None = None
    block_tables: torch.Tensor | None = None
    is_enforce_eager: bool = True
    real_bs: int = -1
    block_size: int = 256
    paged_prefill: bool = False
    decode_attn_ctas: int = 0
    own_gemm: bool = False


_current = Context()


def get_context() -> Context:
    return _current


def set_context(is_prefill, cu_seqlens_q=None, cu_seqlens_k=None, max_seqlen_q=0, max_seqlen_k=0,
                slot_mapping=None, context_lens=None, block_tables=None, is_enforce_eager=None,
                real_bs=None, block_size=None, paged_prefill=False, decode_attn_ctas=0, own_gemm=False) -> None:
    global _current
    _current = Context(is_prefill, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                       slot_mapping, context_lens, block_tables, is_enforce_eager, real_bs,
                       block_size, paged_prefill, decode_attn_ctas, own_gemm)


def reset_context() -> None:
    global _current
    _current = Context()
