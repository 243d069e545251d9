"""Dense decoder stack shared by the Qwen3/Qwen2 and Llama front ends.

Module tree and parameter names are the reference's (models/qwen3.py, models/llama.py):
``model.embed_tokens``, ``model.layers.N.{input_layernorm,post_attention_layernorm}``,
``self_attn.{qkv_proj,o_proj,q_norm,k_norm,attn}``, ``mlp.{gate_up_proj,down_proj}``,
``model.norm``, ``lm_head`` -- so checkpoints, the packed-module mapping and the runner's
"bind every module that has k_cache & v_cache" scan (model_runner.py:223-229) work unchanged.

What differs is the launch structure of a layer. Reference (qwen3.py:79-88,148-161), eager:
  qkv GEMM, split, q_norm (5 kernels), k_norm (5), rope (~10), scatter (2), attention,
  o GEMM, add+norm (7), gate_up GEMM, silu*mul (3), down GEMM.
Here: qkv GEMM -> ONE fused K8 kernel (norm + rope + KV scatter, in place) -> attention ->
o GEMM -> fused add+RMSNorm -> gate_up GEMM -> fused SiluAndMul -> down GEMM.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from .. import ops
from ..layers.activation import SiluAndMul
from ..layers.attention import Attention
from ..layers.embed_head import ParallelLMHead, VocabParallelEmbedding
from ..layers.layernorm import RMSNorm
from ..layers.linear import (MergedColumnParallelLinear, QKVParallelLinear, RowParallelLinear,
                             tp_rank_size)
from ..layers.rotary_embedding import RotaryEmbedding, get_rope
from ..utils.context import get_context


@dataclass
class DecoderSpec:
    """The architecture facts the stack needs, extracted from an HF config by the front ends."""
    vocab_size: int
    hidden_size: int
    num_layers: int
    num_heads: int
    num_kv_heads: int
    head_dim: int
    intermediate_size: int
    rms_eps: float
    rope_theta: float
    max_position: int
    qkv_bias: bool
    o_bias: bool
    mlp_bias: bool
    qk_norm: bool
    tie_embeddings: bool
    shared_rope: bool          # Qwen3 shares one table (lru_cache); Llama builds one per layer
    explicit_scale: bool       # Llama passes head_dim**-0.5; Qwen3 passes None (same value)
    # sparse MoE MLPs (models/qwen3_moe.py:207-219): layer i is sparse when num_experts > 0,
    # i is not in mlp_only_layers and (i + 1) % decoder_sparse_step == 0; dense layers use intermediate_size
    num_experts: int = 0
    top_k: int = 0
    moe_intermediate_size: int = 0
    decoder_sparse_step: int = 1
    mlp_only_layers: tuple = ()

    def is_sparse_layer(self, layer_idx: int) -> bool:
        return (self.num_experts > 0 and layer_idx not in self.mlp_only_layers
                and (layer_idx + 1) % self.decoder_sparse_step == 0)


class DecoderAttention(nn.Module):

    def __init__(self, spec: DecoderSpec, rope: RotaryEmbedding):
        super().__init__()
        _, tp = tp_rank_size()
        if spec.num_heads % tp or (spec.num_kv_heads % tp and tp % spec.num_kv_heads):
            raise ValueError("attention heads must divide by the tensor-parallel size")
        self.num_heads = spec.num_heads // tp
        # fewer kv heads than ranks (Qwen3-30B-A3B: 4 kv heads, TP=8): every kv head is replicated on
        # tp / Hkv consecutive ranks, each rank keeps one (an extension; the reference asserts
        # Hkv % tp == 0, qwen3_moe.py:38-39)
        self.num_kv_heads = max(1, spec.num_kv_heads // tp)
        self.head_dim = spec.head_dim
        self.q_size = self.num_heads * self.head_dim
        self.kv_size = self.num_kv_heads * self.head_dim
        self.eps = spec.rms_eps
        self.qkv_proj = QKVParallelLinear(spec.hidden_size, self.head_dim, spec.num_heads,
                                          spec.num_kv_heads, bias=spec.qkv_bias)
        self.o_proj = RowParallelLinear(spec.num_heads * self.head_dim, spec.hidden_size,
                                        bias=spec.o_bias)
        self.rotary_emb = rope
        self.attn = Attention(self.num_heads, self.head_dim,
                              self.head_dim ** -0.5 if spec.explicit_scale else None,
                              self.num_kv_heads)
        if spec.qk_norm:
            self.q_norm = RMSNorm(self.head_dim, eps=spec.rms_eps)
            self.k_norm = RMSNorm(self.head_dim, eps=spec.rms_eps)
        else:
            self.q_norm = self.k_norm = None

    def forward(self, positions: torch.Tensor, hidden_states: torch.Tensor) -> torch.Tensor:
        qkv = self.qkv_proj(hidden_states)
        ctx = get_context()
        attn = self.attn
        have_cache = attn.k_cache.numel() > 0
        fused_scatter = have_cache and not attn.fp8_cache     # K8 scatters 16-bit caches itself
        ops.qknorm_rope_cache(
            qkv, positions, self.rotary_emb.cos_sin_cache.view(-1, self.head_dim),
            None if self.q_norm is None else self.q_norm.weight,
            None if self.k_norm is None else self.k_norm.weight, self.eps,
            attn.k_cache if fused_scatter else None, attn.v_cache if fused_scatter else None,
            ctx.slot_mapping if fused_scatter else None, self.num_heads, self.num_kv_heads,
            self.head_dim)
        q = qkv[:, :self.q_size].unflatten(1, (self.num_heads, self.head_dim))
        k = qkv[:, self.q_size:self.q_size + self.kv_size].unflatten(1, (self.num_kv_heads, self.head_dim))
        v = qkv[:, self.q_size + self.kv_size:].unflatten(1, (self.num_kv_heads, self.head_dim))
        o = attn(q, k, v, kv_stored=not attn.fp8_cache)   # fp8 cache: attn runs the quantising scatter
        return self.o_proj(o)


class DecoderMLP(nn.Module):

    def __init__(self, spec: DecoderSpec):
        super().__init__()
        self.gate_up_proj = MergedColumnParallelLinear(spec.hidden_size, [spec.intermediate_size] * 2,
                                                       bias=spec.mlp_bias)
        self.down_proj = RowParallelLinear(spec.intermediate_size, spec.hidden_size, bias=spec.mlp_bias)
        self.act_fn = SiluAndMul()

    def forward(self, x):
        w = self.gate_up_proj.weight
        if (get_context().own_gemm and self.gate_up_proj.bias is None and x.dim() == 2 and x.is_contiguous()
                and w.shape[0] % 128 == 0 and w.shape[1] % 64 == 0):
            return self.down_proj(ops.gate_up_silu_umma(x, w))      # gate|up GEMM + SiluAndMul, one launch
        return self.down_proj(self.act_fn(self.gate_up_proj(x)))


class DecoderLayer(nn.Module):

    def __init__(self, spec: DecoderSpec, rope: RotaryEmbedding, layer_idx: int = 0):
        super().__init__()
        self.self_attn = DecoderAttention(spec, rope)
        if spec.is_sparse_layer(layer_idx):
            from ..layers.moe import SparseMoeBlock
            self.mlp = SparseMoeBlock(spec.hidden_size, spec.moe_intermediate_size, spec.num_experts,
                                      spec.top_k)
        else:
            self.mlp = DecoderMLP(spec)
        self.input_layernorm = RMSNorm(spec.hidden_size, eps=spec.rms_eps)
        self.post_attention_layernorm = RMSNorm(spec.hidden_size, eps=spec.rms_eps)
        # both row-parallel outputs of a layer feed a residual-add RMSNorm (post_attention_layernorm,
        # then the next layer's input_layernorm / the final norm): their all-reduce may be fused there
        self.self_attn.o_proj.defer_allreduce = True
        if isinstance(self.mlp, DecoderMLP):
            self.mlp.down_proj.defer_allreduce = True
        else:
            self.mlp.defer_allreduce = True

    def forward(self, positions, hidden_states, residual):
        if residual is None:
            residual = hidden_states
            hidden_states = self.input_layernorm(hidden_states)
        else:
            hidden_states, residual = self.input_layernorm(hidden_states, residual)
        hidden_states = self.self_attn(positions, hidden_states)
        hidden_states, residual = self.post_attention_layernorm(hidden_states, residual)
        hidden_states = self.mlp(hidden_states)
        return hidden_states, residual


class DecoderModel(nn.Module):

    def __init__(self, spec: DecoderSpec):
        super().__init__()
        self.embed_tokens = VocabParallelEmbedding(spec.vocab_size, spec.hidden_size)
        shared = get_rope(spec.head_dim, spec.head_dim, spec.max_position, spec.rope_theta) \
            if spec.shared_rope else None
        self.layers = nn.ModuleList([
            DecoderLayer(spec, shared if shared is not None else
                         RotaryEmbedding(spec.head_dim, spec.head_dim, spec.max_position, spec.rope_theta), i)
            for i in range(spec.num_layers)])
        self.norm = RMSNorm(spec.hidden_size, eps=spec.rms_eps)

    def forward(self, input_ids, positions):
        hidden_states = self.embed_tokens(input_ids)
        residual = None
        for layer in self.layers:
            hidden_states, residual = layer(positions, hidden_states, residual)
        hidden_states, _ = self.norm(hidden_states, residual)
        return hidden_states


class DecoderForCausalLM(nn.Module):
    packed_modules_mapping = {
        "q_proj": ("qkv_proj", "q"),
        "k_proj": ("qkv_proj", "k"),
        "v_proj": ("qkv_proj", "v"),
        "gate_proj": ("gate_up_proj", 0),
        "up_proj": ("gate_up_proj", 1),
    }

    def __init__(self, spec: DecoderSpec):
        super().__init__()
        self.spec = spec
        self.model = DecoderModel(spec)
        self.lm_head = ParallelLMHead(spec.vocab_size, spec.hidden_size)
        if spec.tie_embeddings:
            self.lm_head.weight.data = self.model.embed_tokens.weight.data

    def forward(self, input_ids, positions):
        return self.model(input_ids, positions)

    def virtual_param_shapes(self) -> dict[str, tuple[int, ...]]:
        """Parameters of the full model that live on OTHER ranks only (experts under expert
        parallelism); see utils/loader.py."""
        out = {}
        for i, layer in enumerate(self.model.layers):
            fn = getattr(layer.mlp, "nonlocal_param_shapes", None)
            if fn is not None:
                out.update(fn(f"model.layers.{i}.mlp."))
        return out

    def compute_logits(self, hidden_states):
        return self.lm_head(hidden_states)
