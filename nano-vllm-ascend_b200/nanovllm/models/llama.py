"""Llama dense decoders (reference: models/llama.py). rope_scaling is accepted and ignored,
as in the reference (llama.py:27,65-70)."""
from __future__ import annotations

from .decoder import DecoderForCausalLM, DecoderSpec
from .qwen3 import _rope_theta


def llama_spec(config) -> DecoderSpec:
    bias = bool(getattr(config, "attention_bias", False) or getattr(config, "bias", False))
    head_dim = getattr(config, "head_dim", None) or config.hidden_size // config.num_attention_heads
    return DecoderSpec(
        vocab_size=config.vocab_size, hidden_size=config.hidden_size,
        num_layers=config.num_hidden_layers, num_heads=config.num_attention_heads,
        num_kv_heads=getattr(config, "num_key_value_heads", config.num_attention_heads),
        head_dim=head_dim, intermediate_size=config.intermediate_size,
        rms_eps=config.rms_norm_eps, rope_theta=_rope_theta(config, 10000),
        max_position=getattr(config, "max_position_embeddings", 8192), qkv_bias=bias, o_bias=bias,
        mlp_bias=bool(getattr(config, "mlp_bias", False)), qk_norm=False,
        tie_embeddings=bool(getattr(config, "tie_word_embeddings", False)),
        shared_rope=False, explicit_scale=True)


class LlamaForCausalLM(DecoderForCausalLM):

    def __init__(self, config):
        super().__init__(llama_spec(config))
