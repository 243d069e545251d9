"""Qwen3 / Qwen2 dense decoders (reference: models/qwen3.py).

One class serves both families, as in the reference (qwen3.py:70-72,135): a config without
``attention_bias`` (Qwen2) gets qkv bias and no QK-norm, Qwen3 gets QK-norm and no bias.
"""
from __future__ import annotations

from .decoder import DecoderForCausalLM, DecoderSpec


def _rope_theta(config, default):
    """RoPE base. The reference reads ``getattr(config, "rope_theta", default)`` (qwen3.py:137,
    llama.py:133); transformers>=5 moved the value into ``config.rope_parameters``, where the
    reference no longer finds it and silently uses its default (right for Qwen3, wrong for
    Llama-3.x -- SURVEY.md 8c). Here the checkpoint's real value is honoured in both layouts."""
    theta = getattr(config, "rope_theta", None)
    if theta is None:
        params = getattr(config, "rope_parameters", None) or {}
        theta = params.get("rope_theta") if isinstance(params, dict) else None
    return float(theta if theta is not None else default)


def qwen3_spec(config) -> DecoderSpec:
    qkv_bias = bool(getattr(config, "attention_bias", True))
    head_dim = getattr(config, "head_dim", None) or config.hidden_size // config.num_attention_heads
    return DecoderSpec(
        vocab_size=config.vocab_size, hidden_size=config.hidden_size,
        num_layers=config.num_hidden_layers, num_heads=config.num_attention_heads,
        num_kv_heads=config.num_key_value_heads, head_dim=head_dim,
        intermediate_size=config.intermediate_size, rms_eps=config.rms_norm_eps,
        rope_theta=_rope_theta(config, 1000000), max_position=config.max_position_embeddings,
        qkv_bias=qkv_bias, o_bias=False, mlp_bias=False, qk_norm=not qkv_bias,
        tie_embeddings=bool(getattr(config, "tie_word_embeddings", False)),
        shared_rope=True, explicit_scale=False)


class Qwen3ForCausalLM(DecoderForCausalLM):

    def __init__(self, config):
        if getattr(config, "hidden_act", "silu") != "silu":
            raise ValueError("only silu MLPs are part of the reference path")
        super().__init__(qwen3_spec(config))
