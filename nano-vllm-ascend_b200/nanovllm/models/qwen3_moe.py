"""Qwen3-MoE decoders (reference: models/qwen3_moe.py): the Qwen3 attention stack (QK-norm, no
bias unless ``attention_bias``) with sparse mixture-of-experts MLPs on the layers the config
selects (:207-219). Same parameter names and packed-module mapping as the reference, so its
checkpoints load unchanged; the MoE block itself is layers/moe.py on csrc/moe.cu.
"""
from __future__ import annotations

from dataclasses import replace

from .decoder import DecoderForCausalLM
from .qwen3 import qwen3_spec


def qwen3_moe_spec(config):
    spec = qwen3_spec(config)
    # the reference builds q_norm/k_norm unconditionally (qwen3_moe.py:73-74) and reads the bias flag
    # with a False default (:203)
    bias = bool(getattr(config, "attention_bias", False))
    return replace(
        spec, qkv_bias=bias, qk_norm=True, explicit_scale=True,
        num_experts=int(getattr(config, "num_experts", 0)),
        top_k=int(getattr(config, "num_experts_per_tok", 0)),
        moe_intermediate_size=int(getattr(config, "moe_intermediate_size", 0)),
        decoder_sparse_step=int(getattr(config, "decoder_sparse_step", 1)),
        mlp_only_layers=tuple(getattr(config, "mlp_only_layers", ()) or ()))


class Qwen3MoeForCausalLM(DecoderForCausalLM):

    def __init__(self, config):
        if getattr(config, "hidden_act", "silu") != "silu":
            raise ValueError("only silu experts are part of the reference path (qwen3_moe.py:112)")
        super().__init__(qwen3_moe_spec(config))
