"""Which decoder class serves which HF ``architectures`` entry.

Only the dense families on the hot-path scope are registered: the reference's
registry (models/models_map.py) also lists its MoE, MiniCPM and vision-language stacks, which this
build deliberately leaves out. Qwen2 checkpoints are served by the Qwen3 class, as in the reference:
the config's ``attention_bias`` decides between QKV bias (Qwen2) and QK-norm (Qwen3).
"""
from __future__ import annotations

from .llama import LlamaForCausalLM
from .qwen3 import Qwen3ForCausalLM

model_dict: dict[str, type] = {}


def register(architecture: str, cls: type) -> None:
    model_dict[architecture] = cls


for _arch, _cls in (("Qwen3ForCausalLM", Qwen3ForCausalLM),
                    ("Qwen2ForCausalLM", Qwen3ForCausalLM),
                    ("LlamaForCausalLM", LlamaForCausalLM)):
    register(_arch, _cls)
