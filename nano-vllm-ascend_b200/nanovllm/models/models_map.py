"""Which decoder class serves which HF ``architectures`` entry.

The dense families of the hot-path scope and the Qwen3-MoE stack are registered;
the reference's registry (models/models_map.py) also lists its MiniCPM and vision-language stacks,
which this build deliberately leaves out. Qwen2 checkpoints are served by the Qwen3 class, as in the reference:
the config's ``attention_bias`` decides between QKV bias (Qwen2) and QK-norm (Qwen3).
"""
from __future__ import annotations

from .llama import LlamaForCausalLM
from .qwen3 import Qwen3ForCausalLM
from .qwen3_moe import Qwen3MoeForCausalLM

model_dict: dict[str, type] = {}


def register(architecture: str, cls: type) -> None:
    model_dict[architecture] = cls


for _arch, _cls in (("Qwen3ForCausalLM", Qwen3ForCausalLM),
                    ("Qwen2ForCausalLM", Qwen3ForCausalLM),
                    ("Qwen3MoeForCausalLM", Qwen3MoeForCausalLM),
                    ("LlamaForCausalLM", LlamaForCausalLM)):
    register(_arch, _cls)
