"""HF architecture name -> model class (reference: models/models_map.py). MoE, MiniCPM and the
VL stack are outside the hot-path scope."""
from .llama import LlamaForCausalLM
from .qwen3 import Qwen3ForCausalLM

model_dict = {
    "LlamaForCausalLM": LlamaForCausalLM,
    "Qwen2ForCausalLM": Qwen3ForCausalLM,
    "Qwen3ForCausalLM": Qwen3ForCausalLM,
}
