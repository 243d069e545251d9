"""nanovllm on B200: the reference's front end (``LLM`` / ``SamplingParams``) over a
hand-written sm_100a paged-KV attention path (``libb200attn.so`` via ctypes)."""
from .sampling_params import SamplingParams

__all__ = ["LLM", "SamplingParams"]


def __getattr__(name):
    if name == "LLM":           # lazy: importing the layers alone must not pull the engine in
        from .llm import LLM
        return LLM
    raise AttributeError(name)
