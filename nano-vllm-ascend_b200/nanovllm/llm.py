from .engine.llm_engine import LLMEngine


class LLM(LLMEngine):
    """``LLM(model, **config_fields)`` -- the reference's public entry point (llm.py:4-6)."""
