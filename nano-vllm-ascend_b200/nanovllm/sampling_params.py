from dataclasses import dataclass


@dataclass
class SamplingParams:
    """Reference fields (sampling_params.py:4-11) + an explicit ``greedy`` switch: the
    reference forbids temperature 0 and always samples with multinomial; parity runs need
    argmax, so ``greedy=True`` selects it without touching the temperature contract."""
    temperature: float = 1.0
    max_tokens: int = 64
    ignore_eos: bool = False
    greedy: bool = False

    def __post_init__(self):
        assert self.temperature > 1e-10, "greedy sampling is not permitted (use greedy=True)"
