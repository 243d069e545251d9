"""Per-request sampling controls of the public front end.

Same three knobs as the reference front end (``sampling_params.py:4-11`` there): temperature,
max_tokens, ignore_eos, with the same defaults and the same rule that a (near-)zero temperature
is refused. One addition: ``greedy`` selects the argmax of the logits explicitly -- parity runs
need it, and the reference offers no way to ask for it (its sampler always draws with
torch.multinomial).
"""
from __future__ import annotations

from dataclasses import asdict, dataclass

_MIN_TEMPERATURE = 1e-10


@dataclass
class SamplingParams:
    temperature: float = 1.0     # softmax temperature; must be > 1e-10
    max_tokens: int = 64         # completion length limit
    ignore_eos: bool = False     # keep generating past the EOS token
    greedy: bool = False         # take the argmax instead of sampling (temperature is then unused)

    def __post_init__(self) -> None:
        if not self.temperature > _MIN_TEMPERATURE:
            raise AssertionError("greedy sampling is not permitted through temperature=0; "
                                 "pass greedy=True instead")
        if self.max_tokens < 1:
            raise AssertionError("max_tokens must be at least 1")

    def to_dict(self) -> dict:
        return asdict(self)
