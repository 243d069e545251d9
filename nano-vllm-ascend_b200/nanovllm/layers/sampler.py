"""Sampling tail (reference: layers/sampler.py:9-17).

Reference semantics: trim padded graph rows, divide by temperature, softmax, multinomial.
Here the whole tail is one kernel over the bf16 logits: Gumbel-max sampling
draws from exactly softmax(logits / T) without materialising fp32 logits or probabilities;
``greedy=True`` (temperature <= 0 is still rejected by SamplingParams, as in the reference)
gives the argmax used by the parity tests.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


class Sampler(nn.Module):

    def __init__(self, seed: int = 0):
        super().__init__()
        self.seed = seed
        self.step = 0

    def _next_step(self, step_offset) -> int:
        """The noise is keyed by (seed, step + *step_offset, row, index). With a device counter the
        counter alone must be the key: a host step baked into a CUDA graph at capture differs per
        bucket and would make two launches (bucket A at counter r, bucket B at r') collide on the
        same effective step. Without one, the host counter advances per call."""
        if step_offset is not None:
            return 0
        self.step += 1
        return self.step

    def forward(self, logits: torch.Tensor, temperatures: torch.Tensor | None, greedy: bool = False,
                out: torch.Tensor | None = None, step_offset: torch.Tensor | None = None):
        """temperatures: [real_bs] fp32; rows with temperature <= 0 (or greedy=True) take the
        argmax. step_offset: device int64 counter (CUDA-graph replays advance it)."""
        real_bs = logits.shape[0] if temperatures is None else temperatures.shape[0]
        if logits.shape[0] != real_bs:
            logits = logits[:real_bs]
        return ops.sample(logits, None if greedy else temperatures, seed=self.seed,
                          step=self._next_step(step_offset), out=out, step_offset=step_offset)

    def shard_keys(self, logits: torch.Tensor, temperatures: torch.Tensor | None, index_offset: int,
                   greedy: bool = False, out: torch.Tensor | None = None,
                   step_offset: torch.Tensor | None = None):
        """Tensor parallelism without a logits gather: this rank's vocabulary shard -> one packed
        (score, global index) key per row; the caller MAX-all-reduces the keys over the ranks and
        decodes them with ``ops.decode_sample_keys``. Every rank must call this in lock step (the
        step counter is part of the random stream)."""
        real_bs = logits.shape[0] if temperatures is None else temperatures.shape[0]
        if logits.shape[0] != real_bs:
            logits = logits[:real_bs]
        return ops.sample_shard(logits, None if greedy else temperatures, seed=self.seed,
                                step=self._next_step(step_offset), index_offset=index_offset, out=out,
                                step_offset=step_offset)
