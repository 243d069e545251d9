"""Request state (reference: engine/sequence.py). Host-only bookkeeping."""
from __future__ import annotations

from array import array
from enum import Enum, auto
from itertools import count

from ..sampling_params import SamplingParams


class SequenceStatus(Enum):
    WAITING = auto()
    RUNNING = auto()
    FINISHED = auto()


class FinishReason(Enum):
    EOS = auto()
    LENGTH = auto()
    ABORTED = auto()
    PREEMPTED = auto()


class Sequence:
    _ids = count()

    def __init__(self, token_ids: list[int], sampling_params: SamplingParams | None = None,
                 request_id: str | None = None, block_size: int = 256):
        sp = sampling_params or SamplingParams()
        self.block_size = block_size
        self.seq_id = next(Sequence._ids)
        self.request_id = request_id
        self.status = SequenceStatus.WAITING
        self.token_ids = list(token_ids)
        self.last_token = self.token_ids[-1]
        self.num_tokens = len(self.token_ids)
        self.num_prompt_tokens = self.num_tokens
        self.num_cached_tokens = 0
        self.block_table: list[int] = []
        # chunked prefill only: prompt tokens whose K/V are in the cache, and the [start, end) slice
        # of the prompt the current prefill step computes (None = the whole prompt, reference behaviour)
        self.num_computed_tokens = 0
        self.prefill_window: tuple[int, int] | None = None
        self.temperature = sp.temperature
        self.max_tokens = sp.max_tokens
        self.ignore_eos = sp.ignore_eos
        self.greedy = sp.greedy
        self.finish_reason: FinishReason | None = None
        self.first_token_time: float | None = None   # set by the engine (TTFT)
        self.arrival_time: float | None = None

    def __len__(self):
        return self.num_tokens

    def __getitem__(self, key):
        return self.token_ids[key]

    @property
    def is_finished(self):
        return self.status == SequenceStatus.FINISHED

    @property
    def num_completion_tokens(self):
        return self.num_tokens - self.num_prompt_tokens

    @property
    def prompt_token_ids(self):
        return self.token_ids[:self.num_prompt_tokens]

    @property
    def completion_token_ids(self):
        return self.token_ids[self.num_prompt_tokens:]

    @property
    def num_cached_blocks(self):
        return self.num_cached_tokens // self.block_size

    @property
    def num_blocks(self):
        return (self.num_tokens + self.block_size - 1) // self.block_size

    @property
    def last_block_num_tokens(self):
        return self.num_tokens - (self.num_blocks - 1) * self.block_size

    def block(self, i: int) -> list[int]:
        assert 0 <= i < self.num_blocks
        return self.token_ids[i * self.block_size:(i + 1) * self.block_size]

    def append_token(self, token_id: int) -> None:
        self.token_ids.append(token_id)
        self.last_token = token_id
        self.num_tokens += 1

    # state shipped to tensor-parallel workers each step. Unlike the reference
    # (sequence.py:116-146) block_size and the sampling fields travel too,
    # otherwise num_blocks raises on the worker side.
    # A decode step needs only the last token (the reference makes the same cut, sequence.py:116-125);
    # the prompt travels as packed int32 bytes, and only while the sequence is still being prefilled.
    _WIRE = ("block_size", "seq_id", "num_tokens", "num_prompt_tokens", "num_cached_tokens",
             "last_token", "status", "temperature", "max_tokens",
             "ignore_eos", "greedy", "num_computed_tokens", "prefill_window")
    wire_decode = False          # set by ModelRunner.call around the pickling of a decode batch

    def __getstate__(self):
        state = {k: getattr(self, k) for k in self._WIRE}
        state["block_table"] = array("i", self.block_table).tobytes()
        state["token_ids"] = None if Sequence.wire_decode else array("i", self.token_ids).tobytes()
        return state

    def __setstate__(self, state):
        table, tokens = array("i"), array("i")
        table.frombytes(state.pop("block_table"))
        packed = state.pop("token_ids")
        if packed is not None:
            tokens.frombytes(packed)
        self.__dict__.update(state)
        self.block_table = table.tolist()
        self.token_ids = tokens.tolist()        # empty on decode steps: workers only read last_token
        self.request_id = None
        self.finish_reason = None
        self.first_token_time = self.arrival_time = None

    def __repr__(self):
        return f"Seq(id={self.seq_id}, status={self.status.name}, len={self.num_tokens})"
