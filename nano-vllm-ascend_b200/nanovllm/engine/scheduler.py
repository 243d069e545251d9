"""Prefill-first continuous batching with preemption (reference: engine/scheduler.py)."""
from __future__ import annotations

from collections import deque

from ..config import Config
from .block_manager import BlockManager
from .sequence import FinishReason, Sequence, SequenceStatus


class Scheduler:

    def __init__(self, config: Config):
        self.max_num_seqs = config.max_num_seqs
        self.max_num_batched_tokens = config.max_num_batched_tokens
        self.max_model_len = config.max_model_len
        self.eos = config.eos
        self.chunked_prefill = getattr(config, "chunked_prefill", False)
        # the last physical block is the dummy slot padded graph rows write to
        # (reference scheduler.py:26-30, model_runner.py:310)
        self.block_manager = BlockManager(config.num_kvcache_blocks - 1, config.kvcache_block_size)
        self.waiting: deque[Sequence] = deque()
        self.running: deque[Sequence] = deque()

    def is_finished(self) -> bool:
        return not self.waiting and not self.running

    def add(self, seq: Sequence) -> None:
        self.waiting.append(seq)

    def schedule(self) -> tuple[list[Sequence], bool]:
        if self.chunked_prefill:
            batch = self._schedule_prefill_chunks()
            if batch:
                return batch, True
            return self._schedule_decode(), False
        batch: list[Sequence] = []
        tokens = 0
        # 1) admit waiting prompts while the token budget and free blocks allow
        while self.waiting and len(batch) < self.max_num_seqs:
            seq = self.waiting[0]
            if tokens + len(seq) > self.max_num_batched_tokens or not self.block_manager.can_allocate(seq):
                break
            self.waiting.popleft()
            self.block_manager.allocate(seq)
            seq.status = SequenceStatus.RUNNING
            self.running.append(seq)
            batch.append(seq)
            tokens += len(seq) - seq.num_cached_tokens
        if batch:
            return batch, True
        return self._schedule_decode(), False

    def _schedule_prefill_chunks(self) -> list[Sequence]:
        """Chunked variant of step 1 (an extension; the reference admits whole prompts only). A prompt
        takes its blocks on admission and is then computed ``prefill_window`` by window: it stays at
        the head of ``waiting`` until the window that reaches its end has been scheduled, so nothing
        can be admitted behind it -- and hit its not-yet-computed blocks -- before that step."""
        batch: list[Sequence] = []
        tokens = 0
        while self.waiting and len(batch) < self.max_num_seqs:
            seq = self.waiting[0]
            if not seq.block_table:
                if not self.block_manager.can_allocate(seq):
                    break
                self.block_manager.allocate(seq)
                seq.num_computed_tokens = 0
            n = len(seq)
            start = max(seq.num_computed_tokens, min(seq.num_cached_tokens, n - 1))
            room = self.max_num_batched_tokens - tokens
            if room <= 0:
                break
            end = min(n, start + room)
            seq.prefill_window = (start, end)
            batch.append(seq)
            tokens += end - start
            if end < n:
                break                       # budget exhausted; the prompt continues next step
            self.waiting.popleft()
            seq.status = SequenceStatus.RUNNING
            self.running.append(seq)
        return batch

    def _schedule_decode(self) -> list[Sequence]:
        batch: list[Sequence] = []
        block_size = self.block_manager.block_size
        # 2) one decode step for as many running sequences as fit; evict from the back on OOM
        while self.running and len(batch) < self.max_num_seqs:
            seq = self.running.popleft()
            if seq.num_tokens % block_size > 1:
                # 14 steps out of 16 at block size 16: the newest token neither opened nor completed a
                # block, so can_append is trivially true and may_append has nothing to do
                batch.append(seq)
                continue
            while not self.block_manager.can_append(seq):
                if self.running:
                    self.preempt(self.running.pop())
                else:
                    self.preempt(seq)
                    seq = None
                    break
            if seq is not None:
                self.block_manager.may_append(seq)
                batch.append(seq)
        if batch:
            self.running.extendleft(reversed(batch))
        return batch

    def preempt(self, seq: Sequence) -> None:
        seq.num_computed_tokens, seq.prefill_window = 0, None
        seq.status = SequenceStatus.WAITING
        seq.finish_reason = FinishReason.PREEMPTED
        self.block_manager.deallocate(seq)
        self.waiting.appendleft(seq)

    def abort_seq_group(self, request_id: str) -> None:
        for queue in (self.waiting, self.running):
            for seq in [s for s in queue if s.request_id == request_id]:
                queue.remove(seq)
                self.free_seq(seq, FinishReason.ABORTED)

    def free_seq(self, seq: Sequence, reason: FinishReason) -> None:
        # still WAITING with blocks = a (re-)prefill in progress (chunked mode keeps the sequence at the
        # head of ``waiting`` until its last window ran): blocks past num_computed_tokens carry
        # hashes but no K/V yet, whether or not the sequence already had completions
        mid_prefill = seq.status == SequenceStatus.WAITING and bool(seq.block_table)
        seq.status = SequenceStatus.FINISHED
        seq.finish_reason = reason
        computed = max(seq.num_computed_tokens, seq.num_cached_tokens)   # prefix hits hold K/V already
        if mid_prefill and computed < len(seq):
            self.block_manager.forget_uncomputed(seq, computed)
        self.block_manager.deallocate(seq)

    def postprocess(self, seqs: list[Sequence], token_ids: list[int]) -> None:
        eos, max_model_len = self.eos, self.max_model_len
        for seq, token_id in zip(seqs, token_ids):
            window = seq.prefill_window
            if window is not None:
                seq.prefill_window = None
                seq.num_computed_tokens = window[1]
                if window[1] < seq.num_tokens:
                    continue                # a non-final prompt chunk: its sampled token is meaningless
            seq.append_token(token_id)
            n = seq.num_tokens
            hit_eos = not seq.ignore_eos and token_id == eos
            out_of_room = n - seq.num_prompt_tokens == seq.max_tokens or n == max_model_len
            if hit_eos or out_of_room:
                self.free_seq(seq, FinishReason.EOS if hit_eos else FinishReason.LENGTH)
                self.running.remove(seq)
