"""Front-end engine (reference: engine/llm_engine.py): request intake, the schedule -> run ->
postprocess loop, tensor-parallel worker start-up, and the ``generate`` convenience API."""
from __future__ import annotations

import atexit
from dataclasses import fields
from time import perf_counter

import torch.multiprocessing as mp

from ..config import Config
from ..sampling_params import SamplingParams
from .model_runner import ModelRunner
from .scheduler import Scheduler
from .sequence import Sequence


class LLMEngine:

    def __init__(self, model, **kwargs):
        allowed = {f.name for f in fields(Config)}
        config = Config(model, **{k: v for k, v in kwargs.items() if k in allowed})
        self.config = config
        self.block_size = config.kvcache_block_size
        self.ps, self.events = [], []
        if config.tensor_parallel_size > 1 and config.distributed_launch == "spawn":
            ctx = mp.get_context("spawn")
            for rank in range(1, config.tensor_parallel_size):
                event = ctx.Event()
                proc = ctx.Process(target=ModelRunner, args=(config, rank, event))
                proc.start()
                self.ps.append(proc)
                self.events.append(event)
        rank = 0
        if config.distributed_launch == "env" and config.tensor_parallel_size > 1:
            import os
            rank = int(os.environ.get("RANK", 0)) % config.tensor_parallel_size
        self.model_runner = ModelRunner(config, rank, self.events)
        self.tokenizer = None
        self._tokenizer_tried = False
        self.scheduler = Scheduler(config)
        self._exited = False
        atexit.register(self.exit)

    # ------------------------------------------------------------------ tokenizer (optional)
    def _get_tokenizer(self):
        if not self._tokenizer_tried:
            self._tokenizer_tried = True
            import os
            if os.path.isdir(self.config.model):
                try:
                    from transformers import AutoTokenizer
                    self.tokenizer = AutoTokenizer.from_pretrained(self.config.model, use_fast=True)
                    if self.tokenizer.eos_token_id is not None:
                        self.config.eos = self.scheduler.eos = self.tokenizer.eos_token_id
                except Exception:
                    self.tokenizer = None
        return self.tokenizer

    def exit(self):
        if self._exited:
            return
        self._exited = True
        self.model_runner.call("exit")
        del self.model_runner
        for p in self.ps:
            p.join()

    # ------------------------------------------------------------------ requests
    def add_request(self, prompt: str | list[int], sampling_params: SamplingParams,
                    request_id: str | None = None) -> Sequence:
        if isinstance(prompt, str):
            tok = self._get_tokenizer()
            if tok is None:
                raise ValueError("string prompts need a tokenizer (checkpoint directory); "
                                 "pass token-id lists for preset / random-init models")
            prompt = tok.encode(prompt)
        seq = Sequence(prompt, sampling_params, request_id=request_id, block_size=self.block_size)
        seq.arrival_time = perf_counter()
        self.scheduler.add(seq)
        return seq

    def abort_request(self, request_id: str) -> None:
        self.scheduler.abort_seq_group(request_id)

    def is_finished(self) -> bool:
        return self.scheduler.is_finished()

    def step(self):
        seqs, is_prefill = self.scheduler.schedule()
        if not seqs:
            # nothing runnable: either idle, or the head of the queue can never be admitted (the
            # reference spins forever here, scheduler.py:41-77 returning an empty batch each step)
            if self.scheduler.waiting and not self.scheduler.running:
                seq = self.scheduler.waiting[0]
                raise RuntimeError(
                    f"sequence of {len(seq)} tokens needs {seq.num_blocks} KV blocks but the pool "
                    f"has {len(self.scheduler.block_manager.blocks)}: raise gpu_memory_utilization / "
                    f"num_kvcache_blocks or lower max_model_len")
            return [], 0
        token_ids = self.model_runner.call("run", seqs, is_prefill)
        num_tokens = -len(seqs)
        if is_prefill:
            now = perf_counter()
            num_tokens = 0
            for seq in seqs:
                start, end = seq.prefill_window or (0, len(seq))
                num_tokens += end - start
                if seq.first_token_time is None and end == len(seq):
                    seq.first_token_time = now
        self.scheduler.postprocess(seqs, token_ids)
        outputs = [(s.seq_id, s.completion_token_ids, s.num_prompt_tokens, s.num_cached_tokens)
                   for s in seqs if s.is_finished]
        return outputs, num_tokens

    def generate(self, prompts, sampling_params, use_tqdm: bool = True) -> list[dict]:
        pbar = None
        if use_tqdm:
            from tqdm.auto import tqdm
            pbar = tqdm(total=len(prompts), desc="Generating", dynamic_ncols=True)
        if not isinstance(sampling_params, list):
            sampling_params = [sampling_params] * len(prompts)
        for prompt, sp in zip(prompts, sampling_params):
            self.add_request(prompt, sp)
        done = {}
        prefill_tps = decode_tps = 0.0
        while not self.is_finished():
            t0 = perf_counter()
            finished, num_tokens = self.step()
            dt = perf_counter() - t0
            if num_tokens > 0:
                prefill_tps = num_tokens / dt
            else:
                decode_tps = -num_tokens / dt
            if pbar is not None:
                pbar.set_postfix({"Prefill": f"{int(prefill_tps)}tok/s", "Decode": f"{int(decode_tps)}tok/s"})
            for seq_id, token_ids, prompt_len, cached in finished:
                done[seq_id] = (token_ids, prompt_len, cached)
                if pbar is not None:
                    pbar.update(1)
        if pbar is not None:
            pbar.close()
        tok = self._get_tokenizer()
        return [{"text": tok.decode(ids) if tok is not None else "", "token_ids": ids,
                 "prompt_len": plen, "cache_tokens": cached}
                for _, (ids, plen, cached) in sorted(done.items())]
