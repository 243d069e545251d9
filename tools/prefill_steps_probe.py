"""Developer probe: wall time of every prefill step of the bench workload (256 x 1024-token prompts,
16 sequences per step), synchronised per step, to separate one-time costs from the steady state."""
import os
import random
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nano-vllm-ascend_b200"))
from nanovllm import LLM, SamplingParams  # noqa: E402

llm = LLM("qwen3-0.6b", kvcache_block_size=16, max_num_seqs=256, max_model_len=4096,
          max_num_batched_tokens=16384, gpu_memory_utilization=0.5, seed=0, graph_batch_sizes=(256,))
rnd = random.Random(0)
sp = SamplingParams(temperature=0.6, ignore_eos=True, max_tokens=4)
for rep in range(2):
    for _ in range(256):
        llm.add_request([rnd.randint(0, 10000) for _ in range(1024)], sp)
    times = []
    while llm.scheduler.waiting:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        llm.step()
        torch.cuda.synchronize()
        times.append((time.perf_counter() - t0) * 1e3)
    print(f"rep {rep}: prefill step ms:", " ".join(f"{t:.1f}" for t in times), flush=True)
    while not llm.is_finished():
        llm.step()
llm.exit()
