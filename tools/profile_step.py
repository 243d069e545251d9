"""Profiling harness (run under ncu with --profile-from-start off): builds the config-2 engine,
prefills 256 x 1024-token prompts, then brackets N decode steps (or one prefill step) with
cudaProfilerStart/Stop so the capture holds only the timed-region kernels."""
import argparse
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nano-vllm-ascend_b200"))
from nanovllm import LLM, SamplingParams  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--phase", default="decode", choices=["decode", "prefill"])
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--ctx", type=int, default=1024)
ap.add_argument("--model", default="qwen3-0.6b")
args = ap.parse_args()

llm = LLM(args.model, kvcache_block_size=16, max_num_seqs=args.batch, max_model_len=4096,
          max_num_batched_tokens=16384, seed=0, graph_batch_sizes=(args.batch,))
rnd = random.Random(0)
sp = SamplingParams(temperature=0.6, ignore_eos=True, max_tokens=64)
for _ in range(args.batch):
    llm.add_request([rnd.randint(0, 10000) for _ in range(args.ctx)], sp)
if args.phase == "prefill":
    llm.step()                      # warm
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(args.steps):
        llm.step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    while llm.scheduler.waiting:
        llm.step()
    for _ in range(3):
        llm.step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(args.steps):
        llm.step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
llm.exit()
print("profile_step done")
