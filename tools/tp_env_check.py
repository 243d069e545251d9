"""torchrun smoke for distributed_launch="env" tensor parallelism (debug tool)."""
import faulthandler
import os
import random
import sys
import time

faulthandler.dump_traceback_later(int(os.environ.get("B200_DUMP_AFTER", 100)), exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nano-vllm-ascend_b200"))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from nanovllm import LLM, SamplingParams

def log(*a):
    print(f"[rank{rank} {time.strftime('%X')}]", *a, file=sys.stderr, flush=True)

model = sys.argv[1] if len(sys.argv) > 1 else "tiny-qwen3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 40
kw = dict(num_kvcache_blocks=128, max_model_len=256, max_num_batched_tokens=1024) if model == "tiny-qwen3" else \
    dict(max_model_len=4096, max_num_batched_tokens=16384)
log("building")
llm = LLM(model, kvcache_block_size=16, max_num_seqs=B, tensor_parallel_size=world,
          distributed_launch="env", seed=0, **kw)
log("built; blocks", llm.config.num_kvcache_blocks)
rnd = random.Random(0)
hi = 511 if model == "tiny-qwen3" else 10000
prompts = [[rnd.randint(0, hi) for _ in range(ctx)] for _ in range(B)]
sp = SamplingParams(temperature=1.0, max_tokens=12, ignore_eos=True, greedy=True)
for p in prompts:
    llm.add_request(p, sp)
n = 0
while not llm.is_finished():
    llm.step()
    n += 1
    if n < 4 or n % 8 == 0:
        log("step", n)
log("generate done")
llm.exit()
dist.barrier()
dist.destroy_process_group()
log("clean exit")
