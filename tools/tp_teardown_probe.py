"""Diagnostic (torchrun, >= 2 GPUs): does dist.destroy_process_group() return after CUDA graphs that
embed NCCL kernels were replayed? Times each teardown stage; a hang is reported by faulthandler.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/tp_teardown_probe.py
"""
import faulthandler
import gc
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nano-vllm-ascend_b200")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

faulthandler.dump_traceback_later(240, exit=True)
rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from nanovllm import LLM, SamplingParams  # noqa: E402

t = time.time()


def lap(what):
    global t
    torch.cuda.synchronize()
    now = time.time()
    print(f"[rank{rank}] {what}: {now - t:.2f} s", flush=True)
    t = now


llm = LLM(os.environ.get("PROBE_MODEL", "qwen3-0.6b"), kvcache_block_size=16, max_num_seqs=64,
          max_model_len=1024, max_num_batched_tokens=4096, tensor_parallel_size=world,
          distributed_launch="env", gpu_memory_utilization=0.3, seed=0)
lap("engine built")
outs = llm.generate([[1, 2, 3, 4, 5]] * 48, SamplingParams(temperature=0.6, ignore_eos=True, max_tokens=40),
                    use_tqdm=False)
lap("generated")
keep_graph_ref = llm.model_runner.graphs.get(64) if os.environ.get("PROBE_KEEP_GRAPH") else None
llm.exit()
lap("llm.exit()")
del llm
gc.collect()
torch.cuda.empty_cache()
lap("gc")
dist.barrier()
lap("barrier")
dist.destroy_process_group()
lap("destroy_process_group")
print(f"[rank{rank}] clean exit", flush=True)
