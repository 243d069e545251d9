"""Per-kernel timeline of decode steps (CUPTI via torch.profiler; no nsys in this image).

Builds the same engine as bench.py (optionally tensor-parallel under torchrun), prefills, warms up,
then records N decode steps on rank 0 and prints every kernel name with launches / step, average
and total microseconds and its share of the GPU-busy time, plus the step wall time measured WITHOUT
the profiler just before. Use it to see where a tensor-parallel step goes (GEMMs vs attention vs
all-reduce vs launch gaps) before and after switching the all-reduce form:

  python tools/step_timeline.py --model qwen3-0.6b
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      tools/step_timeline.py --model qwen3-32b --tp --mem-util 0.85
  B200_NVLS_ALLREDUCE=1 python -m torch.distributed.run ... tools/step_timeline.py --model qwen3-32b --tp

Numbers taken under the profiler are for shares only; the un-profiled step time is printed first.
"""
import argparse
import collections
import os
import random
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nano-vllm-ascend_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="qwen3-0.6b")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--ctx", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--tp", action="store_true")
    ap.add_argument("--shard-of", type=int, default=0,
                    help="single process: rank 0's shard of an N-way tensor-parallel model, collectives "
                         "skipped (Config.tp_shard_emulation) -- the per-rank compute of a TP=N step")
    ap.add_argument("--mem-util", type=float, default=0.7)
    ap.add_argument("--kv-cache-dtype", default="auto")
    ap.add_argument("--out", default=None, help="also write the table to this file (rank 0)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from nanovllm import LLM, SamplingParams
    llm = LLM(args.model, kvcache_block_size=16, max_num_seqs=args.batch, max_model_len=4096,
              max_num_batched_tokens=16384, tensor_parallel_size=world if args.tp else 1,
              distributed_launch="env", gpu_memory_utilization=args.mem_util,
              kv_cache_dtype=args.kv_cache_dtype, seed=0, tp_shard_emulation=args.shard_of,
              graph_batch_sizes=(args.batch,))
    rnd = random.Random(0)
    sp = SamplingParams(temperature=0.6, ignore_eos=True, max_tokens=args.steps + 64)
    for _ in range(args.batch):
        llm.add_request([rnd.randint(0, 10000) for _ in range(args.ctx)], sp)
    while llm.scheduler.waiting:
        llm.step()
    for _ in range(5):
        llm.step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(8):
        llm.step()
    torch.cuda.synchronize()
    plain_ms = (time.perf_counter() - t0) / 8 * 1e3

    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            llm.step()
        torch.cuda.synchronize()
    if rank == 0:
        agg = collections.OrderedDict()
        busy = 0.0
        try:
            for ev in prof.events():
                if ev.device_type != torch.autograd.DeviceType.CUDA:
                    continue
                us = ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
                name = ev.name[:110]
                n, tot = agg.get(name, (0, 0.0))
                agg[name] = (n + 1, tot + us)
                busy += us
        except Exception as e:            # profiler event API drift: fall back to torch's own table
            print(f"event walk failed ({type(e).__name__}: {e}); key_averages() instead")
            print(prof.key_averages().table(row_limit=40))
            agg, busy = collections.OrderedDict(), 1.0
        lines = [f"model={args.model} tp={world if args.tp else 1} shard_of={args.shard_of} batch={args.batch} ctx~{args.ctx} "
                 f"step (no profiler) = {plain_ms:.3f} ms; GPU-busy under profiler = "
                 f"{busy / args.steps / 1e3:.3f} ms/step over {args.steps} steps",
                 f"{'launches/step':>13} {'avg us':>9} {'us/step':>10} {'share':>7}  kernel"]
        for name, (n, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            lines.append(f"{n / args.steps:13.1f} {tot / n:9.2f} {tot / args.steps:10.1f} "
                         f"{100 * tot / busy:6.1f}%  {name}")
        text = "\n".join(lines)
        print(text)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as f:
                f.write(text + "\n")
    llm.exit()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
