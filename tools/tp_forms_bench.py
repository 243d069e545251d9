"""A/B of the all-reduce forms on the tensor-parallel decode step (BASELINE configs[3]) in ONE torchrun
launch (8-GPU box time is charged 8x: no repeated interpreter / NCCL start-up). One JSON line per form.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 \
      tools/tp_forms_bench.py --forms auto,nvls_fused,nccl,nvls [--model qwen3-32b] [--steps 24]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nano-vllm-ascend_b200")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--forms", default="auto,nvls_fused,nccl,nvls")
    ap.add_argument("--model", default="qwen3-32b")
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--budget-s", type=float, default=400.0, help="no new form is started after this")
    a = ap.parse_args()
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    if os.environ.get("B200_BENCH_DEBUG"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["B200_BENCH_DEBUG"]), exit=True)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    args = argparse.Namespace(batch=256, ctx=1024, block_size=16, tp_model=a.model)
    kw = dict(kvcache_block_size=16, max_num_seqs=256, max_model_len=4096, max_num_batched_tokens=16384,
              distributed_launch="env", gpu_memory_utilization=0.85, graph_batch_sizes=(256,), seed=0)
    peak, src = bench.peaks()
    t0 = time.time()
    for form in a.forms.split(","):
        go = torch.tensor([1 if time.time() - t0 < a.budget_s else 0], device="cuda")
        dist.broadcast(go, 0)                       # every rank takes rank 0's decision
        if not int(go.item()):
            break
        os.environ["B200_ALLREDUCE"] = form         # read by PeerAllReduce at engine construction
        os.environ["B200_SHARDED_SAMPLING"] = "0" if form == "nccl" else "1"   # nccl = the reference's structure
        out = {"requested_form": form, "model": a.model, "tp": world}
        out.update(bench.tp_engine_measure(args, rank, world, a.steps, a.warmup, peak, src, kw))
        if rank == 0:
            print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
