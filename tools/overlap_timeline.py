"""Per-kernel timelines (start, duration, stream) of ONE replayed decode step of the config-2 engine for a
few variants of the step (one batch / two micro-batches, attention grid size, cuBLAS or the
library's own GEMMs in the split step), plus device-timed ms/step of each variant and an eager check of the
split forward against the one-batch forward. tools/overlap_summary.py turns a trace into the per-kernel
table and the attention-busy fraction quoted in DESIGN 3.8. Kernel records come from the CUPTI activity stream through
torch.profiler (chrome trace per variant under gpurun_out/); timings are taken outside the profiler.
"""
import argparse
import json
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nano-vllm-ascend_b200"))
from nanovllm import LLM, SamplingParams, _C  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--ctx", type=int, default=1024)
ap.add_argument("--reps", type=int, default=40)
ap.add_argument("--outdir", default=os.path.join(ROOT, "gpurun_out"))
args = ap.parse_args()
os.makedirs(args.outdir, exist_ok=True)
sink = open(os.path.join(args.outdir, "overlap_timeline.jsonl"), "a")


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    sink.write(line + "\n")
    sink.flush()


B = args.batch
llm = LLM("qwen3-0.6b", kvcache_block_size=16, max_num_seqs=B, max_model_len=4096,
          max_num_batched_tokens=16384, seed=0, graph_batch_sizes=(B,))
runner = llm.model_runner
lib = _C.load()
rnd = random.Random(0)
sp = SamplingParams(temperature=0.6, ignore_eos=True, max_tokens=100)
for _ in range(B):
    llm.add_request([rnd.randint(0, 10000) for _ in range(args.ctx)], sp)
while llm.scheduler.waiting:
    llm.step()
for _ in range(3):
    llm.step()
torch.cuda.synchronize()


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


def recapture(split, ctas, own):
    """Re-capture bucket B with these step options; returns the graph to replay."""
    runner.config.decode_overlap_min_batch = 16 if split else 0
    runner.config.decode_overlap_min_kv_bytes = 0            # every eligible step splits
    runner.config.decode_overlap_own_gemm = bool(own)
    runner.config.decode_attn_ctas = ctas
    runner.graphs.clear()
    runner.split_graphs.clear()
    runner._capture_decode_graphs()
    torch.cuda.synchronize()
    return runner.decode_graph(B)


# eager: split forward with the library's own GEMMs vs the one-batch cuBLAS forward
with torch.inference_mode():
    runner._set_decode_context(B, B)
    want = runner.model(runner.d_ids[:B], runner.d_pos[:B]).float().clone()
    runner.config.decode_overlap_min_batch, runner.config.decode_overlap_min_kv_bytes = 16, 0
    for own in (False, True):
        runner.config.decode_overlap_own_gemm = own
        runner._set_decode_context(B, B)
        got = runner._forward_overlapped(B, runner._overlap_split(B)).float().clone()
        torch.cuda.synchronize()
        d = (got - want).abs()
        emit(kind="eager_check", own_gemm=own, max_abs=float(d.max()), rms=float(d.pow(2).mean().sqrt()),
             ref_rms=float(want.pow(2).mean().sqrt()))

# name, two micro-batches?, attention CTAs (0 = one per SM), own GEMMs in the split step, profile?
variants = [
    ("one_batch_c148", 0, 0, 0, True),
    ("one_batch_c104", 0, 104, 0, False),
    ("split_cublas_c104", 1, 104, 0, True),
    ("split_own_c148", 1, 0, 1, False),
    ("split_own_c120", 1, 120, 1, False),
    ("split_own_c104", 1, 104, 1, True),
    ("split_own_c88", 1, 88, 1, False),
    ("one_batch_c148_again", 0, 0, 0, False),
]
results = {}
for name, split, ctas, own, prof in variants:
    try:
        g, launches = recapture(split, ctas, own)
        ms = timed(g.replay, args.reps)
        results[name] = (ms, (split, ctas, own))
        emit(kind="step", variant=name, ms_per_step=round(ms, 4), tok_s=round(B / ms * 1e3, 1), launches=launches)
        if prof:
            with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as p:
                g.replay()
                torch.cuda.synchronize()
            p.export_chrome_trace(os.path.join(args.outdir, f"trace_{name}.json"))
    except Exception as e:
        emit(kind="step", variant=name, error=repr(e)[:300])
        torch.cuda.synchronize()

# the engine's own step() loop (wall clock, host work included) for the best variant and the r02 default
import time  # noqa: E402
best = min(results, key=lambda k: results[k][0]) if results else None
for name in dict.fromkeys(["one_batch_c148", best]):
    if name is None or name not in results:
        continue
    recapture(*results[name][1])
    for _ in range(3):
        llm.step()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(12):
        llm.step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 12
    emit(kind="engine_step", variant=name, ms_per_step=round(dt * 1e3, 4), tok_s=round(B / dt, 1),
         ctx=int(runner.h_ctx[:B].max()))
llm.exit()
print("overlap_timeline done")
