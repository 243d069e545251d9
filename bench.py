#!/usr/bin/env python
"""Contract benchmark: decode tokens/s (+ p50 TTFT) of the paged-KV attention path.

Workload (BASELINE.json configs[1]): Qwen3-0.6B (random-init, bf16), 256 sequences x 1024-token
prompts, kvcache_block_size 16, CUDA-graph decode, one engine per GPU.  A "step" is one decode
step of the whole batch (one pass of the hot path: 28 x [K8 epilogue, paged decode attention,
fused norms/activation] + cuBLAS GEMMs + sampling).

  value      device-timed decode throughput: K graph replays with the step's inputs resident in HBM
  e2e        the same K steps through the public API (LLM.step): per step one H2D of the packed
             step metadata from pinned memory and one D2H of the sampled ids, scheduler included
  roofline   the dominant kernel (paged_decode_kernel) timed alone over all 28 layers' caches
             (28 GiB streamed per pass, far beyond L2), algorithmic KV bytes / event time
  cpu_baseline  the reference's CPU path (oracle port of models/qwen3.py + torch-native
             attention) on the host cores, bounded sample of the same decode step

N > 1 (torchrun): independent data-parallel replicas, no data-path collective ("weak" scaling);
`--tp` instead shards one engine with tensor parallelism (row-parallel all-reduce over NCCL).
`--impl reference` times only the CPU reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "nano-vllm-ascend_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "decode_tokens_per_s"
UNIT = "tokens/s"


def moe_expert_roofline(runner, B, peak_gbs):
    """Sparse-MoE models: the expert GEMMs (csrc/moe.cu) are the other HBM stream of a decode step.
    Timed alone with CUDA events over every layer's own expert weights (>= 1 GB per layer, cold in L2);
    algorithmic bytes per layer = experts hit x 3 x hidden x moe_intermediate x 2 B."""
    from nanovllm import ops
    from nanovllm.layers.moe import SparseMoeBlock
    blocks = [m for m in runner.model.modules() if isinstance(m, SparseMoeBlock)]
    blk = blocks[0]
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, blk.hidden_size, device="cuda", generator=g).to(runner.dtype)
    scores = torch.rand(B, blk.num_experts, device="cuda", generator=g)
    ids = scores.topk(blk.top_k, dim=1).indices.to(torch.int32).contiguous()
    w = torch.full((B, blk.top_k), 1.0 / blk.top_k, device="cuda").to(runner.dtype)
    plan = ops.moe_align(ids, blk.expert_begin, blk.num_local)
    hit = int((plan.offsets[1:] > plan.offsets[:-1]).sum())

    def one_pass():
        for b in blocks:
            ops.moe_experts(x, b.w13, b.w2, ids, w, plan)
    for _ in range(2):
        one_pass()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 4
    e0.record()
    for _ in range(reps):
        one_pass()
    e1.record()
    torch.cuda.synchronize()
    per_layer_s = e0.elapsed_time(e1) / 1e3 / (reps * len(blocks))
    nbytes = hit * 3 * blk.hidden_size * blk.inter * x.element_size()
    achieved = nbytes / per_layer_s / 1e9
    return {"bound": "hbm", "kernel": "moe_gemm_kernel<gate_up> + <down> + moe_combine_kernel",
            "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
            "algorithmic_bytes_per_launch": nbytes, "us_per_layer": per_layer_s * 1e6,
            "experts_hit": hit, "local_experts": blk.num_local, "layers": len(blocks)}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="qwen3-0.6b")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--ctx", type=int, default=1024)
    ap.add_argument("--block-size", type=int, default=16)
    ap.add_argument("--tp", action="store_true", help="tensor-parallel over --gpus instead of replicas")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mem-util", type=float, default=0.7, help="gpu_memory_utilization of the engine")
    ap.add_argument("--cpu-sample-seqs", type=int, default=32)
    ap.add_argument("--kv-cache-dtype", default="auto", choices=["auto", "fp8"],
                    help="fp8: e4m3 KV cache (an extra line; the contract line is the default bf16 cache)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed regions."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # under load = samples within 85% of the hottest power draw is unknowable here; use median
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU reference arm
def cpu_reference_step(n_seqs, ctx, block_size, steps, warmup, threads=None):
    """The reference's CPU path (oracle port: eager layers + torch-native attention, bf16) on a
    bounded sample: one full-model decode step for `n_seqs` of the workload's sequences at the
    workload's context length. Returns (tokens/s, seconds per step)."""
    from oracle import model_ref as M
    if threads:
        torch.set_num_threads(threads)
    cfg = M.ModelCfg.qwen3_0_6b()
    w = M.seeded_weights(cfg, seed=0)
    nblk_per = (ctx + 1 + block_size - 1) // block_size
    nb = n_seqs * nblk_per + 1
    g = torch.Generator().manual_seed(0)
    kv = (0.5 * torch.randn(2, cfg.num_layers, nb, block_size, cfg.num_kv_heads, cfg.head_dim,
                            generator=g)).to(cfg.dtype)
    tables = [list(range(i * nblk_per, (i + 1) * nblk_per)) for i in range(n_seqs)]
    lens = [ctx] * n_seqs
    ids = torch.randint(0, 10000, (n_seqs,), generator=g)
    times = []
    with torch.inference_mode():
        for i in range(warmup + steps):
            pos, meta = M.decode_meta(lens, tables, block_size)
            t0 = time.perf_counter()
            logits = M.forward(w, cfg, ids, pos, meta, kv, block_size, attn_mode="reference")
            ids = logits.float().argmax(-1)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    sec = sum(times) / len(times)
    return n_seqs / sec, sec


def run_reference_arm(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # bounded sample: as many of the workload's sequences as keep the whole run near a minute
    # (about 0.07 s per sequence-step on 8 host cores)
    n_seqs = max(1, min(32, args.batch, int(600 / (steps + warmup))))
    args.cpu_sample_seqs = n_seqs
    tps, sec = cpu_reference_step(n_seqs, args.ctx, args.block_size, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": tps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.cpu_sample_seqs} of {args.batch} sequences, full 28-layer decode "
                                   f"step at ctx={args.ctx} (oracle port of the reference CPU path; the "
                                   f"Python reference itself cannot travel to the GPU box)"},
        "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, world):
    par = "replicas" if not args.tp else f"tp{world}"
    return {"workload": f"{args.model} decode, batch {args.batch}/GPU, ctx {args.ctx}, "
                        f"kvcache_block_size {args.block_size}, CUDA-graph decode (BASELINE configs[1])",
            "global_batch": args.batch * (1 if args.tp else world), "seq_len": args.ctx,
            "parallelism": f"{par}x{world}" if world > 1 else "single",
            "l2": "inputs exceed L2 (>= 28 GiB of KV streamed per step)",
            **({} if args.kv_cache_dtype == "auto" else {"kv_cache_dtype": "fp8_e4m3"})}


# ------------------------------------------------------------------ the GPU arm
def dbg(rank, *a):
    if os.environ.get("B200_BENCH_DEBUG"):
        print(f"[bench rank{rank} {time.strftime('%X')}]", *a, file=sys.stderr, flush=True)


def main():
    args = parse_args()
    if os.environ.get("B200_BENCH_DEBUG"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["B200_BENCH_DEBUG"]), exit=True)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the b200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from nanovllm import LLM, SamplingParams, ops
    K, W, B, ctx = args.steps, max(3, args.warmup), args.batch, args.ctx
    max_len = ctx + W + K + 16
    llm = LLM(args.model, kvcache_block_size=args.block_size, max_num_seqs=B,
              max_model_len=max(4096, max_len), max_num_batched_tokens=16384,
              tensor_parallel_size=world if args.tp else 1, distributed_launch="env",
              gpu_memory_utilization=args.mem_util, kv_cache_dtype=args.kv_cache_dtype, seed=0)
    runner = llm.model_runner
    spec = runner.spec
    rnd = random.Random(rank if not args.tp else 0)
    prompts = [[rnd.randint(0, 10000) for _ in range(ctx)] for _ in range(B)]
    sp = SamplingParams(temperature=0.6, ignore_eos=True, max_tokens=W + K + 8)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clocks = ClockSampler(local)
    dbg(rank, 'engine built')
    # ---- phase 1: submit everything at t=0, prefill (TTFT)
    barrier()
    clocks.start()
    t_submit = time.perf_counter()
    seqs = [llm.add_request(p, sp) for p in prompts]
    prefill_steps = 0
    while llm.scheduler.waiting:
        llm.step()
        prefill_steps += 1
    ttfts = sorted((s.first_token_time - t_submit) * 1e3 for s in seqs)
    ttft_p50 = ttfts[len(ttfts) // 2]
    prefill_s = ttfts[-1] / 1e3

    dbg(rank, 'prefill done')
    # ---- phase 2: e2e decode through the public API (host buffers, H2D + D2H inside)
    for _ in range(W):
        llm.step()
    barrier()
    h2d0, d2h0 = runner.h2d_bytes, runner.d2h_bytes
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(K):
        out, ntok = llm.step()
        assert ntok == -B, f"decode batch shrank to {-ntok}"
    e1.record()
    torch.cuda.synchronize()
    e2e_wall = time.perf_counter() - t0
    e2e_s = max(e0.elapsed_time(e1) / 1e3, e2e_wall)
    e2e_s = max_over_ranks(e2e_s)
    h2d_per = (runner.h2d_bytes - h2d0) / K
    d2h_per = (runner.d2h_bytes - d2h0) / K
    ctx_now = ctx + W + K            # context length the resident buffers now describe

    dbg(rank, 'e2e done')
    # ---- phase 3: device-only: replay the captured step with its inputs resident in HBM
    graph = runner.graphs[B]
    for _ in range(W):
        graph.replay()
    barrier()
    e0.record()
    for _ in range(K):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    dev_s = max_over_ranks(e0.elapsed_time(e1) / 1e3)
    launches = runner.launches_per_graph.get(B, 0) * K

    dbg(rank, 'replay done')
    # ---- phase 4: the dominant kernel alone, all layers' caches (cold in L2), CUDA events
    tpw = world if args.tp else 1
    Hq, Hkv, D = spec.num_heads // tpw, max(1, spec.num_kv_heads // tpw), spec.head_dim
    q = torch.randn(B, Hq, D, device="cuda", dtype=runner.dtype)
    o = torch.empty(B, Hq * D, device="cuda", dtype=runner.dtype)
    ws = ops.decode_workspace(B, Hq, Hkv, D, q.device)
    L = spec.num_layers

    attn_fn = ops.paged_decode_attn if args.kv_cache_dtype == "auto" else ops.paged_decode_attn_fp8

    def attn_pass():
        for layer in range(L):
            attn_fn(q, runner.kv_cache[0, layer], runner.kv_cache[1, layer],
                    runner.d_tables[:B], runner.d_ctx[:B], D ** -0.5, Hkv, out=o, workspace=ws)
    for _ in range(3):
        attn_pass()
    barrier()
    reps = max(4, min(K, 32))
    e0.record()
    for _ in range(reps):
        attn_pass()
    e1.record()
    torch.cuda.synchronize()
    attn_launch_s = e0.elapsed_time(e1) / 1e3 / (reps * L)
    dbg(rank, 'attn loop done')
    clk = clocks.stop()
    ctx_sum = int(runner.d_ctx[:B].sum().item())
    attn_bytes = ctx_sum * 2 * Hkv * D * runner.kv_cache.dtype.itemsize  # algorithmic bytes / launch
    peak, peak_src = peaks()
    achieved = attn_bytes / attn_launch_s / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "decode_attn_traffic.json")
    if os.path.exists(tpath) and args.kv_cache_dtype == "auto":
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")

    value = world * B * K / dev_s if not args.tp else B * K / dev_s
    e2e_val = world * B * K / e2e_s if not args.tp else B * K / e2e_s
    # KV-read roofline of the whole step at the measured peak (what `value` is bounded by)
    step_kv_bytes = attn_bytes * L
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_s / K * 1e3, "higher_is_better": True, "scaling": "weak" if not args.tp else "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic (random-init weights, random token ids)",
        "config": dict(workload_config(args, world), ctx_during_timing=[ctx + W, ctx_now]),
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_per, "d2h_bytes_per_step": d2h_per,
                "ms_per_step": e2e_s / K * 1e3},
        "ttft_p50_ms": ttft_p50, "ttft_max_ms": ttfts[-1], "prefill_steps": prefill_steps,
        "prefill_tokens_per_s": world * B * ctx / prefill_s if not args.tp else B * ctx / prefill_s,
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "paged_decode_kernel" if args.kv_cache_dtype == "auto"
                     else "paged_decode_fp8_kernel", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": attn_bytes,
                     "us_per_launch": attn_launch_s * 1e6,
                     "step_kv_roofline_tokens_per_s": B / (step_kv_bytes / (peak * 1e9)),
                     "step_frac_of_kv_roofline": (B * K / dev_s) / (B / (step_kv_bytes / (peak * 1e9)))},
        "clocks": clk,
    }
    if getattr(spec, "num_experts", 0):
        line["moe_roofline"] = moe_expert_roofline(runner, B, peak)
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.model == "qwen3-0.6b":
        tps, sec = cpu_reference_step(args.cpu_sample_seqs, ctx, args.block_size, 3, 1)
        line["cpu_baseline"] = {"value": tps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{args.cpu_sample_seqs} of {B} sequences, one full 28-layer decode "
                                          f"step at ctx={ctx}, 3 timed steps of {sec:.1f} s each"}
    dbg(rank, 'before exit')
    llm.exit()
    dbg(rank, 'after exit')
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        # destroy_process_group() can block for minutes once CUDA graphs that embedded NCCL kernels
        # have been replayed; everything is measured and printed, so leave without the teardown
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
