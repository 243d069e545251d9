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
             (28 GiB streamed per pass, far beyond L2), launched the way the step launches it (the
             runner's grid; two half-batch launches per layer when the step runs as two micro-batches,
             with the one-launch-per-layer figure beside it), algorithmic KV bytes / event time
  cpu_baseline  the reference's CPU path (oracle port of models/qwen3.py + torch-native
             attention) on the host cores, bounded sample of the same decode step

N > 1 (torchrun): independent data-parallel replicas, no data-path collective ("weak" scaling) --
the headline line -- plus a `tp` object: Qwen3-32B (BASELINE configs[3]) decode at the same batch /
context tensor-parallel over all N GPUs (fused NVLink all-reduce + residual add + RMSNorm), with
the all-reduce form and its latency, and the same workload at TP=1 (rank 0) for the speed-up.
`--tp` instead makes the tensor-parallel engine the whole line ("strong" scaling).
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
    ap.add_argument("--tp-model", default="qwen3-32b",
                    help="model of the `tp` object a multi-GPU replica run adds (BASELINE configs[3])")
    ap.add_argument("--no-tp-block", action="store_true", help="multi-GPU replica run without the `tp` object")
    ap.add_argument("--tp-timeout", type=float, default=420.0, help="seconds before the `tp` object is abandoned")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mem-util", type=float, default=0.7, help="gpu_memory_utilization of the engine")
    ap.add_argument("--cpu-sample-seqs", type=int, default=32)
    ap.add_argument("--kv-cache-dtype", default="auto", choices=["auto", "fp8"],
                    help="fp8: e4m3 KV cache (an extra line; the contract line is the default bf16 cache)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except (OSError, ValueError, KeyError, TypeError):
        return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["bf16_tflops"]), "measured burst cuBLAS bf16 (MEASURED_PEAKS.json)"
    except (OSError, ValueError, KeyError, TypeError):
        return 1590.0, "fallback (B200_PROFILING.md)"


def prefill_roofline():
    """The tensor-bound kernel of the path (prefill_umma_kernel) timed alone with CUDA events on the
    shapes BASELINE names: configs[1]'s prefill batch (16 x 1024 tokens, Qwen3-0.6B heads) and configs[2]
    (Llama-3.2-1B, 8192-token prompt: whole, and the prefix-cache-hit form -- 256 new tokens against
    7936 cached ones read through block tables). Flops = 4 * H * D * sum over rows of visible keys
    (causal); inputs rotate over copies totalling > L2."""
    from nanovllm import ops
    peak, src = tensor_peak()
    g = torch.Generator(device="cuda").manual_seed(0)
    bf = torch.bfloat16

    def timed(fns, iters=24):
        for f in fns:
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fns[i % len(fns)]()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 1e3 / iters

    def packed(nseq, L, H, G, D, copies):
        T = nseq * L
        cu = torch.arange(0, T + 1, L, dtype=torch.int32, device="cuda")
        out = torch.empty(T, H * D, dtype=bf, device="cuda")
        fns = []
        for _ in range(copies):
            qkv = torch.randn(T, (H + 2 * G) * D, device="cuda", generator=g).to(bf)
            q = qkv[:, :H * D].unflatten(1, (H, D))
            k = qkv[:, H * D:(H + G) * D].unflatten(1, (G, D))
            v = qkv[:, (H + G) * D:].unflatten(1, (G, D))
            fns.append(lambda q=q, k=k, v=v: ops.prefill_varlen_attn(q, k, v, cu, cu, L, L, D ** -0.5, G, out=out))
        sec = timed(fns)
        flops = 2 * L * L * H * D * nseq
        return {"shape": f"{nseq} x {L} tokens, H={H} Hkv={G} D={D}, packed, causal", "us": sec * 1e6,
                "tflops": flops / sec / 1e12, "frac": flops / sec / 1e12 / peak}

    def prefix_hit(qlen, klen, H, G, D, bs, copies):
        nblk = (klen + bs - 1) // bs
        cu_q = torch.tensor([0, qlen], dtype=torch.int32, device="cuda")
        cu_k = torch.tensor([0, klen], dtype=torch.int32, device="cuda")
        out = torch.empty(qlen, H * D, dtype=bf, device="cuda")
        q = torch.randn(qlen, H, D, device="cuda", generator=g).to(bf)
        fns = []
        for _ in range(copies):
            kc = torch.randn(nblk + 1, bs, G * D, device="cuda", generator=g).to(bf)
            vc = torch.randn(nblk + 1, bs, G * D, device="cuda", generator=g).to(bf)
            bt = torch.randperm(nblk, device="cuda", generator=g).to(torch.int32).view(1, -1).contiguous()
            fns.append(lambda kc=kc, vc=vc, bt=bt: ops.prefill_varlen_attn(
                q, kc, vc, cu_q, cu_k, qlen, klen, D ** -0.5, G, block_tables=bt, out=out))
        sec = timed(fns)
        flops = 4 * H * D * (qlen * (klen - qlen) + qlen * qlen / 2)
        return {"shape": f"q {qlen} / k {klen} tokens, H={H} Hkv={G} D={D}, paged block {bs} (prefix-cache hit)",
                "us": sec * 1e6, "tflops": flops / sec / 1e12, "frac": flops / sec / 1e12 / peak}

    main = packed(16, 1024, 16, 8, 128, 2)
    shapes = [main, packed(4, 4096, 16, 8, 128, 2), packed(1, 8192, 32, 8, 64, 4),
              prefix_hit(256, 8192, 32, 8, 64, 16, 16)]
    return {"bound": "tensor", "kernel": "prefill_umma_kernel", "achieved": main["tflops"], "peak": peak,
            "unit": "TFLOP/s", "frac": main["frac"], "peak_source": src, "traffic": None,
            "shapes": shapes}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed regions. The process
    is started early (its NVML start-up takes driver locks for a few hundred ms: started right before
    the timed region it randomly stalled the first prefill step by 100-400 ms and with it every TTFT);
    only rows that arrive after begin() are counted."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.t_begin = [], None, index, 0.0

    def begin(self):
        self.t_begin = time.perf_counter()

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
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if t < self.t_begin:
                continue
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


def reference_cpu_step(n_seqs, ctx, block_size, steps, warmup, in_process):
    """The CPU arm. With the reference archive present (oracle/_ref/, packed by oracle/stage_ref.py at
    build time) the reference's OWN modules run: models/qwen3.py + layers/attention_torch_native.py on
    all host cores (kind "reference"); otherwise the oracle port of the same graph (kind "port").
    The reference registers a stub package named `nanovllm`, which is also the product's package name:
    the GPU arm therefore runs it in a child process."""
    from oracle import ref_runtime
    cores = os.cpu_count() or 1
    if ref_runtime.reference_dir() is None:
        tps, sec = cpu_reference_step(n_seqs, ctx, block_size, steps, warmup, cores)
        return tps, sec, "port", torch.get_num_threads()
    if in_process:
        tps, sec = ref_runtime.reference_decode_step(n_seqs, ctx, block_size, steps, warmup, cores)
        return tps, sec, "reference", torch.get_num_threads()
    r = subprocess.run([sys.executable, "-m", "oracle.ref_runtime", "--seqs", str(n_seqs), "--ctx", str(ctx),
                        "--block-size", str(block_size), "--steps", str(steps), "--warmup", str(warmup)],
                       cwd=ROOT, capture_output=True, text=True, timeout=1200)
    try:
        if r.returncode != 0:
            raise RuntimeError(r.stderr[-2000:])
        d = json.loads(r.stdout.strip().splitlines()[-1])
        return d["tokens_per_s"], d["s_per_step"], "reference", d["threads"]
    except (RuntimeError, ValueError, KeyError, IndexError) as e:
        # the reference's modules could not be run here: say so and time the oracle port instead
        print(f"bench.py: reference CPU arm failed ({type(e).__name__}: {str(e)[-300:]}); "
              f"timing the oracle port", file=sys.stderr)
        tps, sec = cpu_reference_step(n_seqs, ctx, block_size, steps, warmup, cores)
        return tps, sec, "port", torch.get_num_threads()


def cpu_sample_text(kind, n_seqs, batch, ctx, steps, sec):
    what = ("the reference's own models/qwen3.py + attention_torch_native.py (oracle/_ref archive)"
            if kind == "reference" else
            "oracle port of the reference CPU path (no reference archive on this box)")
    return (f"batch {n_seqs} (of the workload's {batch} sequences), one full 28-layer decode step at "
            f"ctx={ctx}, {steps} timed steps of {sec:.2f} s each; {what}")


def run_reference_arm(args, rank):
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # bounded sample: as many of the workload's sequences as keep the whole run near a minute
    # (about 0.07 s per sequence-step on 8 host cores)
    n_seqs = max(1, min(32, args.batch, int(600 / (steps + warmup))))
    args.cpu_sample_seqs = n_seqs
    tps, sec, kind, cores = reference_cpu_step(n_seqs, args.ctx, args.block_size, steps, warmup, True)
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": dict(workload_config(args, 1), cpu_sample_batch=n_seqs),
        "cpu_baseline": {"value": tps, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": cpu_sample_text(kind, n_seqs, args.batch, args.ctx, steps, sec)},
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


class Group:
    """The ranks a measurement spans: every torchrun rank (replicas, TP=N) or this rank alone (the
    TP=1 comparison run rank 0 makes while the others wait)."""

    def __init__(self, world):
        self.world = world

    def barrier(self):
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max(self, x):
        if self.world == 1:
            return x
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


def decode_workload(llm, group, B, ctx, K, W, prompt_seed, rank):
    """The workload of BASELINE configs[1] on one engine: submit B prompts of `ctx` tokens at t = 0,
    prefill (TTFT), W warm-up decode steps, K decode steps through LLM.step() (e2e: pinned H2D of the
    step metadata + D2H of the sampled ids inside the timed region), then K replays of the captured
    step with its inputs resident (device). Times are CUDA events, max over the group's ranks."""
    from nanovllm import SamplingParams
    runner = llm.model_runner
    rnd = random.Random(prompt_seed)
    prompts = [[rnd.randint(0, 10000) for _ in range(ctx)] for _ in range(B)]
    sp = SamplingParams(temperature=0.6, ignore_eos=True, max_tokens=W + K + 8)
    group.barrier()
    t_submit = time.perf_counter()
    seqs = [llm.add_request(p, sp) for p in prompts]
    prefill_steps = 0
    step_ms = []
    while llm.scheduler.waiting:
        t_step = time.perf_counter()
        llm.step()
        step_ms.append((time.perf_counter() - t_step) * 1e3)
        prefill_steps += 1
    submit_ms = (t_step - t_submit) * 1e3 - sum(step_ms[:-1]) if step_ms else 0.0   # add_request x B
    ttfts = sorted((s.first_token_time - t_submit) * 1e3 for s in seqs)
    dbg(rank, 'prefill done')
    for _ in range(W):
        llm.step()
    group.barrier()
    h2d0, d2h0 = runner.h2d_bytes, runner.d2h_bytes
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(K):
        out, ntok = llm.step()
        assert ntok == -B, f"decode batch shrank to {-ntok}"
    e1.record()
    torch.cuda.synchronize()
    e2e_s = group.max(max(e0.elapsed_time(e1) / 1e3, time.perf_counter() - t0))
    dbg(rank, 'e2e done')
    graph, launches = runner.decode_graph(B)       # the graph run() replays at this batch / context
    for _ in range(W):
        graph.replay()
    group.barrier()
    e0.record()
    for _ in range(K):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    dev_s = group.max(e0.elapsed_time(e1) / 1e3)
    del graph
    dbg(rank, 'replay done')
    return {"e2e_s": e2e_s, "dev_s": dev_s, "ttfts": ttfts, "prefill_steps": prefill_steps,
            "prefill_step_ms": [round(t, 2) for t in step_ms], "submit_ms": round(submit_ms, 2),
            "h2d_per": (runner.h2d_bytes - h2d0) / K, "d2h_per": (runner.d2h_bytes - d2h0) / K,
            "launches": launches * K, "ctx_now": ctx + W + K,
            "two_micro_batches": runner.use_split(B, int(runner.h_ctx[:B].sum()))}


def attention_roofline(runner, group, B, K, tpw, fp8, peak, peak_src):
    """The dominant kernel alone: one launch per layer over that layer's own cache (all layers = tens
    of GB per pass, cold in L2), CUDA events on the launching stream."""
    from nanovllm import ops
    spec = runner.spec
    Hq, Hkv, D = spec.num_heads // tpw, max(1, spec.num_kv_heads // tpw), spec.head_dim
    q = torch.randn(B, Hq, D, device="cuda", dtype=runner.dtype)
    o = torch.empty(B, Hq * D, device="cuda", dtype=runner.dtype)
    ws = ops.decode_workspace(B, Hq, Hkv, D, q.device)
    L = spec.num_layers
    attn_fn = ops.paged_decode_attn_fp8 if fp8 else ops.paged_decode_attn

    # launched the way the step launches it: the runner's grid size, and -- when the step runs as two
    # micro-batches at this batch / context -- two launches of half the rows each
    ctas = runner._attn_ctas(B)
    split = runner._overlap_split(B) if runner.use_split(B, int(runner.h_ctx[:B].sum())) else 0
    parts = [(0, split), (split, B)] if split else [(0, B)]

    def attn_pass(pieces):
        for layer in range(L):
            for lo, hi in pieces:
                attn_fn(q[lo:hi], runner.kv_cache[0, layer], runner.kv_cache[1, layer],
                        runner.d_tables[lo:hi], runner.d_ctx[lo:hi], D ** -0.5, Hkv, out=o[lo:hi], workspace=ws,
                        num_ctas=ctas)

    def seconds_per_pass(pieces, reps):
        for _ in range(3):
            attn_pass(pieces)
        group.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            attn_pass(pieces)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 1e3 / reps
    reps = max(4, min(K, 32))
    launch_s = seconds_per_pass(parts, reps) / (L * len(parts))
    ctx_sum = int(runner.d_ctx[:B].sum().item())
    whole = None
    if split:      # for comparison: the same rows as ONE launch per layer (what the one-batch graph does)
        whole_s = seconds_per_pass([(0, B)], max(4, reps // 2)) / L
        whole_bytes = ctx_sum * 2 * Hkv * D * runner.kv_cache.dtype.itemsize
        whole = {"rows_per_launch": B, "us_per_launch": whole_s * 1e6, "achieved": whole_bytes / whole_s / 1e9,
                 "frac": whole_bytes / whole_s / 1e9 / peak}
    nbytes = ctx_sum * 2 * Hkv * D * runner.kv_cache.dtype.itemsize / len(parts)   # algorithmic bytes / launch
    achieved = nbytes / launch_s / 1e9
    traffic = traffic_ctx = None
    tpath = os.path.join(ROOT, "profiles", "decode_attn_traffic.json")
    if os.path.exists(tpath) and not fp8 and tpw == 1:
        with open(tpath) as f:
            t = json.load(f)
        # the ncu capture was taken at its own context length: scale its bytes/token to this run's
        per_token = t["dram_bytes_per_launch"] / t["context_tokens_per_launch"]
        traffic, traffic_ctx = per_token * ctx_sum / len(parts), t["context_tokens_per_launch"] // B
    return {"bound": "hbm", "kernel": "paged_decode_fp8_kernel" if fp8 else "paged_decode_kernel",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": None if traffic is None else
            f"ncu dram__bytes_read+write per cached token at ctx {traffic_ctx} (profiles/"
            f"decode_attn_traffic.json) x this run's {ctx_sum} cached tokens per launch",
            "peak_source": peak_src, "algorithmic_bytes_per_launch": nbytes,
            "us_per_launch": launch_s * 1e6, "layers": L, "launches_per_layer": len(parts),
            "rows_per_launch": B // len(parts), "grid_ctas": ctas or None, "one_launch_per_layer": whole}


def release(llm):
    """Tear an engine down and give its memory back (the caller must hold no other reference to the
    runner): graphs, KV cache, weights, peer-memory regions."""
    import gc
    llm.exit()
    del llm
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


def allreduce_microbench(runner, rows, reps=200):
    """The layer's collective alone: the fused all-reduce + residual add + RMSNorm kernel over NVLink
    (or NCCL all-reduce + the add-RMSNorm kernel when the fused path does not take the shape)."""
    import torch.distributed as dist
    from nanovllm import ops
    from nanovllm.comm import get_comm
    hidden = runner.spec.hidden_size
    comm = get_comm()
    res = torch.zeros(rows, hidden, dtype=runner.dtype, device="cuda")
    w = torch.ones(hidden, dtype=runner.dtype, device="cuda")
    out = torch.empty_like(res)
    form = comm.mode(rows, hidden, res.element_size()) if comm is not None else None

    def one():
        if form is not None:
            comm.partial_buffer(rows, hidden, runner.dtype)
            comm.allreduce_add_rmsnorm(out, res, w, 1e-6)
        else:
            dist.all_reduce(out)
            ops.add_rmsnorm_(out, res, w, 1e-6)
    for _ in range(10):
        one()
    torch.cuda.synchronize()
    dist.barrier()
    # inside a CUDA graph, like the decode step: device time per call, not the host's launch rate
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            one()
    graph.replay()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    del graph
    return form or "nccl", e0.elapsed_time(e1) * 1e3 / reps


def tp_engine_measure(args, rank, world, K, W, peak, peak_src, kw):
    """One tensor-parallel engine over all `world` ranks: decode workload, all-reduce microbench,
    attention roofline. Returns the `tp` fields that do not need the TP=1 comparison."""
    from nanovllm import LLM
    B, ctx = args.batch, args.ctx
    llm = LLM(args.tp_model, tensor_parallel_size=world, **kw)
    runner = llm.model_runner
    r = decode_workload(llm, Group(world), B, ctx, K, W, 0, rank)
    form, ar_us = allreduce_microbench(runner, B)
    roof = attention_roofline(runner, Group(world), B, K, world, False, peak, peak_src)
    L, spec = runner.spec.num_layers, runner.spec
    weight_bytes = sum(p.numel() * p.element_size() for p in runner.model.parameters())
    step_bytes = weight_bytes + roof["algorithmic_bytes_per_launch"] * L
    out = {
        "tokens_per_s": B * K / r["dev_s"], "ms_per_step": r["dev_s"] / K * 1e3,
        "e2e_tokens_per_s": B * K / r["e2e_s"], "e2e_ms_per_step": r["e2e_s"] / K * 1e3,
        "ttft_p50_ms": r["ttfts"][len(r["ttfts"]) // 2], "gpu_launches_per_step": r["launches"] // K,
        "allreduce_form": form, "allreduce_us": ar_us, "allreduces_per_step": 2 * L,
        "allreduce_message_bytes": B * spec.hidden_size * 2,
        "sampling": "vocab-sharded keys + 8 B/row MAX all-reduce" if runner.sharded_sampling else
                    "logits gathered on rank 0",
        "attention_us_per_layer": roof["us_per_launch"], "attention_frac_of_hbm_peak": roof["frac"],
        "per_rank_bytes_per_step": step_bytes,
        "per_rank_hbm_floor_ms": step_bytes / (peak * 1e9) * 1e3,
    }
    del runner
    release(llm)
    return out


def tp_block(args, rank, world, K, W, peak, peak_src):
    """BASELINE configs[3]: Qwen3-32B decode, batch 256, ctx 1024, tensor-parallel over ALL `world`
    GPUs of this run, next to the same workload at TP=1 (rank 0 alone, the other ranks wait), so the
    driver's scaling record carries the north_star's "TP=8 >= 6x TP=1" figure."""
    import torch.distributed as dist
    from nanovllm import LLM
    B, ctx = args.batch, args.ctx
    kw = dict(kvcache_block_size=args.block_size, max_num_seqs=B, max_model_len=4096,
              max_num_batched_tokens=16384, distributed_launch="env", gpu_memory_utilization=0.88,
              graph_batch_sizes=(B,), seed=0)
    out = {"model": args.tp_model, "tp": world, "batch": B, "ctx": ctx}
    out.update(tp_engine_measure(args, rank, world, K, W, peak, peak_src, kw))
    dist.barrier()
    if rank == 0:
        llm = LLM(args.tp_model, tensor_parallel_size=1, **kw)
        r1 = decode_workload(llm, Group(1), B, ctx, K, W, 0, rank)
        out.update({"tp1_tokens_per_s": B * K / r1["dev_s"], "tp1_ms_per_step": r1["dev_s"] / K * 1e3,
                    "tp1_e2e_tokens_per_s": B * K / r1["e2e_s"]})
        out["speedup_vs_tp1"] = out["tokens_per_s"] / out["tp1_tokens_per_s"]
        out["e2e_speedup_vs_tp1"] = out["e2e_tokens_per_s"] / out["tp1_e2e_tokens_per_s"]
        release(llm)
    dist.barrier()
    return out


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

    from nanovllm import LLM
    K, W, B, ctx = args.steps, max(3, args.warmup), args.batch, args.ctx
    max_len = ctx + W + K + 16
    clocks = ClockSampler(local)
    clocks.start()
    llm = LLM(args.model, kvcache_block_size=args.block_size, max_num_seqs=B,
              max_model_len=max(4096, max_len), max_num_batched_tokens=16384,
              tensor_parallel_size=world if args.tp else 1, distributed_launch="env",
              gpu_memory_utilization=args.mem_util, kv_cache_dtype=args.kv_cache_dtype, seed=0)
    runner = llm.model_runner
    spec = runner.spec
    group = Group(world)
    dbg(rank, 'engine built')
    group.barrier()
    clocks.begin()
    r = decode_workload(llm, group, B, ctx, K, W, rank if not args.tp else 0, rank)
    peak, peak_src = peaks()
    fp8 = args.kv_cache_dtype != "auto"
    roof = attention_roofline(runner, group, B, K, world if args.tp else 1, fp8, peak, peak_src)
    dbg(rank, 'attn loop done')
    clk = clocks.stop()

    ttfts, dev_s, e2e_s = r["ttfts"], r["dev_s"], r["e2e_s"]
    prefill_s = ttfts[-1] / 1e3
    engines = 1 if args.tp else world
    value = engines * B * K / dev_s
    # KV-read roofline of the whole step at the measured peak (what `value` is bounded by)
    step_kv_bytes = roof["algorithmic_bytes_per_launch"] * roof["layers"] * roof["launches_per_layer"]
    roof["step_kv_roofline_tokens_per_s"] = B / (step_kv_bytes / (peak * 1e9))
    roof["step_frac_of_kv_roofline"] = (B * K / dev_s) / roof["step_kv_roofline_tokens_per_s"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_s / K * 1e3, "higher_is_better": True, "scaling": "weak" if not args.tp else "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic (random-init weights, random token ids)",
        "config": dict(workload_config(args, world), ctx_during_timing=[ctx + W, r["ctx_now"]],
                       decode_step="two micro-batches on two streams in one CUDA graph" if r["two_micro_batches"]
                       else "one batch"),
        "e2e": {"value": engines * B * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": r["h2d_per"],
                "d2h_bytes_per_step": r["d2h_per"], "ms_per_step": e2e_s / K * 1e3},
        "ttft_p50_ms": ttfts[len(ttfts) // 2], "ttft_max_ms": ttfts[-1], "prefill_steps": r["prefill_steps"],
        "prefill_step_ms": r["prefill_step_ms"], "submit_ms": r["submit_ms"],
        "prefill_tokens_per_s": engines * B * ctx / prefill_s,
        "gpu_launches": r["launches"],
        "roofline": roof,
        "clocks": clk,
    }
    if getattr(spec, "num_experts", 0):
        line["moe_roofline"] = moe_expert_roofline(runner, B, peak)
    dbg(rank, 'before exit')
    del runner
    release(llm)
    dbg(rank, 'after exit')
    if world == 1 and args.model == "qwen3-0.6b":
        try:
            line["prefill_roofline"] = prefill_roofline()
        except Exception as e:            # an extra object: its failure must not cost the contract line
            line["prefill_roofline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if world > 1 and not args.tp and not args.no_tp_block:
        # the replicas above say nothing about the collective design: the same run also measures the
        # tensor-parallel path (BASELINE configs[3]) over all `world` GPUs, with its TP=1 comparison
        def bail():                       # a hung collective must not cost the replica line
            line["tp"] = {"error": f"tensor-parallel block did not finish within {args.tp_timeout} s"}
            if rank == 0:
                print(json.dumps(line))
                sys.stdout.flush()
            os._exit(0)
        watchdog = threading.Timer(args.tp_timeout, bail)
        watchdog.daemon = True
        watchdog.start()
        try:
            line["tp"] = tp_block(args, rank, world, K, W, peak, peak_src)
        except Exception as e:            # symmetric failures only (OOM on every rank); reported, not hidden
            line["tp"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        watchdog.cancel()
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.model == "qwen3-0.6b":
        tps, sec, kind, cores = reference_cpu_step(args.cpu_sample_seqs, ctx, args.block_size, 3, 1, False)
        line["cpu_baseline"] = {"value": tps, "unit": UNIT, "cores": cores, "kind": kind,
                                "sample": cpu_sample_text(kind, args.cpu_sample_seqs, B, ctx, 3, sec)}
        line["config"]["cpu_sample_batch"] = args.cpu_sample_seqs
    if rank == 0:
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        # everything is measured and printed: a rank that failed inside the tp block must not turn the
        # orderly shutdown into a hang (destroy_process_group itself returns in < 1 s, see
        # profiles/r02_tp_teardown_probe.log)
        closer = threading.Timer(60.0, lambda: os._exit(0))
        closer.daemon = True
        closer.start()
        dist.barrier()
        dist.destroy_process_group()
        closer.cancel()


if __name__ == "__main__":
    main()
