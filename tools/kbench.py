"""Developer micro-benchmark of single kernels (not the contract bench; see bench.py).
usage: python tools/kbench.py [decode|prefill|fa2|all]

`fa2` times the reference's CUDA attention variant (layers/attention_ori.py:63-74: flash_attn 2.8.3
flash_attn_with_kvcache / flash_attn_varlen_func, library code) on the same shapes -- the
"kernel to beat" of SURVEY.md 8(d). FA2's paged path needs page_block_size % 256 == 0, so it runs
at block 256 (the reference default); ours is timed at block 16 and 256."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nano-vllm-ascend_b200"))
from nanovllm import ops  # noqa: E402


def timeit(fn, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2], ts[0]


def bench_decode(B, H, G, D, bs, ctx, ctas=0):
    g = torch.Generator(device="cuda").manual_seed(0)
    nblk = B * ((ctx + bs - 1) // bs)
    kc = torch.randn(nblk + 1, bs, G * D, device="cuda", generator=g).to(torch.bfloat16)
    vc = torch.randn(nblk + 1, bs, G * D, device="cuda", generator=g).to(torch.bfloat16)
    q = torch.randn(B, H, D, device="cuda", generator=g).to(torch.bfloat16)
    bt = torch.randperm(nblk, device="cuda", generator=g).to(torch.int32).view(B, -1).contiguous()
    cl = torch.full((B,), ctx, dtype=torch.int32, device="cuda")
    out = torch.empty(B, H * D, dtype=torch.bfloat16, device="cuda")
    ws = ops.decode_workspace(B, H, G, D, q.device)
    fn = lambda: ops.paged_decode_attn(q, kc, vc, bt, cl, D ** -0.5, G, out=out, workspace=ws,
                                       num_ctas=ctas)
    med, best = timeit(fn)
    byts = B * ctx * 2 * G * D * 2
    return dict(kernel="paged_decode", B=B, H=H, G=G, D=D, bs=bs, ctx=ctx, ctas=ctas,
                ms_med=round(med, 4), ms_best=round(best, 4), GBps_med=round(byts / med / 1e6, 1),
                GBps_best=round(byts / best / 1e6, 1))


def bench_prefill(nseq, L, H, G, D):
    g = torch.Generator(device="cuda").manual_seed(0)
    T = nseq * L
    qkv = torch.randn(T, (H + 2 * G) * D, device="cuda", generator=g).to(torch.bfloat16)
    q = qkv[:, :H * D].unflatten(1, (H, D))
    k = qkv[:, H * D:(H + G) * D].unflatten(1, (G, D))
    v = qkv[:, (H + G) * D:].unflatten(1, (G, D))
    cu = torch.arange(0, T + 1, L, dtype=torch.int32, device="cuda")
    out = torch.empty(T, H * D, dtype=torch.bfloat16, device="cuda")
    fn = lambda: ops.prefill_varlen_attn(q, k, v, cu, cu, L, L, D ** -0.5, G, out=out)
    med, best = timeit(fn, iters=10, warmup=3)
    flops = 2 * L * L * H * D * nseq  # causal: 4*L*L*H*D/2
    return dict(kernel="prefill", nseq=nseq, L=L, H=H, G=G, D=D, ms_med=round(med, 4),
                TFLOPs_med=round(flops / med / 1e9, 1), TFLOPs_best=round(flops / best / 1e9, 1))


def bench_prefill_prefix_hit(qlen, klen, H, G, D, bs):
    """Config-3 form: a prompt whose first klen-qlen tokens hit the prefix cache: q_len << k_len,
    K/V read from the paged cache through block_tables (attention_ori.py:65-66)."""
    g = torch.Generator(device="cuda").manual_seed(0)
    nblk = (klen + bs - 1) // bs
    kc = torch.randn(nblk + 1, bs, G * D, device="cuda", generator=g).to(torch.bfloat16)
    vc = torch.randn(nblk + 1, bs, G * D, device="cuda", generator=g).to(torch.bfloat16)
    q = torch.randn(qlen, H, D, device="cuda", generator=g).to(torch.bfloat16)
    bt = torch.randperm(nblk, device="cuda", generator=g).to(torch.int32).view(1, -1).contiguous()
    cu_q = torch.tensor([0, qlen], dtype=torch.int32, device="cuda")
    cu_k = torch.tensor([0, klen], dtype=torch.int32, device="cuda")
    out = torch.empty(qlen, H * D, dtype=torch.bfloat16, device="cuda")
    fn = lambda: ops.prefill_varlen_attn(q, kc, vc, cu_q, cu_k, qlen, klen, D ** -0.5, G, block_tables=bt, out=out)
    med, best = timeit(fn, iters=10, warmup=3)
    flops = 4 * H * D * (qlen * (klen - qlen) + qlen * qlen / 2)
    return dict(kernel="prefill_prefix_hit", qlen=qlen, klen=klen, H=H, G=G, D=D, bs=bs, ms_med=round(med, 4),
                TFLOPs_med=round(flops / med / 1e9, 1))


def bench_fa2_decode(B, H, G, D, ctx, bs=256):
    """attention_ori.py:72-74 -- flash_attn_with_kvcache over a paged cache."""
    from flash_attn import flash_attn_with_kvcache
    g = torch.Generator(device="cuda").manual_seed(0)
    per = (ctx + bs - 1) // bs
    nblk = B * per
    kc = torch.randn(nblk + 1, bs, G, D, device="cuda", generator=g).to(torch.bfloat16)
    vc = torch.randn(nblk + 1, bs, G, D, device="cuda", generator=g).to(torch.bfloat16)
    q = torch.randn(B, 1, H, D, device="cuda", generator=g).to(torch.bfloat16)
    bt = torch.randperm(nblk, device="cuda", generator=g).to(torch.int32).view(B, per).contiguous()
    cl = torch.full((B,), ctx, dtype=torch.int32, device="cuda")
    fn = lambda: flash_attn_with_kvcache(q, kc, vc, cache_seqlens=cl, block_table=bt,
                                         softmax_scale=D ** -0.5, causal=True)
    med, best = timeit(fn)
    byts = B * ctx * 2 * G * D * 2
    return dict(kernel="fa2_with_kvcache", B=B, H=H, G=G, D=D, bs=bs, ctx=ctx, ms_med=round(med, 4),
                GBps_med=round(byts / med / 1e6, 1), GBps_best=round(byts / best / 1e6, 1))


def bench_fa2_prefill(nseq, L, H, G, D):
    """attention_ori.py:67-70 -- flash_attn_varlen_func, causal, packed K/V."""
    from flash_attn import flash_attn_varlen_func
    g = torch.Generator(device="cuda").manual_seed(0)
    T = nseq * L
    q = torch.randn(T, H, D, device="cuda", generator=g).to(torch.bfloat16)
    k = torch.randn(T, G, D, device="cuda", generator=g).to(torch.bfloat16)
    v = torch.randn(T, G, D, device="cuda", generator=g).to(torch.bfloat16)
    cu = torch.arange(0, T + 1, L, dtype=torch.int32, device="cuda")
    fn = lambda: flash_attn_varlen_func(q, k, v, max_seqlen_q=L, cu_seqlens_q=cu, max_seqlen_k=L,
                                        cu_seqlens_k=cu, softmax_scale=D ** -0.5, causal=True)
    med, best = timeit(fn, iters=10, warmup=3)
    flops = 2 * L * L * H * D * nseq
    return dict(kernel="fa2_varlen", nseq=nseq, L=L, H=H, G=G, D=D, ms_med=round(med, 4),
                TFLOPs_med=round(flops / med / 1e9, 1), TFLOPs_best=round(flops / best / 1e9, 1))


def bench_fa2_prefix_hit(qlen, klen, H, G, D, bs=256):
    """attention_ori.py:65-66 -- flash_attn_varlen_func with block_table (prefix-cache hit: q shorter
    than k, K/V from the paged cache); FA2 pages are 256 tokens."""
    from flash_attn import flash_attn_varlen_func
    g = torch.Generator(device="cuda").manual_seed(0)
    nblk = (klen + bs - 1) // bs
    kc = torch.randn(nblk + 1, bs, G, D, device="cuda", generator=g).to(torch.bfloat16)
    vc = torch.randn(nblk + 1, bs, G, D, device="cuda", generator=g).to(torch.bfloat16)
    q = torch.randn(qlen, H, D, device="cuda", generator=g).to(torch.bfloat16)
    bt = torch.randperm(nblk, device="cuda", generator=g).to(torch.int32).view(1, -1).contiguous()
    cu_q = torch.tensor([0, qlen], dtype=torch.int32, device="cuda")
    cu_k = torch.tensor([0, klen], dtype=torch.int32, device="cuda")
    fn = lambda: flash_attn_varlen_func(q, kc, vc, max_seqlen_q=qlen, cu_seqlens_q=cu_q, max_seqlen_k=klen,
                                        cu_seqlens_k=cu_k, softmax_scale=D ** -0.5, causal=True,
                                        block_table=bt)
    med, best = timeit(fn, iters=10, warmup=3)
    flops = 4 * H * D * (qlen * (klen - qlen) + qlen * qlen / 2)
    return dict(kernel="fa2_varlen_prefix_hit", qlen=qlen, klen=klen, H=H, G=G, D=D, bs=bs,
                ms_med=round(med, 4), TFLOPs_med=round(flops / med / 1e9, 1))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="all")
    args = ap.parse_args()
    res = []
    if args.what in ("decode", "all"):
        for bs in (16, 256):
            res.append(bench_decode(256, 16, 8, 128, bs, 1024))
        res.append(bench_decode(256, 16, 8, 128, 16, 1024, ctas=296))
        res.append(bench_decode(256, 8, 1, 128, 16, 1024))      # 32B TP=8 per-rank shape
        res.append(bench_decode(256, 32, 8, 64, 16, 1024))      # Llama-1B
        res.append(bench_decode(1, 16, 8, 128, 16, 8192))
        res.append(bench_decode(16, 16, 8, 128, 16, 1024))
    if args.what == "prefill16":                 # one shape, few launches: the ncu target
        res.append(bench_prefill(16, 1024, 16, 8, 128))
    if args.what in ("prefill", "all"):
        res.append(bench_prefill(16, 1024, 16, 8, 128))
        res.append(bench_prefill(4, 4096, 16, 8, 128))
        res.append(bench_prefill(1, 8192, 32, 8, 64))
        res.append(bench_prefill_prefix_hit(256, 8192, 32, 8, 64, 16))
    if args.what in ("fa2", "all"):
        try:
            res.append(bench_fa2_decode(256, 16, 8, 128, 1024))
            res.append(bench_fa2_decode(256, 8, 1, 128, 1024))
            res.append(bench_fa2_decode(256, 32, 8, 64, 1024))
            res.append(bench_fa2_prefill(16, 1024, 16, 8, 128))
            res.append(bench_fa2_prefill(4, 4096, 16, 8, 128))
            res.append(bench_fa2_prefill(1, 8192, 32, 8, 64))
            res.append(bench_fa2_prefix_hit(256, 8192, 32, 8, 64))
        except Exception as e:                     # the comparator is optional library code
            res.append(dict(kernel="fa2", unavailable=f"{type(e).__name__}: {e}"[:200]))
    for r in res:
        print(json.dumps(r))
