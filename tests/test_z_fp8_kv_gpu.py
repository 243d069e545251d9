"""-m gpu: FP8 (e4m3) KV-cache variant of the scatter (K1/K2) and of paged decode attention (K6),
SURVEY.md 8f row f3, through the C ABI, against the oracle's statement of the same semantics
(oracle/attention_ref.py: quantize_e4m3 / decode_attention_fp8).

Tolerances: the quantising scatter is bit-exact (same fp32 product, same round-to-nearest-even
saturating cast as torch's float8_e4m3fn); attention is compared to the oracle's fp32 evaluation over
the DE-QUANTISED cache, |diff| <= 2e-2 on O(1) outputs as for the 16-bit kernel (K and V are exact in
the kernel's f16 operands, P is rounded to f16, the output to bf16).

This file sorts last on purpose and its first fixture runs one tiny launch of each kernel in a
child process with a timeout: a kernel that deadlocked on new hardware would otherwise hang the
whole GPU test session instead of failing one file.
"""
import os
import subprocess
import sys

import pytest
import torch

from oracle import attention_ref as A

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_PROBE = r"""
import sys, torch
sys.path[:0] = [r"%s", r"%s"]
from nanovllm import ops
G, D, bs, nb = 8, 128, 16, 6
k = torch.randn(40, G, D, device="cuda").bfloat16(); v = torch.randn(40, G, D, device="cuda").bfloat16()
kc = torch.zeros(nb, bs, G * D, dtype=torch.uint8, device="cuda"); vc = torch.zeros_like(kc)
slots = torch.arange(40, dtype=torch.int32, device="cuda")
ops.reshape_and_cache_fp8(k, v, kc, vc, slots)
q = torch.randn(2, 16, D, device="cuda").bfloat16()
bt = torch.tensor([[0, 1, 2], [1, 2, -1]], dtype=torch.int32, device="cuda")
ctx = torch.tensor([40, 17], dtype=torch.int32, device="cuda")
o = ops.paged_decode_attn_fp8(q, kc, vc, bt, ctx, D ** -0.5, G)
torch.cuda.synchronize()
assert torch.isfinite(o).all()
print("probe ok")
""" % (ROOT, os.path.join(ROOT, "nano-vllm-ascend_b200"))


@pytest.fixture(scope="module")
def fp8_alive():
    try:
        r = subprocess.run([sys.executable, "-c", _PROBE], capture_output=True, text=True, timeout=300)
    except subprocess.TimeoutExpired:
        pytest.fail("fp8 KV kernels did not finish a tiny launch within 300 s (deadlock?)")
    assert r.returncode == 0 and "probe ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def _ops():
    from nanovllm import ops
    return ops


@pytest.mark.parametrize("two_d", [False, True])
@pytest.mark.parametrize("scale", [(1.0, 1.0), (0.5, 2.0), (0.37, 1.9)])
@pytest.mark.parametrize("dtype", [BF, torch.float16])
def test_fp8_scatter_bit_exact(fp8_alive, two_d, scale, dtype):
    ops = _ops()
    ks, vs = scale
    g = torch.Generator().manual_seed(11)
    T, G, D, bs, nb = 300, 8, 128, 16, 40
    # wide dynamic range: subnormal e4m3 results, exact ties, saturation beyond 448
    key = (torch.randn(T, G, D, generator=g) * torch.logspace(-4, 3, T).view(T, 1, 1)).to(dtype)
    val = (torch.randn(T, G, D, generator=g) * 3).to(dtype)
    slots = torch.randperm(nb * bs, generator=g)[:T].to(torch.int32)
    slots[7], slots[100] = -1, nb * bs + 5                       # dropped rows
    kc = torch.full((nb, bs, G, D), 0x55, dtype=torch.uint8)
    vc = torch.full((nb, bs, G, D), 0x55, dtype=torch.uint8)
    want_k, want_v = kc.clone(), vc.clone()
    A.store_kvcache_fp8(key, val, want_k, want_v, slots, bs, ks, vs)
    d_kc, d_vc = kc.view(nb, bs, G * D).cuda(), vc.view(nb, bs, G * D).cuda()
    sl = torch.stack([slots // bs, slots % bs], 1).contiguous() if two_d else slots
    if two_d:
        sl[7] = torch.tensor([-1, 0], dtype=torch.int32)
        sl[100] = torch.tensor([nb + 1, 3], dtype=torch.int32)
    # strided views, as the engine passes them (slices of the qkv projection)
    packed = torch.cat([key.view(T, -1), val.view(T, -1)], 1).cuda()
    ops.reshape_and_cache_fp8(packed[:, :G * D].unflatten(1, (G, D)), packed[:, G * D:].unflatten(1, (G, D)),
                              d_kc, d_vc, sl.cuda(), ks, vs)
    got_k, got_v = d_kc.cpu().view(nb, bs, G, D), d_vc.cpu().view(nb, bs, G, D)
    # compare values (so that +0 / -0 encodings of a result rounded to zero cannot matter) ...
    assert torch.equal(A.dequantize_e4m3(got_k), A.dequantize_e4m3(want_k))
    assert torch.equal(A.dequantize_e4m3(got_v), A.dequantize_e4m3(want_v))
    # ... and the untouched bytes literally
    untouched = torch.ones(nb * bs, dtype=torch.bool)
    ok = (slots >= 0) & (slots < nb * bs)
    untouched[slots[ok].long()] = False
    assert (got_k.view(nb * bs, -1)[untouched] == 0x55).all()


def _fp8_case(B, H, G, D, bs, lens, seed, k_scale=1.0, v_scale=1.0):
    g = torch.Generator().manual_seed(seed)
    need = [(n + bs - 1) // bs for n in lens]
    nb = sum(need) + 3
    perm = torch.randperm(nb, generator=g).tolist()
    tables, p = [], 0
    for n in need:
        tables.append(perm[p:p + n])
        p += n
    width = max(1, max(need))
    bt = torch.tensor([t + [-1] * (width - len(t)) for t in tables], dtype=torch.int32)
    kc = A.quantize_e4m3(0.5 * torch.randn(nb, bs, G, D, generator=g), k_scale)
    vc = A.quantize_e4m3(torch.randn(nb, bs, G, D, generator=g), v_scale)
    q = torch.randn(B, H, D, generator=g).to(BF)
    return q, kc, vc, torch.tensor(lens, dtype=torch.int32), bt


def _run(q, kc, vc, ctx, bt, G, D, k_scale=1.0, v_scale=1.0, **kw):
    nb, bs = kc.shape[0], kc.shape[1]
    return _ops().paged_decode_attn_fp8(q.cuda(), kc.cuda().view(nb, bs, G * D), vc.cuda().view(nb, bs, G * D),
                                        bt.cuda(), ctx.cuda(), D ** -0.5, G, k_scale, v_scale, **kw).cpu()


@pytest.mark.parametrize("H,G,D,bs", [
    (16, 8, 128, 16),    # Qwen3-0.6B, BASELINE block size
    (16, 8, 128, 256),
    (8, 1, 128, 16),     # Qwen3-32B at TP=8 (per-rank heads): 8 warps share the head
    (32, 8, 64, 16),     # Llama-3.2-1B: two heads per 128-byte panel row
    (8, 2, 64, 32),
    (16, 4, 128, 64),
    (16, 16, 128, 16),   # two head groups per sequence
    (14, 2, 64, 16),     # group of 7
])
def test_fp8_decode_ragged(fp8_alive, H, G, D, bs):
    lens = [1, 15, 16, 17, 127, 128, 129, 300, 1024, 0, 33, 2, 511]
    q, kc, vc, ctx, bt = _fp8_case(len(lens), H, G, D, bs, lens, seed=H * 7 + bs)
    want = A.decode_attention_fp8(q, kc, vc, ctx, bt, G, D, D ** -0.5)
    for ctas in (0, 3, 148, 296):
        o = _run(q, kc, vc, ctx, bt, G, D, num_ctas=ctas)
        err = (o.float() - want).abs().max()
        assert err < 2e-2, f"num_ctas={ctas}: {err}"
        assert torch.count_nonzero(o[9]) == 0          # context_len == 0 row -> zeros


def test_fp8_decode_scales_and_stale_nan_rows(fp8_alive):
    """Per-tensor scales; stale rows past the context hold the e4m3 NaN pattern and must not leak."""
    lens, G, D, bs = [5, 40, 77, 130], 2, 128, 16
    q, kc, vc, ctx, bt = _fp8_case(len(lens), 8, G, D, bs, lens, seed=5, k_scale=0.25, v_scale=3.0)
    want = A.decode_attention_fp8(q, kc, vc, ctx, bt, G, D, D ** -0.5, 0.25, 3.0)
    for i, n in enumerate(lens):
        blk, off = int(bt[i, (n - 1) // bs]), n % bs
        if off:
            kc[blk, off:] = 0x7f
            vc[blk, off:] = 0xff
    o = _run(q, kc, vc, ctx, bt, G, D, 0.25, 3.0)
    assert torch.isfinite(o).all()
    assert (o.float() - want).abs().max() < 6e-2      # outputs are O(3): same relative bound


def test_fp8_decode_many_sequences(fp8_alive):
    g = torch.Generator().manual_seed(77)
    lens = torch.randint(1, 300, (96,), generator=g).tolist()
    lens[5], lens[40], lens[41] = 3000, 0, 1
    q, kc, vc, ctx, bt = _fp8_case(len(lens), 16, 8, 128, 16, lens, seed=123)
    want = A.decode_attention_fp8(q, kc, vc, ctx, bt, 8, 128, 128 ** -0.5)
    for ctas in (0, 7, 148, 300):
        o = _run(q, kc, vc, ctx, bt, 8, 128, num_ctas=ctas)
        assert (o.float() - want).abs().max() < 2e-2, ctas
        assert torch.count_nonzero(o[40]) == 0


def test_fp8_engine_tracks_bf16_engine(fp8_alive):
    """End to end through LLM (config flag -> uint8 cache -> quantising scatter -> fp8 decode inside
    the CUDA graph). The first generated token comes from the prefill, which attends over the
    un-quantised K/V of the step, so it must be identical to the 16-bit-cache run; later tokens
    read the e4m3 cache and may only drift where logits nearly tie."""
    import random
    from nanovllm import LLM, SamplingParams
    rnd = random.Random(4)
    prompts = [[rnd.randint(0, 511) for _ in range(n)] for n in (9, 40, 17, 33)]
    outs = {}
    for kvd in ("auto", "fp8"):
        llm = LLM("tiny-qwen3", seed=3, norm_jitter=0.1, kvcache_block_size=16, max_num_seqs=8,
                  max_model_len=256, max_num_batched_tokens=1024, num_kvcache_blocks=64,
                  kv_cache_dtype=kvd)
        res = llm.generate(prompts, SamplingParams(temperature=1.0, max_tokens=8, ignore_eos=True, greedy=True),
                           use_tqdm=False)
        outs[kvd] = [o["token_ids"] for o in res]
        llm.exit()
    assert [o[0] for o in outs["auto"]] == [o[0] for o in outs["fp8"]]
    assert all(len(o) == 8 and all(0 <= t < 512 for t in o) for o in outs["fp8"])
    agree = sum(x == y for a, b in zip(outs["auto"], outs["fp8"]) for x, y in zip(a[:4], b[:4]))
    assert agree >= 8, outs          # of 16 positions; e4m3 noise (2^-4 relative per element) flips few


@pytest.mark.parametrize("G,D,bs,dtype,scales", [(8, 128, 16, BF, (1.0, 1.0)), (2, 64, 32, torch.float16, (0.25, 3.0)),
                                                  (1, 128, 256, BF, (2.0, 0.5))])
def test_fp8_gather_dequant_bit_exact(fp8_alive, G, D, bs, dtype, scales):
    """b200_gather_dequant_fp8: the cached tokens of every sequence, gathered through its block table and
    de-quantised -- bit-exact against the oracle's dequantize_e4m3 (exact fp32 product, one rounding)."""
    lens = [1, 17, 300, 64, 5]
    _, kc, vc, _, bt = _fp8_case(len(lens), G, G, D, bs, lens, seed=bs + D, k_scale=scales[0], v_scale=scales[1])
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
    nb = kc.shape[0]
    k, v = _ops().gather_dequant_fp8(kc.cuda().view(nb, bs, G * D), vc.cuda().view(nb, bs, G * D), bt.cuda(),
                                     cu.cuda(), max(lens), G, D, scales[0], scales[1], dtype=dtype)
    assert k.shape[0] == len(lens) * max(lens)
    kd, vd = A.dequantize_e4m3(kc, scales[0], dtype), A.dequantize_e4m3(vc, scales[1], dtype)
    for i, n in enumerate(lens):
        for t in (0, n // 2, n - 1):
            blk = int(bt[i, t // bs])
            assert torch.equal(k[int(cu[i]) + t].cpu(), kd[blk, t % bs]), (i, t)
            assert torch.equal(v[int(cu[i]) + t].cpu(), vd[blk, t % bs]), (i, t)
    assert torch.count_nonzero(k[int(cu[-1]):]) == 0          # rows past the last sequence stay zero


def test_fp8_chunked_and_prefix_hit_prefill(fp8_alive):
    """kv_cache_dtype="fp8" with chunked prefill / true prefix-hit prefill (r01 rejected the combination):
    later windows read the earlier ones back from the e4m3 cache through gather + de-quantise. Same
    workload as the 16-bit chunked-prefill test; tokens may only drift from the un-chunked fp8 run where
    logits nearly tie (the un-chunked prefill attends over un-quantised K/V of its own step)."""
    import random
    from nanovllm import LLM, SamplingParams
    rnd = random.Random(12)
    prompts = [[rnd.randint(0, 511) for _ in range(n)] for n in (300, 40, 17, 200, 129)]
    prompts.append(prompts[0][:160] + [3] * 30)            # shares 10 full blocks with the first prompt
    sp = SamplingParams(temperature=1.0, max_tokens=8, ignore_eos=True, greedy=True)
    outs = []
    for kw in (dict(max_num_batched_tokens=1024), dict(max_num_batched_tokens=128, chunked_prefill=True),
               dict(max_num_batched_tokens=1024, prefix_hit_prefill=True)):
        llm = LLM("tiny-qwen3", seed=3, norm_jitter=0.1, kvcache_block_size=16, max_num_seqs=8,
                  max_model_len=512, num_kvcache_blocks=128, kv_cache_dtype="fp8", **kw)
        outs.append([o["token_ids"] for o in llm.generate(prompts, sp, use_tqdm=False)])
        llm.exit()
    for other in outs[1:]:
        assert all(len(o) == 8 and all(0 <= t < 512 for t in o) for o in other)
        agree = sum(x == y for a, b in zip(outs[0], other) for x, y in zip(a[:3], b[:3]))
        assert agree >= 12, (agree, outs[0], other)          # of 18 positions
