"""-m gpu parity tests proper: every CUDA kernel, called through the C ABI (ctypes), against
the oracle on the same seeded inputs and against the reference-generated golden fixtures.

Tolerances (stated per test):
  * KV scatter: bit-exact (byte movement).
  * RMSNorm / RoPE / SiluAndMul / K8 epilogue: same rounding points as the reference eager ops;
    the only freedom is fp32 summation order / expf ulps => at most 1 bf16 ulp on a tiny
    fraction of elements.
  * attention: kernels accumulate in fp32 where the reference's torch-native path rounds every
    matmul to bf16. Compared (a) to the oracle's fp32 ("exact") evaluation of the same function
    at bf16-output resolution: |diff| <= 2e-2 absolute on O(1) outputs, and (b) to the
    reference-arithmetic outputs with the looser 6e-2 bound the two CPU evaluations differ by.
"""
import math

import pytest
import torch

from oracle import attention_ref as A
from oracle import layers_ref as L

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _ops():
    from nanovllm import ops
    return ops


def _ulp_close(got, want, max_frac=2e-3):
    """Equal up to 2 units in the last place of bf16/fp16 on at most max_frac of the elements."""
    got, want = got.cpu(), want.cpu()
    if torch.equal(got, want):
        return True
    gi = got.view(torch.int16).int()
    wi = want.view(torch.int16).int()
    d = (gi - wi).abs()
    # two rounding points (normalise, then scale by the weight): a 1-ulp flip at the first can
    # show as 2 ulps after the second
    assert int(d.max()) <= 2, f"max ulp distance {int(d.max())}"
    frac = float((d > 0).float().mean())
    assert frac <= max_frac, f"{frac:.2e} of elements differ by <= 2 ulp"
    return True


# ------------------------------------------------------------------ KV scatter (bit-exact)
@pytest.mark.parametrize("two_d", [False, True])
def test_kv_scatter_golden(golden_kernels, two_d):
    ops = _ops()
    g = golden_kernels["kv_scatter"]
    bs = g["block_size"]
    kc, vc = g["k_cache"].cuda().clone(), g["v_cache"].cuda().clone()
    slots = g["slot_mapping"].clone()
    if two_d:
        nb = kc.shape[0]
        blk = torch.where(slots >= 0, slots // bs, torch.full_like(slots, -1))
        pairs = torch.stack([blk, torch.where(slots >= 0, slots % bs, torch.zeros_like(slots))], 1)
        slots = pairs.to(torch.int32).contiguous()
    ops.reshape_and_cache(g["key"].cuda(), g["value"].cuda(), kc, vc, slots.cuda())
    assert torch.equal(kc.cpu(), g["k_cache_out"]) and torch.equal(vc.cpu(), g["v_cache_out"])


@pytest.mark.parametrize("T,G,D,bs", [(1, 8, 128, 16), (256, 8, 128, 16), (4099, 1, 128, 256),
                                      (777, 4, 64, 32), (0, 2, 64, 16)])
def test_kv_scatter_random(T, G, D, bs):
    ops = _ops()
    g = torch.Generator().manual_seed(T + G)
    nb = max(2, (T + bs - 1) // bs + 3)
    # strided views, as the model slices k/v out of the fused qkv activation
    big = torch.randn(max(T, 1), (3 * G) * D, generator=g).to(BF)
    key = big[:T, G * D:2 * G * D].unflatten(1, (G, D))
    val = big[:T, 2 * G * D:].unflatten(1, (G, D))
    kc = torch.randn(nb, bs, G, D, generator=g).to(BF)
    vc = torch.randn(nb, bs, G, D, generator=g).to(BF)
    slots = torch.randperm(nb * bs, generator=g)[:T].to(torch.int32)
    if T > 10:
        slots[::7] = -1
    kc_ref, vc_ref = kc.clone(), vc.clone()
    A.store_kvcache(key, val, kc_ref, vc_ref, slots, bs)
    kc_d, vc_d = kc.cuda(), vc.cuda()
    ops.reshape_and_cache(big.cuda()[:T, G * D:2 * G * D].unflatten(1, (G, D)),
                          big.cuda()[:T, 2 * G * D:].unflatten(1, (G, D)),
                          kc_d.view(nb, bs, G * D), vc_d.view(nb, bs, G * D), slots.cuda())
    assert torch.equal(kc_d.cpu(), kc_ref) and torch.equal(vc_d.cpu(), vc_ref)


# ------------------------------------------------------------------ eager-layer kernels
def test_layers_golden(golden_kernels):
    ops = _ops()
    for hidden in (128, 1024):
        g = golden_kernels[f"rmsnorm_{hidden}"]
        y = ops.rmsnorm(g["x"].cuda(), g["weight"].cuda(), g["eps"])
        _ulp_close(y, g["out"])
        x, r = g["x"].cuda().clone(), g["residual"].cuda().clone()
        ops.add_rmsnorm_(x, r, g["weight"].cuda(), g["eps"])
        _ulp_close(x, g["add_out"])
        assert torch.equal(r.cpu(), g["add_residual"])
    g = golden_kernels["silu_and_mul"]
    _ulp_close(ops.silu_and_mul(g["x"].cuda()), g["out"])


@pytest.mark.parametrize("rows,hidden", [(1, 1024), (256, 1024), (33, 5120), (3, 2048), (5, 8192),
                                         (2, 64)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_rmsnorm_random(rows, hidden, dtype):
    ops = _ops()
    g = torch.Generator().manual_seed(rows * hidden)
    x = (2 * torch.randn(rows, hidden, generator=g)).to(dtype)
    r = torch.randn(rows, hidden, generator=g).to(dtype)
    w = (1 + 0.2 * torch.randn(hidden, generator=g)).to(dtype)
    _ulp_close(ops.rmsnorm(x.cuda(), w.cuda(), 1e-6), L.rmsnorm(x, w, 1e-6))
    want, want_r = L.add_rmsnorm(x, r, w, 1e-6)
    xd, rd = x.cuda(), r.cuda()
    ops.add_rmsnorm_(xd, rd, w.cuda(), 1e-6)
    _ulp_close(xd, want)
    assert torch.equal(rd.cpu(), want_r)


@pytest.mark.parametrize("rows,inter", [(1, 3072), (256, 3072), (7, 25600), (3, 8)])
def test_silu_and_mul_random(rows, inter):
    ops = _ops()
    g = torch.Generator().manual_seed(rows + inter)
    x = (3 * torch.randn(rows, 2 * inter, generator=g)).to(BF)
    _ulp_close(ops.silu_and_mul(x.cuda()), L.silu_and_mul(x))


@pytest.mark.parametrize("H,G,D", [(16, 8, 128), (8, 1, 128), (32, 8, 64), (4, 2, 64)])
@pytest.mark.parametrize("qk_norm", [True, False])
@pytest.mark.parametrize("slot_2d", [False, True])
def test_qknorm_rope_cache(H, G, D, qk_norm, slot_2d):
    """K8 against the chain the reference runs op by op (qwen3.py:79-88)."""
    ops = _ops()
    g = torch.Generator().manual_seed(H * 131 + D)
    T, nb, bs, maxpos = 37, 9, 16, 512
    qkv = torch.randn(T, (H + 2 * G) * D, generator=g).to(BF)
    pos = torch.randint(0, maxpos, (T,), generator=g)
    qw = (1 + 0.1 * torch.randn(D, generator=g)).to(BF) if qk_norm else None
    kw = (1 + 0.1 * torch.randn(D, generator=g)).to(BF) if qk_norm else None
    cache = L.rope_cache(D, maxpos, 1000000.0)
    slots = torch.randperm(nb * bs, generator=g)[:T].to(torch.int32)
    slots[5] = -1
    q, k, v = L.qkv_epilogue(qkv, pos, cache, qw, kw, 1e-6, H, G, D)
    kc = torch.zeros(nb, bs, G, D, dtype=BF)
    vc = torch.zeros_like(kc)
    A.store_kvcache(k, v, kc, vc, slots, bs)

    qkv_d = qkv.cuda()
    kc_d = torch.zeros(nb, bs, G * D, dtype=BF, device="cuda")
    vc_d = torch.zeros_like(kc_d)
    if slot_2d:
        blk = torch.where(slots >= 0, slots // bs, torch.full_like(slots, -1))
        sl = torch.stack([blk, torch.where(slots >= 0, slots % bs, torch.zeros_like(slots))], 1)
        sl = sl.to(torch.int32).contiguous().cuda()
    else:
        sl = slots.cuda()
    ops.qknorm_rope_cache(qkv_d, pos.cuda(), cache.squeeze(1).contiguous().cuda(),
                          None if qw is None else qw.cuda(), None if kw is None else kw.cuda(),
                          1e-6, kc_d, vc_d, sl, H, G, D)
    got = qkv_d.cpu()
    _ulp_close(got[:, :H * D].reshape(T, H, D), q.contiguous())
    _ulp_close(got[:, H * D:(H + G) * D].reshape(T, G, D), k.contiguous())
    assert torch.equal(got[:, (H + G) * D:].reshape(T, G, D), v)       # v untouched
    _ulp_close(kc_d.cpu().view(nb, bs, G, D), kc)
    assert torch.equal(vc_d.cpu().view(nb, bs, G, D), vc)


def test_rope_golden(golden_kernels):
    """RoPE alone (no QK-norm) on the reference-generated vectors."""
    ops = _ops()
    for D in (64, 128):
        g = golden_kernels[f"rope_{D}"]
        T = g["q"].shape[0]
        qkv = torch.cat([g["q"].reshape(T, -1), g["k"].reshape(T, -1),
                         torch.zeros(T, 2 * D, dtype=BF)], 1).cuda()
        ops.qknorm_rope_cache(qkv, g["positions"].cuda(),
                              g["cos_sin_cache"].squeeze(1).contiguous().cuda(), None, None, 1e-6,
                              None, None, None, 4, 2, D)
        got = qkv.cpu()
        assert torch.equal(got[:, :4 * D].reshape(T, 4, D), g["q_out"])
        assert torch.equal(got[:, 4 * D:6 * D].reshape(T, 2, D), g["k_out"])


# ------------------------------------------------------------------ paged decode attention
def _decode_case(B, H, G, D, bs, lens, seed, dtype=BF, extra_blocks=3):
    g = torch.Generator().manual_seed(seed)
    need = [(n + bs - 1) // bs for n in lens]
    nb = sum(need) + extra_blocks
    perm = torch.randperm(nb, generator=g).tolist()
    tables, p = [], 0
    for n in need:
        tables.append(perm[p:p + n])
        p += n
    width = max(1, max(need))
    bt = torch.tensor([t + [-1] * (width - len(t)) for t in tables], dtype=torch.int32)
    kc = (0.5 * torch.randn(nb, bs, G, D, generator=g)).to(dtype)
    vc = torch.randn(nb, bs, G, D, generator=g).to(dtype)
    q = torch.randn(B, H, D, generator=g).to(dtype)
    return q, kc, vc, torch.tensor(lens, dtype=torch.int32), bt


def _run_decode(q, kc, vc, ctx, bt, G, D, **kw):
    ops = _ops()
    nb, bs = kc.shape[0], kc.shape[1]
    return ops.paged_decode_attn(q.cuda(), kc.cuda().view(nb, bs, G * D), vc.cuda().view(nb, bs, G * D),
                                 bt.cuda(), ctx.cuda(), D ** -0.5, G, **kw).cpu()


def test_decode_golden(golden_kernels):
    for tag in ("d64", "d128"):
        g = golden_kernels[f"decode_{tag}"]
        G, D = g["num_kv_heads"], g["head_dim"]
        o = _run_decode(g["q"], g["k_cache"], g["v_cache"], g["context_lens"], g["block_tables"], G, D)
        exact = A.decode_attention(g["q"], g["k_cache"], g["v_cache"], g["context_lens"],
                                   g["block_tables"], G, D, mode="exact", scale=D ** -0.5)
        assert (o.float() - exact).abs().max() < 2e-2
        assert (o.float() - g["out"].float()).abs().max() < 6e-2


@pytest.mark.parametrize("H,G,D,bs", [
    (16, 8, 128, 16),    # Qwen3-0.6B, BASELINE block size
    (16, 8, 128, 256),   # reference default block size
    (8, 1, 128, 16),     # Qwen3-32B at TP=8 (per-rank heads)
    (8, 1, 128, 128),
    (32, 8, 64, 16),     # Llama-3.2-1B
    (8, 2, 64, 32),
    (16, 4, 128, 64),
    (16, 16, 128, 16),   # MHA, two head groups per sequence
    (14, 2, 64, 16),     # Qwen2-0.5B (group of 7)
])
def test_decode_ragged(H, G, D, bs):
    lens = [1, 15, 16, 17, 127, 128, 129, 300, 1024, 0, 33, 2, 511]
    q, kc, vc, ctx, bt = _decode_case(len(lens), H, G, D, bs, lens, seed=H * 7 + bs)
    want = A.decode_attention(q, kc, vc, ctx, bt, G, D, mode="exact", scale=D ** -0.5)
    for ctas in (0, 3, 148, 296):   # auto grid, many boundaries, ... : stream-K partition + combine
        o = _run_decode(q, kc, vc, ctx, bt, G, D, num_ctas=ctas)
        err = (o.float() - want).abs().max()
        assert err < 2e-2, f"num_ctas={ctas}: {err}"
        assert torch.count_nonzero(o[9]) == 0          # context_len == 0 row -> zeros


def test_decode_many_sequences_ragged():
    """>= 64 sequences (the one-block-per-boundary combine mode), ragged lengths incl. one sequence
    far longer than the per-CTA share, several persistent grid sizes."""
    g = torch.Generator().manual_seed(77)
    lens = torch.randint(1, 300, (96,), generator=g).tolist()
    lens[5], lens[40], lens[41] = 3000, 0, 1
    q, kc, vc, ctx, bt = _decode_case(len(lens), 16, 8, 128, 16, lens, seed=123)
    want = A.decode_attention(q, kc, vc, ctx, bt, 8, 128, mode="exact", scale=128 ** -0.5)
    for ctas in (0, 7, 148, 300):
        o = _run_decode(q, kc, vc, ctx, bt, 8, 128, num_ctas=ctas)
        assert (o.float() - want).abs().max() < 2e-2, ctas
        assert torch.count_nonzero(o[40]) == 0


def test_decode_fp16_and_stale_nan_rows():
    """fp16 caches; stale rows past the context hold NaN bit patterns and must not leak."""
    lens = [5, 40, 77]
    q, kc, vc, ctx, bt = _decode_case(3, 8, 2, 128, 16, lens, seed=5, dtype=torch.float16)
    want = A.decode_attention(q, kc, vc, ctx, bt, 2, 128, mode="exact", scale=128 ** -0.5)
    for i, n in enumerate(lens):     # poison everything after the context in the last block
        blk, off = int(bt[i, (n - 1) // 16]), n % 16
        if off:
            kc[blk, off:] = float("nan")
            vc[blk, off:] = float("nan")
    o = _run_decode(q, kc, vc, ctx, bt, 2, 128)
    assert torch.isfinite(o).all()
    assert (o.float() - want).abs().max() < 2e-2


def test_decode_full_size_properties():
    """BASELINE config-2 shape (B=256, ctx=1024, block 16): size-independent properties.
    (a) permuting the physical blocks + block_tables leaves the output bit-identical;
    (b) V := 1 gives exactly 1 (softmax sums to one); (c) 40 rows spread over the batch (every
    8th + the ends, ragged context lengths in a second launch) match the oracle's fp32 evaluation."""
    ops = _ops()
    B, H, G, D, bs, n = 256, 16, 8, 128, 16, 1024
    g = torch.Generator(device="cuda").manual_seed(0)
    nblk = B * n // bs
    kc = (0.5 * torch.randn(nblk + 1, bs, G * D, generator=g, device="cuda")).to(BF)
    vc = torch.randn(nblk + 1, bs, G * D, generator=g, device="cuda").to(BF)
    q = torch.randn(B, H, D, generator=g, device="cuda").to(BF)
    perm = torch.randperm(nblk, device="cuda", generator=g).to(torch.int32)
    bt = perm.view(B, n // bs).contiguous()
    ctx = torch.full((B,), n, dtype=torch.int32, device="cuda")
    o1 = ops.paged_decode_attn(q, kc, vc, bt, ctx, D ** -0.5, G)
    # (a) physically shuffle the blocks
    perm2 = torch.randperm(nblk, device="cuda", generator=g)
    kc2, vc2 = torch.empty_like(kc), torch.empty_like(vc)
    kc2[perm2] = kc[:nblk]
    vc2[perm2] = vc[:nblk]
    bt2 = perm2[bt.long()].to(torch.int32)
    o2 = ops.paged_decode_attn(q, kc2, vc2, bt2, ctx, D ** -0.5, G)
    assert torch.equal(o1, o2)
    # (b)
    ones = torch.ones_like(vc)
    o3 = ops.paged_decode_attn(q, kc, ones, bt, ctx, D ** -0.5, G)
    assert (o3.float() - 1).abs().max() < 1e-2
    # (c)
    rows = sorted(set(range(0, B, 8)) | {1, 101, 127, 128, 129, 254, 255})
    assert len(rows) >= 32
    kc4, vc4 = kc.cpu().view(-1, bs, G, D), vc.cpu().view(-1, bs, G, D)
    want = A.decode_attention(q[rows].cpu(), kc4, vc4, ctx[rows].cpu(), bt[rows].cpu(), G, D,
                              mode="exact", scale=D ** -0.5)
    err = (o1[rows].cpu().float() - want).abs()
    print(f"config-2 decode shape, {len(rows)} rows vs oracle(exact): max {err.max():.3e} rms {err.pow(2).mean().sqrt():.3e}")
    assert err.max() < 1e-2 and err.pow(2).mean().sqrt() < 1e-3
    # ragged lengths at the same batch size (what a running batch looks like after a few hundred steps)
    ctx_r = torch.randint(1, n + 1, (B,), generator=torch.Generator().manual_seed(5)).to(torch.int32)
    o4 = ops.paged_decode_attn(q, kc, vc, bt, ctx_r.cuda(), D ** -0.5, G)
    want = A.decode_attention(q[rows].cpu(), kc4, vc4, ctx_r[rows], bt[rows].cpu(), G, D,
                              mode="exact", scale=D ** -0.5)
    err = (o4[rows].cpu().float() - want).abs()
    assert err.max() < 1e-2 and err.pow(2).mean().sqrt() < 1e-3


def test_decode_rejects_bad_arguments():
    from nanovllm import _C
    ops = _ops()
    q, kc, vc, ctx, bt = _decode_case(2, 24, 2, 128, 16, [5, 9], seed=1)   # group of 12 > 8
    with pytest.raises(_C.B200Error):
        _run_decode(q, kc, vc, ctx, bt, 2, 128)
    with pytest.raises(RuntimeError):
        ops.rmsnorm(torch.zeros(2, 64, dtype=BF), torch.ones(64, dtype=BF), 1e-6)  # CPU tensor


# ------------------------------------------------------------------ varlen prefill attention
@pytest.fixture(params=[0, 1, 2], ids=["layout_auto", "one_tile_per_cta", "two_tiles_per_cta"])
def prefill_layout(request):
    """Both CTA layouts of the tcgen05 prefill kernel (b200_prefill_set_query_tiles) and the per-launch
    choice between them must pass every prefill parity test."""
    from nanovllm import _C
    lib = _C.load()
    _C.check(lib.b200_prefill_set_query_tiles(request.param), "b200_prefill_set_query_tiles")
    yield request.param
    _C.check(lib.b200_prefill_set_query_tiles(0), "b200_prefill_set_query_tiles")


def test_prefill_golden(golden_kernels, prefill_layout):
    ops = _ops()
    for tag in ("d64", "d128"):
        g = golden_kernels[f"prefill_{tag}"]
        G, D = g["num_kv_heads"], g["head_dim"]
        cu = g["cu_seqlens"].cuda()
        mx = int((g["cu_seqlens"][1:] - g["cu_seqlens"][:-1]).max())
        o = ops.prefill_varlen_attn(g["q"].cuda(), g["k"].cuda(), g["v"].cuda(), cu, cu, mx, mx,
                                    D ** -0.5, G).cpu()
        exact = A.prefill_attention(g["q"], g["k"], g["v"], g["cu_seqlens"], g["cu_seqlens"], G, D,
                                    mode="exact", scale=D ** -0.5)
        assert (o.float() - exact).abs().max() < 2e-2
        assert (o.float() - g["out"].float()).abs().max() < 6e-2


@pytest.mark.parametrize("H,G,D", [(16, 8, 128), (32, 8, 64), (8, 1, 128), (4, 4, 64)])
def test_prefill_ragged_packed(H, G, D, prefill_layout):
    ops = _ops()
    lens = [1, 63, 64, 65, 200, 1024, 7]
    g = torch.Generator().manual_seed(H + D)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
    T = int(cu[-1])
    # q/k/v as strided slices of one fused activation, like the model passes them
    qkv = torch.randn(T, (H + 2 * G) * D, generator=g).to(BF)
    qkv[:, H * D:(H + G) * D] *= 0.5
    q = qkv[:, :H * D].unflatten(1, (H, D))
    k = qkv[:, H * D:(H + G) * D].unflatten(1, (G, D))
    v = qkv[:, (H + G) * D:].unflatten(1, (G, D))
    want = A.prefill_attention(q, k, v, cu, cu, G, D, mode="exact", scale=D ** -0.5)
    d = qkv.cuda()
    o = ops.prefill_varlen_attn(d[:, :H * D].unflatten(1, (H, D)),
                                d[:, H * D:(H + G) * D].unflatten(1, (G, D)),
                                d[:, (H + G) * D:].unflatten(1, (G, D)), cu.cuda(), cu.cuda(),
                                max(lens), max(lens), D ** -0.5, G).cpu()
    assert (o.float() - want).abs().max() < 2e-2


@pytest.mark.parametrize("bs", [16, 256])
def test_prefill_paged_prefix_hit(bs, prefill_layout):
    """Prefix-cache hit form (attention_ori.py:65-70): q shorter than k, K/V read from the
    paged cache through block_tables, bottom-right causal alignment."""
    ops = _ops()
    H, G, D = 16, 8, 128
    klens, qlens = [300, 1024, 40], [44, 256, 40]
    q_, kc, vc, _, bt = _decode_case(3, H, G, D, bs, klens, seed=bs)
    g = torch.Generator().manual_seed(9)
    cu_q = torch.tensor([0] + list(torch.tensor(qlens).cumsum(0)), dtype=torch.int32)
    q = torch.randn(int(cu_q[-1]), H, D, generator=g).to(BF)
    kp, vp, cu_k = A.gather_paged_kv(kc, vc, bt, klens)
    want = A.prefill_attention(q, kp, vp, cu_q, cu_k, G, D, mode="exact", scale=D ** -0.5,
                               causal_align="bottom_right")
    nb = kc.shape[0]
    o = ops.prefill_varlen_attn(q.cuda(), kc.cuda().view(nb, bs, G * D), vc.cuda().view(nb, bs, G * D),
                                cu_q.cuda(), cu_k.cuda(), max(qlens), max(klens), D ** -0.5, G,
                                block_tables=bt.cuda()).cpu()
    assert (o.float() - want).abs().max() < 2e-2


@pytest.mark.parametrize("H,G,D,bs,qlens,klens", [
    (32, 8, 64, 16, [256], [8192]),                     # BASELINE config 3: Llama-3.2-1B, 7936-token prefix hit
    (16, 8, 128, 16, [200, 33], [4096, 150]),           # one sequence with fewer key tiles than splits
    (8, 1, 128, 256, [256, 256], [2048, 1999]),         # 32B TP=8 shard shape, reference page size
])
def test_prefill_split_kv(H, G, D, bs, qlens, klens, prefill_layout):
    """Few query rows against many cached keys: the launch splits every (sequence, head, 256 rows)
    along the keys over several CTAs (un-normalised partials in the workspace + merge kernel). Against
    the oracle's fp32 evaluation, and against the same launch with the split disabled."""
    ops = _ops()
    n = len(qlens)
    _, kc, vc, _, bt = _decode_case(n, H, G, D, bs, klens, seed=H + D)
    g = torch.Generator().manual_seed(4)
    cu_q = torch.tensor([0] + list(torch.tensor(qlens).cumsum(0)), dtype=torch.int32)
    q = torch.randn(int(cu_q[-1]), H, D, generator=g).to(BF)
    kp, vp, cu_k = A.gather_paged_kv(kc, vc, bt, klens)
    want = A.prefill_attention(q, kp, vp, cu_q, cu_k, G, D, mode="exact", scale=D ** -0.5,
                               causal_align="bottom_right")
    nb = kc.shape[0]
    args = (q.cuda(), kc.cuda().view(nb, bs, G * D), vc.cuda().view(nb, bs, G * D), cu_q.cuda(), cu_k.cuda(),
            max(qlens), max(klens), D ** -0.5, G)
    before = ops.launch_count()
    split = ops.prefill_varlen_attn(*args, block_tables=bt.cuda()).cpu()
    assert ops.launch_count() - before == 2, "expected the split launch (attention + merge kernel)"
    whole = ops.prefill_varlen_attn(*args, block_tables=bt.cuda(), workspace=None).cpu()
    assert (split.float() - want).abs().max() < 2e-2
    assert (whole.float() - want).abs().max() < 2e-2
    assert (split.float() - whole.float()).abs().max() < 1.6e-2     # both round the same fp32 value to bf16
    # packed K/V take the same path
    packed = ops.prefill_varlen_attn(q.cuda(), kp.cuda(), vp.cuda(), cu_q.cuda(), cu_k.cuda(), max(qlens),
                                     max(klens), D ** -0.5, G).cpu()
    assert (packed.float() - want).abs().max() < 2e-2


# ------------------------------------------------------------------ sampling tail
def test_sample_argmax_and_gumbel():
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(64, 151936, generator=g).to(BF)
    got = ops.sample(logits.cuda()).cpu()
    assert torch.equal(got, L.greedy(logits))
    # temperature sampling: the empirical distribution follows softmax(logits / T)
    small = torch.tensor([[2.0, 1.0, 0.0, -1.0]] * 4096).to(BF)
    t = torch.full((4096,), 0.7)
    draws = ops.sample(small.cuda(), t.cuda(), seed=11, step=5).cpu()
    freq = torch.bincount(draws, minlength=4).float() / 4096
    want = torch.softmax(small[0].float() / 0.7, -1)
    assert (freq - want).abs().max() < 0.03
    assert not torch.equal(draws, ops.sample(small.cuda(), t.cuda(), seed=11, step=6).cpu())


def test_sharded_sampling_equals_full_vocabulary_sampling():
    """b200_sample_shard on 1..8 vocabulary shards + a MAX over the keys == b200_sample on the whole
    row, bit for bit: greedy rows, temperature rows (the Gumbel draw is keyed by the GLOBAL index),
    ties (lowest index wins), a shard that is entirely -inf. Single GPU: the all-reduce is a torch max."""
    from nanovllm import ops
    g = torch.Generator().manual_seed(3)
    rows, vocab = 37, 151936
    logits = (3 * torch.randn(rows, vocab, generator=g)).to(torch.bfloat16).cuda()
    logits[5, 1000:1010] = 30.0                      # a 10-way tie for the maximum inside one shard
    logits[6, 10] = logits[6, vocab - 7] = 31.0      # a tie across the first and the last shard
    logits[7, : vocab // 2] = float("-inf")          # an all -inf shard (world = 2)
    temps = torch.rand(rows, generator=g) * 1.5
    temps[::3] = 0.0                                 # every third row greedy
    temps = temps.cuda()
    off = torch.tensor([11], dtype=torch.int64, device="cuda")
    for world in (1, 2, 4, 8):
        per = vocab // world
        for t in (None, temps):
            want = ops.sample(logits, t, seed=5, step=9, step_offset=off)
            keys = None
            for r in range(world):
                shard = logits[:, r * per:(r + 1) * per].contiguous()
                k = ops.sample_shard(shard, t, seed=5, step=9, index_offset=r * per, step_offset=off)
                keys = k if keys is None else torch.maximum(keys, k)
            got = ops.decode_sample_keys(keys)
            assert torch.equal(got, want), (world, t is None, (got != want).nonzero().flatten().tolist())
    assert int(want[5]) >= 0 and int(ops.sample(logits, None)[5]) == 1000 and int(ops.sample(logits, None)[6]) == 10
