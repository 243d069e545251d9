"""Oracle: KV scatter, paged decode attention, varlen prefill attention.

Two arithmetic modes per attention function:
  * ``mode="reference"``: the reference's torch-native arithmetic, step for step --
    every matmul in the model dtype, scale = dtype(1/sqrt(D)), scores clamped to +-100,
    softmax in the model dtype (attention_torch_native.py:80,127,139-140,188-192).
  * ``mode="exact"``: the same mathematical function evaluated in fp32 (what a fused kernel
    with fp32 accumulation approximates); used to state kernel tolerances.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def store_kvcache(key, value, k_cache, v_cache, slot_mapping, block_size):
    """Scatter rows of key/value [N, G, D] into caches [nb, bs, G, D] at flat slots.

    Follows attention_torch_native.py:13-51 (store_kvcache): slot -1 means "no slot";
    slots that fall outside the cache are dropped (:33-35). Negative slots other than -1
    do not occur in the reference runner (model_runner.py:262-270) and are dropped too.
    """
    nb, bs = k_cache.shape[0], k_cache.shape[1]
    slots = slot_mapping.to(torch.int64)
    keep = (slots >= 0) & (slots < nb * block_size) & ((slots % block_size) < bs)
    rows = torch.nonzero(keep).flatten()
    blk = torch.div(slots[rows], block_size, rounding_mode="floor")
    off = slots[rows] - blk * block_size
    k_cache[blk, off] = key[rows].to(k_cache.dtype)
    v_cache[blk, off] = value[rows].to(v_cache.dtype)


def _expand_kv(x, n_rep):
    """[L, G, D] -> [L, G*n_rep, D], each kv head repeated n_rep times consecutively
    (attention_torch_native.py:53-58, without its H<=32 guard)."""
    if n_rep == 1:
        return x
    L, G, D = x.shape
    return x[:, :, None, :].expand(L, G, n_rep, D).reshape(L, G * n_rep, D)


def _ref_scale(head_dim, dtype):
    # attention_torch_native.py:80 -- 1/sqrt(D) computed in fp32 then rounded to bf16
    return (1.0 / torch.sqrt(torch.tensor(head_dim, dtype=torch.float32))).to(dtype)


def _attend(q_hld, k_hdl, v_hld, mask, mode, scale):
    """q [H, Lq, D], k [H, D, Lk], v [H, Lk, D]; mask additive [Lq, Lk] or None."""
    if mode == "reference":
        att = torch.matmul(q_hld, k_hdl) * scale
        if mask is not None:
            att = att + mask.to(att.dtype)[None]
        att = F.softmax(att.clamp(-100, 100), dim=-1)
        return torch.matmul(att, v_hld)
    att = torch.matmul(q_hld.float(), k_hdl.float()) * float(scale)
    if mask is not None:
        att = att + mask.float()[None]
    att = F.softmax(att, dim=-1)
    return torch.matmul(att, v_hld.float())


def decode_attention(q, k_cache, v_cache, context_lens, block_tables, num_kv_heads, head_dim,
                     mode="reference", scale=None):
    """One query token per sequence against its paged history.

    q [B, H, D]; caches [nb, bs, G, D]; block_tables [B, max_blk] (-1 padded);
    context_lens [B]. Returns [B, H*D] (fp32 in exact mode, model dtype otherwise).
    Follows attention_torch_native.py:147-196 (_decode_attn): per sequence, concatenate the
    blocks named by the non-negative table entries, keep the first context_len tokens,
    expand kv heads, softmax(q k^T * scale) v. A row with context_len == 0 yields zeros here
    (the reference would softmax over an empty axis; its runner never consumes such rows,
    model_runner.py:303-311 + sampler.py:11-12).
    """
    B = q.shape[0]
    H = q.shape[1]
    n_rep = H // num_kv_heads
    sc = _ref_scale(head_dim, q.dtype) if scale is None else scale
    odt = q.dtype if mode == "reference" else torch.float32
    out = torch.zeros(B, H, head_dim, dtype=odt)
    nb = k_cache.shape[0]
    for i in range(B):
        n = int(context_lens[i])
        ids = [int(b) for b in block_tables[i].tolist() if 0 <= b < nb]
        if n <= 0 or not ids:
            continue
        idx = torch.tensor(ids, dtype=torch.long)
        k_hist = k_cache[idx].reshape(-1, num_kv_heads, head_dim)[:n]
        v_hist = v_cache[idx].reshape(-1, num_kv_heads, head_dim)[:n]
        k_hist = _expand_kv(k_hist, n_rep)
        v_hist = _expand_kv(v_hist, n_rep)
        o = _attend(q[i][:, None, :], k_hist.permute(1, 2, 0), v_hist.transpose(0, 1), None,
                    mode, sc)
        out[i] = o[:, 0, :]
    return out.reshape(B, H * head_dim)


def prefill_attention(q, k, v, cu_seqlens_q, cu_seqlens_k, num_kv_heads, head_dim,
                      mode="reference", scale=None, causal_align="top_left"):
    """Varlen causal self-attention over packed sequences.

    q [Tq, H, D], k/v [Tk, G, D]. Follows attention_torch_native.py:103-145 (_prefill_attn):
    per sequence slices, kv-head expansion, additive -inf mask above the main diagonal
    (``torch.triu(..., diagonal=1)``, i.e. TOP-LEFT aligned when Lq != Lk), clamp, softmax.
    ``causal_align="bottom_right"`` gives the flash-attn / NPU sparse_mode=3 alignment
    (query r sees keys <= r + Lk - Lq); both coincide when Lq == Lk, the only case the
    reference runner produces (model_runner.py:252-256).
    """
    H = q.shape[1]
    n_rep = H // num_kv_heads
    sc = _ref_scale(head_dim, q.dtype) if scale is None else scale
    outs = []
    for i in range(len(cu_seqlens_q) - 1):
        qs, qe = int(cu_seqlens_q[i]), int(cu_seqlens_q[i + 1])
        ks, ke = int(cu_seqlens_k[i]), int(cu_seqlens_k[i + 1])
        Lq, Lk = qe - qs, ke - ks
        kk = _expand_kv(k[ks:ke], n_rep)
        vv = _expand_kv(v[ks:ke], n_rep)
        shift = 1 if causal_align == "top_left" else 1 + (Lk - Lq)
        mask = torch.triu(torch.full((Lq, Lk), float("-inf")), diagonal=shift)
        o = _attend(q[qs:qe].transpose(0, 1), kk.permute(1, 2, 0), vv.transpose(0, 1), mask,
                    mode, sc)
        if mode != "reference":
            o = torch.nan_to_num(o, nan=0.0)  # rows with no visible key (Lk < Lq) -> 0
        outs.append(o.transpose(0, 1))
    odt = q.dtype if mode == "reference" else torch.float32
    if not outs:
        return torch.zeros(0, H * head_dim, dtype=odt)
    return torch.cat(outs, 0).reshape(-1, H * head_dim)


def gather_paged_kv(k_cache, v_cache, block_tables, seq_lens):
    """Helper for prefix-hit prefill tests: materialise each sequence's first seq_len cached
    tokens as packed [Tk, G, D] tensors + cu_seqlens_k (what flash_attn_varlen_func reads
    through block_table, attention_ori.py:65-70)."""
    ks, vs, cu = [], [], [0]
    for i, n in enumerate(seq_lens):
        ids = torch.tensor([int(b) for b in block_tables[i].tolist() if b >= 0], dtype=torch.long)
        kk = k_cache[ids].reshape(-1, *k_cache.shape[2:])[:n]
        vv = v_cache[ids].reshape(-1, *v_cache.shape[2:])[:n]
        ks.append(kk), vs.append(vv), cu.append(cu[-1] + n)
    return torch.cat(ks), torch.cat(vs), torch.tensor(cu, dtype=torch.int32)


# ---------------------------------------------------------------------------------------------
# FP8 (e4m3) KV cache -- SURVEY.md 8f row f3. NOT a restatement of reference code: the reference
# keeps its caches in the model dtype (engine/model_runner.py:218-229). This states the semantics
# the fp8 kernels (csrc/decode_attn_fp8.cu) implement, with torch's own float8_e4m3fn cast as the
# arithmetic authority, so the scatter can be checked bit for bit and attention against
# ``decode_attention(mode="exact")`` over the de-quantised cache.
E4M3_MAX = 448.0


def quantize_e4m3(x, scale=1.0):
    """bytes of e4m3(x * (1/scale)): the product and the reciprocal are fp32, the cast is round to
    nearest even, saturating at +-448 (torch's cast would turn overflow into NaN, so clamp first)."""
    inv = (torch.tensor(1.0, dtype=torch.float32) / torch.tensor(scale, dtype=torch.float32))
    y = (x.float() * inv).clamp(-E4M3_MAX, E4M3_MAX)
    return y.to(torch.float8_e4m3fn).view(torch.uint8)


def dequantize_e4m3(q, scale=1.0, dtype=torch.float32):
    return (q.view(torch.float8_e4m3fn).float() * float(scale)).to(dtype)


def store_kvcache_fp8(key, value, k_cache, v_cache, slot_mapping, block_size, k_scale=1.0, v_scale=1.0):
    """store_kvcache into uint8 caches [nb, bs, G, D] holding e4m3 bytes (same slot rules)."""
    store_kvcache(quantize_e4m3(key, k_scale), quantize_e4m3(value, v_scale), k_cache, v_cache,
                  slot_mapping, block_size)


def decode_attention_fp8(q, k_cache, v_cache, context_lens, block_tables, num_kv_heads, head_dim,
                         scale, k_scale=1.0, v_scale=1.0):
    """Exact-mode decode attention over the de-quantised e4m3 caches. Returns fp32 [B, H*D]."""
    return decode_attention(q, dequantize_e4m3(k_cache, k_scale), dequantize_e4m3(v_cache, v_scale),
                            context_lens, block_tables, num_kv_heads, head_dim, mode="exact", scale=scale)
