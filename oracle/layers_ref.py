"""Oracle: the eager layers around attention (RMSNorm, RoPE, SiluAndMul, sampler).

Each function restates the reference's eager op sequence so the rounding points match.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def rmsnorm(x, weight, eps):
    """layers/layernorm.py:16-25 (rms_forward): fp32 statistics, cast back, multiply by the
    weight in the model dtype."""
    xf = x.float()
    xf = xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)
    return xf.to(x.dtype) * weight


def add_rmsnorm(x, residual, weight, eps):
    """layers/layernorm.py:27-38 (add_rms_forward): returns (normed, new_residual); the new
    residual is the fp32 sum rounded to the model dtype, the norm uses the unrounded sum."""
    s = x.float() + residual.float()
    new_res = s.to(x.dtype)
    s = s * torch.rsqrt(s.pow(2).mean(dim=-1, keepdim=True) + eps)
    return s.to(x.dtype) * weight, new_res


def rope_cache(head_dim, max_position, base):
    """layers/rotary_embedding.py:26-35: [max_position, 1, D] fp32, cos half then sin half."""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2, dtype=torch.float) / head_dim))
    ang = torch.outer(torch.arange(max_position, dtype=torch.float), inv_freq)
    return torch.cat((ang.cos(), ang.sin()), dim=-1)[:, None, :]


def apply_rope(positions, x, cos_sin_cache):
    """layers/rotary_embedding.py:6-14,43-46 (NeoX half-split rotation, fp32, cast back).
    x [T, heads, D]."""
    cs = cos_sin_cache[positions]
    half = cs.shape[-1] // 2
    cos, sin = cs[..., :half], cs[..., half:]
    xf = x.float()
    a, b = xf[..., :half], xf[..., half:]
    return torch.cat((a * cos - b * sin, b * cos + a * sin), dim=-1).to(x.dtype)


def silu_and_mul(x):
    """layers/activation.py:11-13."""
    half = x.shape[-1] // 2
    return F.silu(x[..., :half]) * x[..., half:]


def qkv_epilogue(qkv, positions, cos_sin_cache, q_norm_w, k_norm_w, eps, num_heads,
                 num_kv_heads, head_dim):
    """models/qwen3.py:79-87: split, optional per-head q/k RMSNorm, RoPE. Returns q,k,v 3-D."""
    qs, kvs = num_heads * head_dim, num_kv_heads * head_dim
    q, k, v = qkv.split([qs, kvs, kvs], dim=-1)
    q = q.reshape(-1, num_heads, head_dim)
    k = k.reshape(-1, num_kv_heads, head_dim)
    v = v.reshape(-1, num_kv_heads, head_dim)
    if q_norm_w is not None:
        q = rmsnorm(q, q_norm_w, eps)
        k = rmsnorm(k, k_norm_w, eps)
    q = apply_rope(positions, q, cos_sin_cache)
    k = apply_rope(positions, k, cos_sin_cache)
    return q, k, v


def greedy(logits):
    """Parity sampler: argmax of the logits (the reference forbids temperature 0,
    sampling_params.py:11, and samples with torch.multinomial, sampler.py:9-17; greedy parity
    runs take the argmax of the same logits tensor instead)."""
    return logits.float().argmax(dim=-1)
