"""Oracle (TEST INFRASTRUCTURE): the sparse mixture-of-experts MLP of the reference's Qwen3-MoE
stack on CPU -- ``models/qwen3_moe.py`` ``Qwen3MoeSparseMoeBlock`` (:125-186) and the expert MLP
``Qwen3MoeMLP`` (:95-122).

Pinned against the reference: ``oracle/gen_golden.py::gen_moe`` runs the unmodified reference block
and a tiny Qwen3-MoE model on seeded inputs (tests/golden/moe.pt); ``tests/test_oracle_golden.py``
requires these functions to reproduce them bit for bit.

The same module also states, in numpy-level Python, WHAT the device-side dispatch plan must contain
(``dispatch_plan``): the product computes it with ``b200_moe_align``; the GPU tests compare the two.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import layers_ref as L


def route(hidden, gate_w, top_k):
    """qwen3_moe.py:150-158: router logits in the model dtype, softmax in fp32 over ALL experts,
    top-k, renormalise the k probabilities, round to the model dtype.
    Returns (weights [T, k] dtype, expert ids [T, k] int64, router logits [T, E] dtype)."""
    logits = F.linear(hidden, gate_w)
    w = F.softmax(logits, dim=1, dtype=torch.float)
    w, ids = torch.topk(w, top_k, dim=-1)
    w = w / w.sum(dim=-1, keepdim=True)
    return w.to(hidden.dtype), ids, logits


def expert_mlp(x, w13, w2):
    """qwen3_moe.py:117-122 (gate_up GEMM, SiluAndMul, down GEMM), one expert's weights."""
    return F.linear(L.silu_and_mul(F.linear(x, w13)), w2)


def sparse_moe_block(hidden, gate_w, w13, w2, top_k, experts=None):
    """qwen3_moe.py:146-186. w13 [E, 2I, h], w2 [E, h, I] (or lists of per-expert matrices).
    Experts are visited in ascending index; each adds ``expert(x) * weight`` (rounded to the model
    dtype) into a model-dtype accumulator with index_add_ (one rounding per expert hit).
    ``experts`` restricts the sum to a subset (an expert-parallel rank's partial result)."""
    T, h = hidden.shape
    rw, ids, _ = route(hidden, gate_w, top_k)
    E = len(w13)
    out = torch.zeros_like(hidden)
    mask = F.one_hot(ids, num_classes=E).permute(2, 1, 0)          # [E, k, T]
    hit = torch.greater(mask.sum(dim=(-1, -2)), 0).nonzero()
    for e in hit.flatten().tolist():
        if experts is not None and e not in experts:
            continue
        slot, tok = torch.where(mask[e])
        cur = hidden[None, tok].reshape(-1, h)
        cur = expert_mlp(cur, w13[e], w2[e]) * rw[tok, slot, None]
        out.index_add_(0, tok, cur.to(hidden.dtype))
    return out


def dispatch_plan(ids, expert_begin, num_local):
    """What b200_moe_align must produce for expert ids [T, k] and the local expert range
    [expert_begin, expert_begin + num_local): pairs (t, j) -> pair index t*k + j, stably sorted by
    local expert. Returns (offsets [num_local + 1], sorted_pair [n_local_pairs], pos_of_pair [T*k]
    with -1 for pairs routed to experts of other ranks)."""
    flat = ids.reshape(-1).to(torch.int64) - expert_begin
    local = (flat >= 0) & (flat < num_local)
    key = torch.where(local, flat, torch.full_like(flat, num_local))
    order = torch.sort(key, stable=True).indices
    n_local = int(local.sum())
    sorted_pair = order[:n_local].to(torch.int32)
    counts = torch.bincount(flat[local], minlength=num_local)
    offsets = torch.zeros(num_local + 1, dtype=torch.int32)
    offsets[1:] = counts.cumsum(0).to(torch.int32)
    pos = torch.full((flat.numel(),), -1, dtype=torch.int32)
    pos[sorted_pair.long()] = torch.arange(n_local, dtype=torch.int32)
    return offsets, sorted_pair, pos
