"""Weights: safetensors checkpoints through per-parameter ``weight_loader`` hooks (reference:
utils/loader.py:12-59) and seeded random initialisation for synthetic benchmarks / parity."""
from __future__ import annotations

import os
from glob import glob

import torch
from torch import nn


def load_model(model: nn.Module, path: str) -> int:
    """Fill `model` from every *.safetensors under `path`; returns the number of tensors read.
    Checkpoint names containing a packed-module key (q_proj, gate_proj, ...) are routed to the
    fused parameter with their shard id, everything else by exact name."""
    from safetensors import safe_open
    packed = getattr(model, "packed_modules_mapping", {})
    elsewhere = getattr(model, "virtual_param_shapes", dict)()    # experts held by other ranks
    count = 0
    for file in sorted(glob(os.path.join(path, "*.safetensors"))):
        with safe_open(file, "pt", "cpu") as f:
            for name in f.keys():
                fused_name = name
                for key, (fused, _) in packed.items():
                    if key in name:
                        fused_name = name.replace(key, fused)
                        break
                if fused_name in elsewhere:
                    continue
                tensor = f.get_tensor(name)
                for key, (fused, shard_id) in packed.items():
                    if key in name:
                        param = model.get_parameter(name.replace(key, fused))
                        param.weight_loader(param, tensor.to(param.dtype), shard_id)
                        break
                else:
                    param = model.get_parameter(name)
                    loader = getattr(param, "weight_loader", lambda p, w: p.data.copy_(w))
                    loader(param, tensor.to(param.dtype))
                count += 1
    return count


def full_shapes(model: nn.Module, tp_size: int) -> dict[str, tuple[int, ...]]:
    """Unsharded shape of every parameter (what a checkpoint would hold)."""
    shapes = {}
    for name, p in model.named_parameters():
        shape = list(p.shape)
        mod = model.get_submodule(name.rsplit(".", 1)[0])
        tp_dim = getattr(mod, "tp_dim", 0 if hasattr(mod, "vocab_start_idx") else None)
        if tp_size > 1 and tp_dim is not None and len(shape) > tp_dim and not (
                len(shape) == 1 and tp_dim == 1):
            if getattr(mod, "kv_replicas", 1) > 1:     # q sharded, kv heads replicated (linear.py)
                shape[tp_dim] = (mod.num_heads * tp_size + 2 * mod.total_num_kv_heads) * mod.head_size
            else:
                shape[tp_dim] *= tp_size
        shapes[name] = tuple(shape)
    return shapes


def device_init(model: nn.Module, seed: int = 0, std: float = 0.02) -> None:
    """Random init generated ON the device (for models too large to draw on the host in reasonable
    time, e.g. Qwen3-32B = 65 GB): same distribution as seeded_init, not the same numbers; each
    rank of a tensor-parallel group seeds by (seed, rank), which is fine for throughput runs."""
    from ..layers.linear import tp_rank_size
    rank, _ = tp_rank_size()
    tied = getattr(getattr(model, "spec", None), "tie_embeddings", False)
    for idx, (name, p) in enumerate(sorted(model.named_parameters())):
        if tied and name == "lm_head.weight":
            continue
        gen = torch.Generator(device=p.device).manual_seed(seed * 1000003 + idx * 17 + rank)
        if p.dim() == 1:
            p.data.fill_(1.0)
        else:
            p.data.normal_(0.0, std, generator=gen)
    if tied:
        model.lm_head.weight.data = model.model.embed_tokens.weight.data


def seeded_init(model: nn.Module, seed: int = 0, norm_jitter: float = 0.0, std: float = 0.02) -> None:
    """Deterministic random init keyed by parameter NAME (identical to oracle/model_ref.py's
    ``seeded_weights``): the i-th name in sorted order draws from Generator(seed*1000003 + i);
    matrices ~ N(0, std), 1-D (norm) weights = 1 + jitter*N(0,1). With tensor parallelism the
    full tensor is generated and this rank's shard is cut by the parameter's weight_loader, so
    every TP layout computes the same function."""
    from ..layers.linear import tp_rank_size
    _, tp = tp_rank_size()
    tied = getattr(getattr(model, "spec", None), "tie_embeddings", False)
    # numbered over the FULL parameter list (experts of other ranks included) so that every
    # parallel layout draws the same weights as the single-GPU model and the oracle
    elsewhere = getattr(model, "virtual_param_shapes", dict)()
    names = sorted({n for n, _ in model.named_parameters() if not (tied and n == "lm_head.weight")}
                   | set(elsewhere))
    shapes = full_shapes(model, tp)
    params = dict(model.named_parameters())
    for idx, name in enumerate(names):
        if name in elsewhere:
            continue
        gen = torch.Generator().manual_seed(seed * 1000003 + idx)
        shape = shapes[name]
        if len(shape) == 1:
            full = torch.ones(shape) + norm_jitter * torch.randn(shape, generator=gen)
        else:
            full = std * torch.randn(shape, generator=gen)
        p = params[name]
        full = full.to(p.dtype)
        loader = getattr(p, "weight_loader", None)
        if loader is None or tuple(p.shape) == shape:
            p.data.copy_(full)
        else:
            _load_fused(model, name, p, full)
    if tied:
        model.lm_head.weight.data = model.model.embed_tokens.weight.data


def _load_fused(model, name, p, full):
    """Route a full (unsharded) fused tensor through the shard-aware weight_loader."""
    mod = model.get_submodule(name.rsplit(".", 1)[0])
    tp = mod.tp_size
    if hasattr(mod, "output_sizes"):                      # gate|up
        off = 0
        for i, n in enumerate(mod.output_sizes):
            p.weight_loader(p, full.narrow(0, off, n), i)
            off += n
    elif hasattr(mod, "num_kv_heads") and hasattr(mod, "head_size"):   # q|k|v
        q = mod.num_heads * tp * mod.head_size
        kv = getattr(mod, "total_num_kv_heads", mod.num_kv_heads * tp) * mod.head_size
        p.weight_loader(p, full.narrow(0, 0, q), "q")
        p.weight_loader(p, full.narrow(0, q, kv), "k")
        p.weight_loader(p, full.narrow(0, q + kv, kv), "v")
    else:
        p.weight_loader(p, full)
