"""Oracle: the decoder stack and the runner glue around the attention path, on CPU.

Restates, as pure functions over a flat ``{name: tensor}`` weight dict, what the reference
computes with ``models/qwen3.py`` / ``models/llama.py`` built from its eager layers and the
torch-native attention, driven by the metadata its NPU-only runner would build
(``engine/model_runner.py:238-366``). Parameter names are the reference's own
(``model.layers.N.self_attn.qkv_proj.weight`` ...), so one seeded initialiser fills the
reference model, this oracle and the CUDA product identically.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

from . import attention_ref as A
from . import layers_ref as L


@dataclass
class ModelCfg:
    vocab_size: int = 512
    hidden_size: int = 128
    num_layers: int = 2
    num_heads: int = 4
    num_kv_heads: int = 2
    head_dim: int = 64
    intermediate_size: int = 256
    rms_eps: float = 1e-6
    rope_theta: float = 1000000.0
    max_position: int = 512
    qk_norm: bool = True          # Qwen3 (no qkv bias) -> per-head q/k RMSNorm (qwen3.py:70-72)
    tie_embeddings: bool = True
    dtype: torch.dtype = torch.bfloat16
    # Qwen3-MoE (models/qwen3_moe.py): num_experts > 0 makes EVERY layer's MLP a sparse MoE block
    # (decoder_sparse_step = 1, mlp_only_layers = [] as in Qwen3-30B-A3B); intermediate_size is then unused
    num_experts: int = 0
    top_k: int = 0
    moe_intermediate_size: int = 0

    @staticmethod
    def tiny_qwen3_moe():
        return ModelCfg(vocab_size=512, hidden_size=128, num_layers=2, num_heads=4, num_kv_heads=2,
                        head_dim=64, intermediate_size=256, num_experts=8, top_k=2,
                        moe_intermediate_size=64, tie_embeddings=False)

    @staticmethod
    def qwen3_0_6b():
        return ModelCfg(vocab_size=151936, hidden_size=1024, num_layers=28, num_heads=16,
                        num_kv_heads=8, head_dim=128, intermediate_size=3072,
                        max_position=40960)

    @staticmethod
    def llama_3_2_1b():
        # rope_theta: the model card's 500000. (Under transformers>=5 the reference's
        # getattr(config, "rope_theta", 10000) at llama.py:133 falls back to 10000; the base is
        # an input of this oracle, parity tests pass the same value to both sides.)
        return ModelCfg(vocab_size=128256, hidden_size=2048, num_layers=16, num_heads=32,
                        num_kv_heads=8, head_dim=64, intermediate_size=8192, rms_eps=1e-5,
                        rope_theta=500000.0, max_position=131072, qk_norm=False)


def param_shapes(cfg: ModelCfg) -> dict[str, tuple[int, ...]]:
    """Names and shapes of every parameter of the reference Qwen3/Llama graph at TP=1."""
    h, d = cfg.hidden_size, cfg.head_dim
    shapes = {"model.embed_tokens.weight": (cfg.vocab_size, h), "model.norm.weight": (h,)}
    if not cfg.tie_embeddings:
        shapes["lm_head.weight"] = (cfg.vocab_size, h)
    for i in range(cfg.num_layers):
        p = f"model.layers.{i}."
        shapes[p + "input_layernorm.weight"] = (h,)
        shapes[p + "post_attention_layernorm.weight"] = (h,)
        shapes[p + "self_attn.qkv_proj.weight"] = ((cfg.num_heads + 2 * cfg.num_kv_heads) * d, h)
        shapes[p + "self_attn.o_proj.weight"] = (h, cfg.num_heads * d)
        if cfg.qk_norm:
            shapes[p + "self_attn.q_norm.weight"] = (d,)
            shapes[p + "self_attn.k_norm.weight"] = (d,)
        if cfg.num_experts:
            shapes[p + "mlp.gate.weight"] = (cfg.num_experts, h)
            for e in range(cfg.num_experts):
                shapes[p + f"mlp.experts.{e}.gate_up_proj.weight"] = (2 * cfg.moe_intermediate_size, h)
                shapes[p + f"mlp.experts.{e}.down_proj.weight"] = (h, cfg.moe_intermediate_size)
            continue
        shapes[p + "mlp.gate_up_proj.weight"] = (2 * cfg.intermediate_size, h)
        shapes[p + "mlp.down_proj.weight"] = (h, cfg.intermediate_size)
    return shapes


def seeded_weights(cfg: ModelCfg, seed: int = 0, norm_jitter: float = 0.0,
                   std: float = 0.02) -> dict[str, torch.Tensor]:
    """Deterministic random-init weights keyed by parameter name: matrices ~ N(0, std),
    norm weights 1 (+ jitter * N(0,1) when asked, so the weight multiply is exercised).
    Each tensor draws from its own generator seeded by (seed, sorted position) so the result
    does not depend on which consumer asks first."""
    out = {}
    for idx, (name, shape) in enumerate(sorted(param_shapes(cfg).items())):
        gen = torch.Generator().manual_seed(seed * 1000003 + idx)
        if len(shape) == 1:
            w = torch.ones(shape) + norm_jitter * torch.randn(shape, generator=gen)
        else:
            w = std * torch.randn(shape, generator=gen)
        out[name] = w.to(cfg.dtype)
    return out


@dataclass
class StepMeta:
    """What the reference publishes through its global Context (utils/context.py:5-17)."""
    is_prefill: bool
    slot_mapping: torch.Tensor                 # flat int32 [T]
    cu_seqlens_q: torch.Tensor | None = None
    cu_seqlens_k: torch.Tensor | None = None
    context_lens: torch.Tensor | None = None
    block_tables: torch.Tensor | None = None


def forward(w, cfg: ModelCfg, input_ids, positions, meta: StepMeta, kv_cache, block_size,
            attn_mode="reference", return_hidden=False, trace: dict | None = None,
            fp32_logits=False):
    """One model step. kv_cache: [2, L, nb, bs, Hkv, D] (written in place).
    Mirrors Qwen3Model.forward (qwen3.py:175-185) / Qwen3DecoderLayer.forward (:148-161) /
    Qwen3Attention.forward (:74-90) / Qwen3MLP.forward (:115-119) / compute_logits (:216-218)."""
    H, G, D = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
    cos_sin = _rope_cache(cfg)
    x = F.embedding(input_ids, w["model.embed_tokens.weight"])
    residual = None
    for i in range(cfg.num_layers):
        p = f"model.layers.{i}."
        if residual is None:
            residual, x = x, L.rmsnorm(x, w[p + "input_layernorm.weight"], cfg.rms_eps)
        else:
            x, residual = L.add_rmsnorm(x, residual, w[p + "input_layernorm.weight"], cfg.rms_eps)
        if trace is not None:
            trace[p + "attn_in"] = x
        qkv = F.linear(x, w[p + "self_attn.qkv_proj.weight"])
        if trace is not None:
            trace[p + "qkv"] = qkv
        q, k, v = L.qkv_epilogue(qkv, positions, cos_sin,
                                 w.get(p + "self_attn.q_norm.weight"),
                                 w.get(p + "self_attn.k_norm.weight"), cfg.rms_eps, H, G, D)
        kc, vc = kv_cache[0, i], kv_cache[1, i]
        A.store_kvcache(k, v, kc, vc, meta.slot_mapping, block_size)
        if meta.is_prefill:
            o = A.prefill_attention(q, k, v, meta.cu_seqlens_q, meta.cu_seqlens_k, G, D,
                                    mode=attn_mode)
        else:
            o = A.decode_attention(q, kc, vc, meta.context_lens, meta.block_tables, G, D,
                                   mode=attn_mode)
        o = o.to(cfg.dtype)
        if trace is not None:
            trace[p + "q"], trace[p + "k"], trace[p + "attn_out"] = q, k, o
        x = F.linear(o, w[p + "self_attn.o_proj.weight"])
        if trace is not None:
            trace[p + "o_proj"] = x
        x, residual = L.add_rmsnorm(x, residual, w[p + "post_attention_layernorm.weight"],
                                    cfg.rms_eps)
        if trace is not None:
            trace[p + "mlp_in"] = x
        if cfg.num_experts:      # Qwen3MoeSparseMoeBlock (qwen3_moe.py:146-186)
            from . import moe_ref
            x = moe_ref.sparse_moe_block(
                x, w[p + "mlp.gate.weight"],
                [w[p + f"mlp.experts.{e}.gate_up_proj.weight"] for e in range(cfg.num_experts)],
                [w[p + f"mlp.experts.{e}.down_proj.weight"] for e in range(cfg.num_experts)],
                cfg.top_k)
        else:
            x = F.linear(x, w[p + "mlp.gate_up_proj.weight"])
            x = F.linear(L.silu_and_mul(x), w[p + "mlp.down_proj.weight"])
        if trace is not None:
            trace[p + "mlp_out"] = x
    x, _ = L.add_rmsnorm(x, residual, w["model.norm.weight"], cfg.rms_eps)
    if return_hidden:
        return x
    if meta.is_prefill:  # embed_head.py:57-60: logits only for each sequence's last token
        x = x[(meta.cu_seqlens_q[1:] - 1).long()].contiguous()
    head = w["model.embed_tokens.weight"] if cfg.tie_embeddings else w["lm_head.weight"]
    if fp32_logits:      # the LM-head product before its cast to the model dtype (parity yardstick)
        return F.linear(x.float(), head.float())
    return F.linear(x, head)


_ROPE = {}


def _rope_cache(cfg):
    key = (cfg.head_dim, cfg.max_position, cfg.rope_theta)
    if key not in _ROPE:
        _ROPE[key] = L.rope_cache(cfg.head_dim, cfg.max_position, cfg.rope_theta)
    return _ROPE[key]


# ---------------------------------------------------------------- runner glue (metadata)
def prefill_meta(seq_lens, block_tables_py, block_size):
    """model_runner.py:238-290 (prepare_prefill): every prompt token is fed (the reference
    recomputes cached prefixes too, :247-261), positions 0..len-1, flat slots block*bs+off."""
    ids_pos, cu, slots = [], [0], []
    for n, table in zip(seq_lens, block_tables_py):
        ids_pos.extend(range(n))
        cu.append(cu[-1] + n)
        nblk = (n + block_size - 1) // block_size
        for j in range(nblk):
            cnt = block_size if j != nblk - 1 else n - (nblk - 1) * block_size
            slots.extend(range(table[j] * block_size, table[j] * block_size + cnt))
    cu_t = torch.tensor(cu, dtype=torch.int32)
    width = max(len(t) for t in block_tables_py)
    bt = torch.tensor([t + [-1] * (width - len(t)) for t in block_tables_py], dtype=torch.int32)
    meta = StepMeta(True, torch.tensor(slots, dtype=torch.int32), cu_t, cu_t.clone(), None, bt)
    return torch.tensor(ids_pos, dtype=torch.int64), meta


def decode_meta(seq_lens, block_tables_py, block_size):
    """model_runner.py:344-366 (prepare_decode): one token per sequence at position len-1,
    context_len = len, slot = (last block, len-1 mod bs) -- flattened for the oracle's
    store_kvcache, which takes flat slots (attention_torch_native.py:30-31)."""
    pos, slots = [], []
    for n, table in zip(seq_lens, block_tables_py):
        pos.append(n - 1)
        slots.append(table[(n - 1) // block_size] * block_size + (n - 1) % block_size)
    width = max(len(t) for t in block_tables_py)
    bt = torch.tensor([t + [-1] * (width - len(t)) for t in block_tables_py], dtype=torch.int32)
    meta = StepMeta(False, torch.tensor(slots, dtype=torch.int32), None, None,
                    torch.tensor(seq_lens, dtype=torch.int32), bt)
    return torch.tensor(pos, dtype=torch.int64), meta


@dataclass
class GreedyTrace:
    tokens: list[list[int]] = field(default_factory=list)   # per sequence
    logits: list[torch.Tensor] = field(default_factory=list)  # per step [B, vocab] fp32


def generate_greedy(w, cfg: ModelCfg, prompts, max_tokens, block_size=16, num_blocks=None,
                    attn_mode="reference", keep_logits=True, fp32_logits=False, force_tokens=None):
    """Greedy (argmax) generation for a fixed batch of prompts, all decoded in lock step;
    blocks are handed out sequentially (sequence i gets a private ascending run).
    Follows the engine loop llm_engine.py:120-176 + model_runner.py:398-452 with the sampler
    replaced by argmax (sampling_params.py:11 forbids T=0 in the reference front end)."""
    B = len(prompts)
    need = [(len(p) + max_tokens + block_size - 1) // block_size for p in prompts]
    if num_blocks is None:
        num_blocks = sum(need) + 1
    tables, nxt = [], 0
    for n in need:
        tables.append(list(range(nxt, nxt + n)))
        nxt += n
    kv = torch.zeros(2, cfg.num_layers, num_blocks, block_size, cfg.num_kv_heads, cfg.head_dim,
                     dtype=cfg.dtype)
    seqs = [list(p) for p in prompts]
    trace = GreedyTrace(tokens=[[] for _ in range(B)])
    with torch.inference_mode():
        lens = [len(s) for s in seqs]
        positions, meta = prefill_meta(lens, [t[:(n + block_size - 1) // block_size]
                                              for t, n in zip(tables, lens)], block_size)
        ids = torch.tensor([t for s in seqs for t in s], dtype=torch.int64)
        logits = forward(w, cfg, ids, positions, meta, kv, block_size, attn_mode, fp32_logits=fp32_logits)
        for step in range(max_tokens):
            nxt_tok = L.greedy(logits).tolist()
            if keep_logits:
                trace.logits.append(logits.float().clone())
            for i, t in enumerate(nxt_tok):
                seqs[i].append(t if force_tokens is None else force_tokens[step][i])
                trace.tokens[i].append(t)
            if step == max_tokens - 1:
                break
            lens = [len(s) for s in seqs]
            positions, meta = decode_meta(lens, [t[:(n + block_size - 1) // block_size]
                                                 for t, n in zip(tables, lens)], block_size)
            ids = torch.tensor([s[-1] for s in seqs], dtype=torch.int64)
            logits = forward(w, cfg, ids, positions, meta, kv, block_size, attn_mode, fp32_logits=fp32_logits)
    return trace
