"""Test infrastructure: import and drive the UNMODIFIED reference modules on CPU.

The reference package cannot be imported whole (torch_npu / torchair are absent), so a namespace
stub is registered and only its CPU-capable modules are imported: the torch-native attention
(layers/attention_torch_native.py), the eager layers and models/qwen3.py with the NPU attention
module replaced by a constructor-adapting shim around the torch-native class. The
runner metadata (prepare_prefill / prepare_decode, engine/model_runner.py:238-366, NPU-bound) is
restated in oracle/model_ref.py.

Two locations serve as the reference tree:
  /root/reference        the build container (oracle/gen_golden.py writes tests/golden/ from it)
  oracle/_ref/           reference_path.zip: the handful of modules this path needs, packed by
                         oracle/stage_ref.py; git-ignored, travels to the GPU box with the snapshot
                         so that `bench.py --impl reference` and the `cpu_baseline` leg time the
                         reference's own code there (cpu_baseline.kind = "reference").
Only tests/, __graft_entry__.smoke()/build() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import os
import sys
import time
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref", "reference_path.zip")


def reference_dir() -> str | None:
    """Where to import the reference from: $B200_REFERENCE_DIR, /root/reference, or the archive
    oracle/stage_ref.py packed (zipimport serves the unmodified modules straight from it)."""
    if os.environ.get("B200_REFERENCE_DIR") == "archive":        # tests: force the packed copy
        return STAGED if os.path.isfile(STAGED) else None
    for cand in (os.environ.get("B200_REFERENCE_DIR"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "nanovllm", "layers", "attention_torch_native.py")):
            return cand
    return STAGED if os.path.isfile(STAGED) else None


def import_reference():
    pkg = types.ModuleType("nanovllm")
    ref = reference_dir()
    if ref is None:
        raise RuntimeError("no reference tree: run oracle/stage_ref.py in the build container")
    pkg.__path__ = [os.path.join(ref, "nanovllm")]
    sys.modules["nanovllm"] = pkg
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29713", world_size=1, rank=0)
    from nanovllm.layers import attention_torch_native as native

    class ShimAttention(native.Attention):
        """models/*.py construct Attention(num_heads, head_dim, scaling, num_kv_heads);
        the torch-native class takes (num_heads, head_dim, num_kv_heads, block_size)."""
        block_size = 16

        def __init__(self, num_heads, head_dim, scaling, num_kv_heads):
            super().__init__(num_heads, head_dim, num_kv_heads, ShimAttention.block_size)

    shim = types.ModuleType("nanovllm.layers.attention")
    shim.Attention = ShimAttention
    sys.modules["nanovllm.layers.attention"] = shim
    return native, ShimAttention


def set_ctx(meta):
    from nanovllm.utils.context import set_context
    set_context(meta.is_prefill, cu_seqlens_q=meta.cu_seqlens_q, cu_seqlens_k=meta.cu_seqlens_k,
                max_seqlen_q=0, max_seqlen_k=0, slot_mapping=meta.slot_mapping,
                context_lens=meta.context_lens, block_tables=meta.block_tables,
                is_enforce_eager=True, real_bs=-1, block_size=16)


def build_reference_model(cfg, weights):
    """Reference Qwen3ForCausalLM (models/qwen3.py) on CPU in cfg.dtype with `weights`."""
    from transformers import Qwen3Config
    from nanovllm.models.qwen3 import Qwen3ForCausalLM
    from nanovllm.layers.rotary_embedding import get_rope
    get_rope.cache_clear()
    hf = Qwen3Config(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size,
                     num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads,
                     num_key_value_heads=cfg.num_kv_heads, head_dim=cfg.head_dim,
                     intermediate_size=cfg.intermediate_size, rms_norm_eps=cfg.rms_eps,
                     max_position_embeddings=cfg.max_position, tie_word_embeddings=True,
                     attention_bias=False, hidden_act="silu")
    old = torch.get_default_dtype()
    torch.set_default_dtype(cfg.dtype)
    try:
        model = Qwen3ForCausalLM(hf)
    finally:
        torch.set_default_dtype(old)
    params = dict(model.named_parameters())
    for name, w in weights.items():
        params[name].data.copy_(w)
    model.lm_head.weight.data = model.model.embed_tokens.weight.data
    return model.eval()


def reference_greedy(model, cfg, prompts, max_tokens, bs):
    """Drive the reference model with restated runner metadata; argmax sampling."""
    from oracle import model_ref as M
    from nanovllm.utils.context import reset_context
    B = len(prompts)
    need = [(len(p) + max_tokens + bs - 1) // bs for p in prompts]
    nb = sum(need) + 1
    tables, nxt = [], 0
    for n in need:
        tables.append(list(range(nxt, nxt + n)))
        nxt += n
    kv = torch.zeros(2, cfg.num_layers, nb, bs, cfg.num_kv_heads, cfg.head_dim, dtype=cfg.dtype)
    lid = 0
    for mod in model.modules():
        if hasattr(mod, "k_cache") and hasattr(mod, "v_cache"):
            mod.k_cache, mod.v_cache = kv[0, lid], kv[1, lid]
            lid += 1
    seqs = [list(p) for p in prompts]
    toks, logs = [[] for _ in range(B)], []
    with torch.inference_mode():
        lens = [len(s) for s in seqs]
        pos, meta = M.prefill_meta(lens, [t[:(n + bs - 1) // bs] for t, n in zip(tables, lens)], bs)
        ids = torch.tensor([t for s in seqs for t in s], dtype=torch.int64)
        set_ctx(meta)
        logits = model.compute_logits(model(ids, pos))
        for step in range(max_tokens):
            nxt_tok = logits.float().argmax(-1).tolist()
            logs.append(logits.float().clone())
            for i, t in enumerate(nxt_tok):
                seqs[i].append(t)
                toks[i].append(t)
            if step == max_tokens - 1:
                break
            lens = [len(s) for s in seqs]
            pos, meta = M.decode_meta(lens, [t[:(n + bs - 1) // bs] for t, n in zip(tables, lens)], bs)
            ids = torch.tensor([s[-1] for s in seqs], dtype=torch.int64)
            set_ctx(meta)
            logits = model.compute_logits(model(ids, pos))
        reset_context()
    return toks, logs



def reference_decode_step(n_seqs, ctx, block_size, steps, warmup, threads=None):
    """One full-model decode step of the reference's own Qwen3ForCausalLM (models/qwen3.py, eager
    layers, torch-native paged attention, bf16, CPU) for `n_seqs` sequences at context `ctx`; the
    same workload slice oracle port `bench.cpu_reference_step` times. Returns (tokens/s, s/step)."""
    from oracle import model_ref as M
    if threads:
        torch.set_num_threads(threads)
    _, shim = import_reference()          # first: registers the stub package the next import resolves in
    from nanovllm.utils.context import reset_context
    shim.block_size = block_size
    cfg = M.ModelCfg.qwen3_0_6b()
    model = build_reference_model(cfg, M.seeded_weights(cfg, seed=0))
    nblk_per = (ctx + 1 + block_size - 1) // block_size
    nb = n_seqs * nblk_per + 1
    g = torch.Generator().manual_seed(0)
    kv = (0.5 * torch.randn(2, cfg.num_layers, nb, block_size, cfg.num_kv_heads, cfg.head_dim,
                            generator=g)).to(cfg.dtype)
    lid = 0
    for mod in model.modules():
        if hasattr(mod, "k_cache") and hasattr(mod, "v_cache"):
            mod.k_cache, mod.v_cache = kv[0, lid], kv[1, lid]
            lid += 1
    tables = [list(range(i * nblk_per, (i + 1) * nblk_per)) for i in range(n_seqs)]
    lens = [ctx] * n_seqs
    ids = torch.randint(0, 10000, (n_seqs,), generator=g)
    times = []
    with torch.inference_mode():
        for i in range(warmup + steps):
            pos, meta = M.decode_meta(lens, tables, block_size)
            t0 = time.perf_counter()
            set_ctx(meta)
            logits = model.compute_logits(model(ids, pos))
            ids = logits.float().argmax(-1)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
        reset_context()
    sec = sum(times) / len(times)
    return n_seqs / sec, sec


def main(argv=None):
    """`python -m oracle.ref_runtime --seqs N --ctx C ...`: one JSON line with the timing of the
    reference's decode step (bench.py runs this in a child process: the stub package registered
    here shadows the product's own `nanovllm` package name)."""
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("--seqs", type=int, default=8)
    ap.add_argument("--ctx", type=int, default=1024)
    ap.add_argument("--block-size", type=int, default=16)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args(argv)
    threads = a.threads or os.cpu_count() or 1
    tps, sec = reference_decode_step(a.seqs, a.ctx, a.block_size, a.steps, a.warmup, threads)
    print(json.dumps({"tokens_per_s": tps, "s_per_step": sec, "threads": torch.get_num_threads(),
                      "seqs": a.seqs, "ctx": a.ctx, "source": reference_dir()}))


if __name__ == "__main__":
    main()
