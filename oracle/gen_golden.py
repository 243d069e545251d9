"""Generate tests/golden/*.pt by running the UNMODIFIED reference code on CPU.

Run in the build container only (needs /root/reference):  python -m oracle.gen_golden
The reference package cannot be imported whole (torch_npu/torchair are absent), so a
namespace stub is registered and only its CPU-capable modules are imported: the torch-native
attention (layers/attention_torch_native.py), the eager layers, and models/qwen3.py with the
NPU attention module replaced by a constructor-adapting shim around the torch-native class
(SURVEY.md 8c). The runner metadata (prepare_prefill / prepare_decode) is restated in
oracle/model_ref.py because engine/model_runner.py is NPU-bound.

Outputs are inputs + reference outputs for seeded cases; tests/test_oracle_golden.py replays
them through oracle/*.py (bit-exact on CPU) and the -m gpu tests replay them on the B200.
"""
from __future__ import annotations

import os
import random
import sys
import types

import torch

from oracle.ref_runtime import (build_reference_model, import_reference, reference_greedy,  # noqa: E402,F401
                                set_ctx)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def gen_kernels(native):
    from oracle import layers_ref  # only for the rope table constructor arguments
    g = torch.Generator().manual_seed(1234)
    bf = torch.bfloat16
    out = {}

    # ---- KV scatter (attention_torch_native.store_kvcache)
    N, G, D, nb, bs = 40, 2, 64, 6, 16
    key = torch.randn(N, G, D, generator=g).to(bf)
    val = torch.randn(N, G, D, generator=g).to(bf)
    kc = torch.randn(nb, bs, G, D, generator=g).to(bf)
    vc = torch.randn(nb, bs, G, D, generator=g).to(bf)
    slots = torch.randperm(nb * bs, generator=g)[:N].to(torch.int32)
    slots[3] = -1
    slots[17] = -1
    slots[29] = nb * bs + 5          # out of range -> dropped
    kc_out, vc_out = kc.clone(), vc.clone()
    native.store_kvcache(key, val, kc_out, vc_out, slots, bs)
    out["kv_scatter"] = dict(key=key, value=val, k_cache=kc, v_cache=vc, slot_mapping=slots,
                             block_size=bs, k_cache_out=kc_out, v_cache_out=vc_out)

    # ---- decode + prefill attention (Attention._decode_attn / _prefill_attn)
    for tag, (H, G, D) in {"d64": (4, 2, 64), "d128": (4, 2, 128)}.items():
        nb, bs = 14, 16
        attn = native.Attention(H, D, G, bs)
        attn.k_cache = (0.5 * torch.randn(nb, bs, G, D, generator=g)).to(bf)
        attn.v_cache = torch.randn(nb, bs, G, D, generator=g).to(bf)
        ctx = torch.tensor([1, 16, 17, 40, 33, 64], dtype=torch.int32)
        perm = torch.randperm(nb, generator=g).tolist()
        tables, p = [], 0
        for n in ctx.tolist():
            k = (n + bs - 1) // bs
            tables.append(perm[p:p + k] if p + k <= nb else perm[:k])
            p = (p + k) % nb
        width = max(len(t) for t in tables)
        bt = torch.tensor([t + [-1] * (width - len(t)) for t in tables], dtype=torch.int32)
        q = torch.randn(len(ctx), H, D, generator=g).to(bf)
        o = attn._decode_attn(q, ctx, bt, torch.device("cpu")).reshape(len(ctx), -1)
        out[f"decode_{tag}"] = dict(q=q, k_cache=attn.k_cache, v_cache=attn.v_cache,
                                    context_lens=ctx, block_tables=bt, num_kv_heads=G,
                                    head_dim=D, out=o)
        lens = [5, 16, 33, 70]
        cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
        T = int(cu[-1])
        q = torch.randn(T, H, D, generator=g).to(bf)
        k = (0.5 * torch.randn(T, G, D, generator=g)).to(bf)
        v = torch.randn(T, G, D, generator=g).to(bf)
        o = attn._prefill_attn(q, k, v, cu, cu, torch.device("cpu")).reshape(T, -1)
        out[f"prefill_{tag}"] = dict(q=q, k=k, v=v, cu_seqlens=cu, num_kv_heads=G, head_dim=D,
                                     out=o)

    # ---- eager layers
    from nanovllm.layers.layernorm import RMSNorm
    from nanovllm.layers.rotary_embedding import RotaryEmbedding
    from nanovllm.layers.activation import SiluAndMul
    for hidden in (128, 1024):
        norm = RMSNorm(hidden, 1e-6)
        norm.weight.data = (1 + 0.1 * torch.randn(hidden, generator=g)).to(bf)
        x = torch.randn(7, hidden, generator=g).to(bf)
        r = torch.randn(7, hidden, generator=g).to(bf)
        y = norm(x.clone())
        y2, r2 = norm(x.clone(), r.clone())
        out[f"rmsnorm_{hidden}"] = dict(x=x, residual=r, weight=norm.weight.data.clone(), eps=1e-6,
                                        out=y, add_out=y2, add_residual=r2)
    for D in (64, 128):
        rope = RotaryEmbedding(D, D, 256, 1000000.0)
        pos = torch.tensor([0, 1, 2, 17, 100, 255], dtype=torch.int64)
        q = torch.randn(6, 4, D, generator=g).to(bf)
        k = torch.randn(6, 2, D, generator=g).to(bf)
        qo, ko = rope(pos, q, k)
        out[f"rope_{D}"] = dict(positions=pos, q=q, k=k, base=1000000.0, max_position=256,
                                q_out=qo, k_out=ko, cos_sin_cache=rope.cos_sin_cache.clone())
    x = (3 * torch.randn(5, 2 * 96, generator=g)).to(bf)
    out["silu_and_mul"] = dict(x=x, out=SiluAndMul()(x))
    return out


def gen_models():
    from oracle import model_ref as M
    out = {}
    # tiny Qwen3-shaped model: full logits kept
    cfg = M.ModelCfg()
    w = M.seeded_weights(cfg, seed=7, norm_jitter=0.1)
    model = build_reference_model(cfg, w)
    rnd = random.Random(0)
    prompts = [[rnd.randint(0, cfg.vocab_size - 1) for _ in range(n)] for n in (20, 5, 33)]
    toks, logs = reference_greedy(model, cfg, prompts, 10, 16)
    out["tiny_qwen3"] = dict(seed=7, norm_jitter=0.1, prompts=prompts, max_tokens=10,
                             block_size=16, tokens=toks, logits=torch.stack(logs))
    # BASELINE config 1: full-size random-init Qwen3-0.6B, 1 prompt, block 16, 32 tokens
    cfg = M.ModelCfg.qwen3_0_6b()
    w = M.seeded_weights(cfg, seed=0)
    model = build_reference_model(cfg, w)
    rnd = random.Random(0)
    prompts = [[rnd.randint(0, 10000) for _ in range(24)]]
    toks, logs = reference_greedy(model, cfg, prompts, 32, 16)
    lg = torch.stack(logs)[:, 0]                      # [32, vocab]
    top = lg.topk(8, dim=-1)
    out["qwen3_0_6b_config1"] = dict(seed=0, norm_jitter=0.0, prompts=prompts, max_tokens=32,
                                     block_size=16, tokens=toks, top_values=top.values,
                                     top_indices=top.indices,
                                     logit_mean=lg.mean(-1), logit_std=lg.std(-1))
    # The same configuration with a livelier init (std 0.04, norm weights 1 + 0.1 N(0,1), seed 1) and 64
    # tokens: the N(0, 0.02) model above collapses to one repeated token, which makes "greedy ids
    # identical" thin evidence; this one emits ~60 distinct ids with a median top-1/top-2 margin of 0.25.
    # (Larger init std makes the random model chaotic: at 0.05 the reference's own bf16-rounding attention
    # and an fp32 evaluation of the same graph already differ by 0.2-0.4 in the logits.)
    w = M.seeded_weights(cfg, seed=1, norm_jitter=0.1, std=0.04)
    model = build_reference_model(cfg, w)
    toks, logs = reference_greedy(model, cfg, prompts, 64, 16)
    lg = torch.stack(logs)[:, 0]
    top = lg.topk(8, dim=-1)
    out["qwen3_0_6b_config1_varied"] = dict(seed=1, norm_jitter=0.1, init_std=0.04, prompts=prompts,
                                            max_tokens=64, block_size=16, tokens=toks,
                                            top_values=top.values, top_indices=top.indices,
                                            logit_mean=lg.mean(-1), logit_std=lg.std(-1))
    return out


def moe_block_weights(E, hidden, inter, seed, dtype=torch.bfloat16):
    """Seeded weights of one sparse-MoE block (regenerated identically by the tests, so the
    fixture stores outputs only): router [E, h] ~ N(0, 1) (spread-out routing probabilities),
    expert matrices ~ N(0, 0.1)."""
    g = torch.Generator().manual_seed(seed)
    gate = torch.randn(E, hidden, generator=g).to(dtype)
    w13 = (0.1 * torch.randn(E, 2 * inter, hidden, generator=g)).to(dtype)
    w2 = (0.1 * torch.randn(E, hidden, inter, generator=g)).to(dtype)
    return gate, w13, w2


MOE_BLOCK_CASES = {            # name: (tokens, hidden, moe_intermediate, experts, top_k, seed)
    "e8_k2": (37, 128, 64, 8, 2, 11),
    "e32_k4": (64, 128, 64, 32, 4, 12),
    "e128_k8": (48, 256, 64, 128, 8, 13),
    "single_token": (1, 128, 64, 8, 2, 14),
}


def gen_moe():
    """Qwen3MoeSparseMoeBlock (models/qwen3_moe.py:125-186) and a tiny Qwen3MoeForCausalLM, both
    the unmodified reference classes on CPU."""
    from transformers import Qwen3MoeConfig
    from nanovllm.models.qwen3_moe import Qwen3MoeForCausalLM, Qwen3MoeSparseMoeBlock
    from nanovllm.layers.rotary_embedding import get_rope
    from oracle import model_ref as M
    from oracle import moe_ref
    out = {}
    old = torch.get_default_dtype()
    for name, (T, hidden, inter, E, k, seed) in MOE_BLOCK_CASES.items():
        hf = Qwen3MoeConfig(hidden_size=hidden, moe_intermediate_size=inter, num_experts=E,
                            num_experts_per_tok=k, hidden_act="silu")
        torch.set_default_dtype(torch.bfloat16)
        try:
            block = Qwen3MoeSparseMoeBlock(hf)
        finally:
            torch.set_default_dtype(old)
        gate, w13, w2 = moe_block_weights(E, hidden, inter, seed)
        block.gate.weight.data.copy_(gate)
        for e, ex in enumerate(block.experts):
            ex.gate_up_proj.weight.data.copy_(w13[e])
            ex.down_proj.weight.data.copy_(w2[e])
        g = torch.Generator().manual_seed(seed + 100)
        x = torch.randn(T, hidden, generator=g).to(torch.bfloat16)
        with torch.inference_mode():
            y = block(x)
        rw, ids, logits = moe_ref.route(x, gate, k)      # recorded for the kernel-level tests
        out[name] = dict(tokens=T, hidden=hidden, inter=inter, experts=E, top_k=k, seed=seed, x=x,
                         out=y.clone(), topk_ids=ids.to(torch.int32), topk_weights=rw, logits=logits)
    # tiny Qwen3-MoE model, greedy
    cfg = M.ModelCfg.tiny_qwen3_moe()
    w = M.seeded_weights(cfg, seed=5, norm_jitter=0.1)
    get_rope.cache_clear()
    hf = Qwen3MoeConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size,
                        num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads,
                        num_key_value_heads=cfg.num_kv_heads, head_dim=cfg.head_dim,
                        intermediate_size=cfg.intermediate_size,
                        moe_intermediate_size=cfg.moe_intermediate_size, num_experts=cfg.num_experts,
                        num_experts_per_tok=cfg.top_k, decoder_sparse_step=1, mlp_only_layers=[],
                        rms_norm_eps=cfg.rms_eps, max_position_embeddings=cfg.max_position,
                        tie_word_embeddings=False, attention_bias=False, hidden_act="silu")
    torch.set_default_dtype(cfg.dtype)
    try:
        model = Qwen3MoeForCausalLM(hf)
    finally:
        torch.set_default_dtype(old)
    params = dict(model.named_parameters())
    assert set(params) == set(w), (set(params) ^ set(w))
    for pname, t in w.items():
        params[pname].data.copy_(t)
    model.eval()
    rnd = random.Random(3)
    prompts = [[rnd.randint(0, cfg.vocab_size - 1) for _ in range(n)] for n in (18, 7, 33, 16)]
    toks, logs = reference_greedy(model, cfg, prompts, 10, 16)
    out["tiny_qwen3_moe"] = dict(seed=5, norm_jitter=0.1, prompts=prompts, max_tokens=10,
                                 block_size=16, tokens=toks, logits=torch.stack(logs))
    return out


def gen_scheduler_trace():
    """Drive the reference's own Scheduler + BlockManager (engine/scheduler.py,
    engine/block_manager.py, imported unmodified) through a seeded workload that exercises
    prefix sharing, block growth, finishing and preemption; record every step's decision."""
    from types import SimpleNamespace
    from nanovllm.engine.scheduler import Scheduler
    from nanovllm.engine.sequence import Sequence
    from nanovllm.sampling_params import SamplingParams
    traces = {}
    for name, nb, bs, max_seqs, budget in (("roomy", 64, 16, 8, 256), ("tight", 14, 16, 6, 128)):
        cfg = SimpleNamespace(max_num_seqs=max_seqs, max_num_batched_tokens=budget, eos=7,
                              num_kvcache_blocks=nb, kvcache_block_size=bs, is_multimodal=False,
                              hf_config=None, max_model_len=200)
        sched = Scheduler(cfg)
        rnd = random.Random(42)
        shared = [rnd.randint(10, 99) for _ in range(40)]
        reqs, seqs = [], []
        for i in range(12):
            prompt = (shared[:rnd.choice([16, 32, 40])] if i % 3 == 0 else []) + \
                [rnd.randint(10, 99) for _ in range(rnd.randint(1, 30))]
            sp = dict(temperature=0.6, max_tokens=rnd.randint(1, 40), ignore_eos=(i % 4 != 0))
            reqs.append(dict(prompt=prompt, **sp))
            s = Sequence(prompt, SamplingParams(**sp), block_size=bs)
            seqs.append(s)
            sched.add(s)
        index = {s.seq_id: i for i, s in enumerate(seqs)}
        steps = []
        tok_rnd = random.Random(1)
        while not sched.is_finished() and len(steps) < 2000:
            batch, is_prefill = sched.schedule()
            rec = dict(is_prefill=is_prefill,
                       seqs=[(index[s.seq_id], list(s.block_table), s.num_cached_tokens, len(s)) for s in batch])
            toks = [tok_rnd.randint(5, 12) for _ in batch]     # 7 == eos shows up now and then
            rec["tokens"] = toks
            sched.postprocess(batch, toks)
            rec["finished"] = [index[s.seq_id] for s in batch if s.is_finished]
            steps.append(rec)
        traces[name] = dict(num_blocks=nb, block_size=bs, max_num_seqs=max_seqs,
                            max_num_batched_tokens=budget, eos=7, max_model_len=200,
                            requests=reqs, steps=steps)
    return traces


def main():
    torch.manual_seed(0)
    native, _ = import_reference()
    os.makedirs(OUT, exist_ok=True)
    k = gen_kernels(native)
    torch.save(k, os.path.join(OUT, "kernels.pt"))
    m = gen_models()
    torch.save(m, os.path.join(OUT, "models.pt"))
    moe = gen_moe()
    torch.save(moe, os.path.join(OUT, "moe.pt"))
    tr = gen_scheduler_trace()
    torch.save(tr, os.path.join(OUT, "scheduler.pt"))
    print({k: len(v["steps"]) for k, v in tr.items()})
    print("tiny moe tokens", moe["tiny_qwen3_moe"]["tokens"])
    for name in ("kernels.pt", "models.pt", "moe.pt", "scheduler.pt"):
        print(name, os.path.getsize(os.path.join(OUT, name)), "bytes")
    print("tiny tokens", m["tiny_qwen3"]["tokens"])
    print("0.6B tokens", m["qwen3_0_6b_config1"]["tokens"])


if __name__ == "__main__":
    main()
