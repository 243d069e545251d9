import os, sys, random, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nano-vllm-ascend_b200"))
from oracle import model_ref as M
from nanovllm import LLM
from nanovllm.engine.sequence import Sequence
from nanovllm.utils.context import reset_context
cfg = M.ModelCfg(); w = M.seeded_weights(cfg, seed=7, norm_jitter=0.1)
llm = LLM("tiny-qwen3", seed=7, norm_jitter=0.1, enforce_eager=True, kvcache_block_size=16, max_num_seqs=8,
          max_model_len=256, max_num_batched_tokens=1024, num_kvcache_blocks=128)
r = llm.model_runner
rnd = random.Random(0); prompt = [rnd.randint(0, 511) for _ in range(20)]
seq = Sequence(prompt, block_size=16); seq.block_table = [3, 5]
got = {}
def hook(name):
    def f(mod, inp, out):
        got[name] = (out[0] if isinstance(out, tuple) else out).detach().float().cpu().clone()
    return f
for i, layer in enumerate(r.model.model.layers):
    p = f"model.layers.{i}."
    layer.self_attn.qkv_proj.register_forward_hook(hook(p + "qkv"))
    layer.self_attn.attn.register_forward_hook(hook(p + "attn_out"))
    layer.self_attn.o_proj.register_forward_hook(hook(p + "o_proj"))
    layer.mlp.register_forward_hook(hook(p + "mlp_out"))
    layer.input_layernorm.register_forward_hook(hook(p + "attn_in"))
    layer.post_attention_layernorm.register_forward_hook(hook(p + "mlp_in"))
with torch.inference_mode():
    ids, pos = r.prepare_prefill([seq])
    logits = r.model.compute_logits(r.model(ids, pos)).float().cpu()
reset_context()
trace = {}
kv = torch.zeros(2, cfg.num_layers, 128, 16, cfg.num_kv_heads, cfg.head_dim, dtype=cfg.dtype)
positions, meta = M.prefill_meta([20], [[3, 5]], 16)
with torch.inference_mode():
    ref = M.forward(w, cfg, torch.tensor(prompt), positions, meta, kv, 16, attn_mode="exact", trace=trace)
for k in trace:
    if k in got:
        a, b = got[k], trace[k].float().reshape(got[k].shape)
        print(f"{k:45s} max|d|={float((a-b).abs().max()):.4f} ref|max|={float(b.abs().max()):.3f}")
print("logits", float((logits - ref.float()).abs().max()))
emb = r.model.model.embed_tokens.weight.detach().cpu()
print("emb equal", torch.equal(emb, w["model.embed_tokens.weight"]))
for n, p_ in r.model.named_parameters():
    if n in w and not torch.equal(p_.detach().cpu(), w[n]): print("MISMATCH weight", n)
llm.exit()
