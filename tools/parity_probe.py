"""Debug: per-step logit error of the config-1 run vs the reference golden."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nano-vllm-ascend_b200"))
from nanovllm import LLM, SamplingParams
g = torch.load(os.path.join(ROOT, "tests/golden/models.pt"), weights_only=False)["qwen3_0_6b_config1"]
llm = LLM("qwen3-0.6b", seed=0, enforce_eager=True, max_num_seqs=4, kvcache_block_size=16, max_model_len=256,
          max_num_batched_tokens=1024, num_kvcache_blocks=128)
logs = []
orig = llm.model_runner.sampler.forward
def spy(logits, *a, **k):
    logs.append(logits.float().cpu()); return orig(logits, *a, **k)
llm.model_runner.sampler.forward = spy
out = llm.generate(g["prompts"], SamplingParams(temperature=1.0, max_tokens=32, ignore_eos=True, greedy=True), use_tqdm=False)
got = torch.stack(logs)[:, 0]
err = (got.gather(1, g["top_indices"]) - g["top_values"]).abs().max(-1).values
print(os.environ.get("B200_PREFILL_KERNEL", "umma"), "per-step err:", [round(float(e), 3) for e in err])
print("tokens equal:", out[0]["token_ids"] == g["tokens"][0])
llm.exit()
