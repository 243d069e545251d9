"""The reference's own throughput benchmark (bench/bench.py:16-42) through this engine, same
workload and same arithmetic: seed(0); 256 sequences, prompt length U(100, 1024) of random token
ids, output length U(100, 1024), temperature 0.6, ignore_eos; throughput = total OUTPUT tokens /
wall time of one `LLM.generate` call (prefill included), after a one-request warm-up.

Its published numbers for this exact workload (other hardware: 1 x Ascend 910C, README.md:315-350):
Qwen3-0.6B graph mode 3954.20 tok/s, eager 845-880 tok/s; Llama-3.2-1B 5648.50 tok/s. This is NOT
the contract metric of bench.py (steady-state decode tokens/s at bs=256, ctx=1024); it is the
number to put beside the reference's README table.

    python tools/ref_bench_workload.py [--model qwen3-0.6b] [--block-size 256] [--enforce-eager]
"""
import argparse
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nano-vllm-ascend_b200"))


def workload(n_requests: int = 256, longest_prompt: int = 1024, longest_output: int = 1024):
    """The request set of the reference's benchmark, drawn in its order from Random(0): every prompt's
    length then its ids, and only afterwards every output budget (bench/bench.py:29-31)."""
    rng = random.Random(0)
    prompts = []
    for _ in range(n_requests):
        length = rng.randint(100, longest_prompt)
        prompts.append([rng.randint(0, 10000) for _ in range(length)])
    budgets = [rng.randint(100, longest_output) for _ in range(n_requests)]
    return prompts, budgets


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="qwen3-0.6b")
    ap.add_argument("--block-size", type=int, default=256, help="reference default (config.py:25)")
    ap.add_argument("--max-num-seqs", type=int, default=256)
    ap.add_argument("--tensor-parallel-size", type=int, default=1)
    ap.add_argument("--enforce-eager", action="store_true")
    ap.add_argument("--kv-cache-dtype", default="auto")
    args = ap.parse_args()
    from nanovllm import LLM, SamplingParams
    prompts, budgets = workload()
    llm = LLM(args.model, enforce_eager=args.enforce_eager, max_model_len=4096,
              max_num_seqs=args.max_num_seqs, tensor_parallel_size=args.tensor_parallel_size,
              kvcache_block_size=args.block_size, kv_cache_dtype=args.kv_cache_dtype)
    params = [SamplingParams(temperature=0.6, ignore_eos=True, max_tokens=n) for n in budgets]
    llm.generate([[1, 2, 3, 4]], SamplingParams(max_tokens=16), use_tqdm=False)     # warm-up request
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    llm.generate(prompts, params, use_tqdm=False)
    t = time.perf_counter() - t0
    total_tokens = sum(budgets)
    print(json.dumps({"workload": "reference bench/bench.py (256 seqs, in U(100,1024), out U(100,1024))",
                      "model": args.model, "kvcache_block_size": args.block_size,
                      "tensor_parallel_size": args.tensor_parallel_size, "enforce_eager": args.enforce_eager,
                      "kv_cache_dtype": args.kv_cache_dtype, "total_output_tokens": total_tokens,
                      "seconds": round(t, 3), "throughput_tok_s": round(total_tokens / t, 2),
                      "reference_published_tok_s": {"qwen3-0.6b": 3954.20, "llama-3.2-1b": 5648.50}.get(args.model)}))
    llm.exit()


if __name__ == "__main__":
    main()
