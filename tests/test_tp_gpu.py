"""-m gpu, needs >= 2 GPUs: tensor-parallel engine (NCCL all-reduce in the row-parallel linears,
vocab-parallel embedding / LM head gather, sampled ids broadcast) against the single-GPU engine."""
import random

import pytest
import torch

pytestmark = pytest.mark.gpu


def _need_gpus(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs, have {torch.cuda.device_count()}")


def _run(tp, prompts, enforce_eager, port):
    from nanovllm import LLM, SamplingParams
    llm = LLM("tiny-qwen3", seed=3, norm_jitter=0.1, kvcache_block_size=16, max_num_seqs=8,
              max_model_len=256, max_num_batched_tokens=1024, num_kvcache_blocks=128,
              tensor_parallel_size=tp, enforce_eager=enforce_eager, hccl_port=port)
    outs = llm.generate(prompts, SamplingParams(temperature=1.0, max_tokens=24, ignore_eos=True,
                                                greedy=True), use_tqdm=False)
    llm.exit()
    return [o["token_ids"] for o in outs]


@pytest.mark.parametrize("enforce_eager", [True, False])
def test_tp2_matches_tp1(enforce_eager):
    _need_gpus(2)
    rnd = random.Random(8)
    prompts = [[rnd.randint(0, 511) for _ in range(n)] for n in (9, 40, 17)]
    want = _run(1, prompts, True, 28011)
    got = _run(2, prompts, enforce_eager, 28013 if enforce_eager else 28015)
    # the sharded sums are reassociated (two partial products + all-reduce), so a token may only
    # differ where the logits nearly tie; the tiny model's margins are >= 0.1, far above that
    assert got == want
