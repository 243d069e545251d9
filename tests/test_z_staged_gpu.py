"""-m gpu, OPT-IN (B200_RUN_STAGED=1): code paths that compile and pass their CPU-side checks
(tools/simulate_umma2_protocol.py, tools/emulate_fp8_fragments.py) but have not yet run on a B200.
They are staged here so that the default `pytest -m gpu` run only exercises validated kernels; each
case runs in a child process with a timeout, because a first run on hardware may deadlock.

  B200_RUN_STAGED=1 python -m pytest tests/test_z_staged_gpu.py -q

Move a case into the regular GPU test files once it has passed on hardware.
"""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("B200_RUN_STAGED") != "1",
                                 reason="staged (not yet validated on hardware); set B200_RUN_STAGED=1")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pytest_child(args, env_extra, timeout):
    env = dict(os.environ, **env_extra)
    env.pop("B200_RUN_STAGED", None)
    try:
        r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", *args], cwd=ROOT, env=env,
                           capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired as e:
        pytest.fail(f"child pytest {args} timed out after {timeout} s (deadlock?)\n{(e.stdout or b'')[-3000:]}")
    assert r.returncode == 0, r.stdout[-6000:] + r.stderr[-2000:]


def test_prefill_umma2_kernel_passes_the_prefill_parity_tests():
    """prefill_umma2.cu (128-key S tiles, P from tensor memory) against the same golden / oracle
    tests as the default prefill kernel."""
    _pytest_child(["tests/test_kernels_gpu.py", "-k", "prefill"], {"B200_PREFILL_KERNEL": "umma2"}, 600)


def test_prefill_umma2_engine_parity():
    _pytest_child(["tests/test_engine_gpu.py", "-k", "tiny_model or ragged or prefix_cache or llama"],
                  {"B200_PREFILL_KERNEL": "umma2"}, 900)


_CHUNKED = r"""
import random, sys
sys.path[:0] = [r"%s", r"%s"]
from nanovllm import LLM, SamplingParams
rnd = random.Random(12)
prompts = [[rnd.randint(0, 511) for _ in range(n)] for n in (300, 40, 17, 200, 129)]
prompts.append(prompts[0][:160] + [3] * 30)            # shares 10 full blocks with the first prompt
sp = SamplingParams(temperature=1.0, max_tokens=12, ignore_eos=True, greedy=True)
outs = []
for kw in (dict(max_num_batched_tokens=1024), dict(max_num_batched_tokens=128, chunked_prefill=True)):
    llm = LLM("tiny-qwen3", seed=3, norm_jitter=0.1, kvcache_block_size=16, max_num_seqs=8,
              max_model_len=512, num_kvcache_blocks=128, **kw)
    outs.append([o["token_ids"] for o in llm.generate(prompts, sp, use_tqdm=False)])
    llm.exit()
assert outs[0] == outs[1], outs
print("chunked ok")
""" % (ROOT, os.path.join(ROOT, "nano-vllm-ascend_b200"))


def test_chunked_prefill_engine_matches_unchunked():
    """chunked_prefill=True with a 128-token step budget against whole-prompt prefill: same greedy
    tokens (only validated kernels are involved: K8 scatter + paged prefill attention; staged because
    the host logic has only run on CPU)."""
    r = subprocess.run([sys.executable, "-c", _CHUNKED], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "chunked ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_nvls_allreduce_kernel_and_tp2_engine():
    """Multicast all-reduce (comm.py `_map_symmetric`, allreduce.cu nvls kernel): kernel cases on 2
    GPUs, then the TP=2 engine against TP=1 with every fused all-reduce forced through it."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _pytest_child(["tests/test_tp_gpu.py", "-k", "nvls"], {"B200_NVLS_ALLREDUCE": "1"}, 400)
    _pytest_child(["tests/test_tp_gpu.py", "-k", "tp2_matches_tp1"], {"B200_NVLS_ALLREDUCE": "all"}, 600)


def test_moe_kernels_and_engine():
    """csrc/moe.cu + layers/moe.py + models/qwen3_moe.py: tests/test_z_moe_gpu.py."""
    _pytest_child(["tests/test_z_moe_gpu.py"], {"B200_RUN_MOE": "1"}, 900)


def test_decode_l2_weight_prefetch_keeps_results():
    """B200_DECODE_L2_PREFETCH=1: the side-stream L2 prefetch next to decode attention (forked and
    joined inside the captured graph) must not change a single token."""
    _pytest_child(["tests/test_engine_gpu.py", "-k", "tiny_model or cuda_graph_decode_equals_eager"],
                  {"B200_DECODE_L2_PREFETCH": "1"}, 600)


def test_sharded_sampling_equals_full_vocabulary_sampling():
    """b200_sample_shard on 1..8 vocabulary shards + a MAX over the keys == b200_sample on the whole
    row, bit for bit: greedy rows, temperature rows (the Gumbel draw is keyed by the GLOBAL index),
    ties (lowest index wins), a shard that is entirely -inf. Single GPU: the all-reduce is a torch max."""
    import torch
    from nanovllm import ops
    g = torch.Generator().manual_seed(3)
    rows, vocab = 37, 151936
    logits = (3 * torch.randn(rows, vocab, generator=g)).to(torch.bfloat16).cuda()
    logits[5, 1000:1010] = 30.0                      # a 10-way tie for the maximum inside one shard
    logits[6, 10] = logits[6, vocab - 7] = 31.0      # a tie across the first and the last shard
    logits[7, : vocab // 2] = float("-inf")          # an all -inf shard (world = 2)
    temps = torch.rand(rows, generator=g) * 1.5
    temps[::3] = 0.0                                 # every third row greedy
    temps = temps.cuda()
    off = torch.tensor([11], dtype=torch.int64, device="cuda")
    for world in (1, 2, 4, 8):
        per = vocab // world
        for t in (None, temps):
            want = ops.sample(logits, t, seed=5, step=9, step_offset=off)
            keys = None
            for r in range(world):
                shard = logits[:, r * per:(r + 1) * per].contiguous()
                k = ops.sample_shard(shard, t, seed=5, step=9, index_offset=r * per, step_offset=off)
                keys = k if keys is None else torch.maximum(keys, k)
            got = ops.decode_sample_keys(keys)
            assert torch.equal(got, want), (world, t is None, (got != want).nonzero().flatten().tolist())
    assert int(want[5]) >= 0 and int(ops.sample(logits, None)[5]) == 1000 and int(ops.sample(logits, None)[6]) == 10


def test_sharded_sampling_tp2_engine_matches_tp1():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _pytest_child(["tests/test_tp_gpu.py", "-k", "tp2_matches_tp1"], {"B200_SHARDED_SAMPLING": "1"}, 600)
