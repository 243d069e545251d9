"""CPU tests (-m "not gpu"): host logic, the C-ABI surface, and the N>1 sharding logic on gloo."""
import ctypes
import os
import re
import socket
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ C ABI surface
def test_library_exports_every_declared_symbol():
    """Every function declared in include/b200attn.h is exported by the built library and bound
    by the ctypes layer (no compute calls: there is no GPU here)."""
    from nanovllm import _C
    header = open(os.path.join(ROOT, "include", "b200attn.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_C.PROTOTYPES), declared ^ set(_C.PROTOTYPES)
    lib = _C.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.b200_version() == 100
    assert b"unsupported" in lib.b200_error_string(-3)
    assert lib.b200_launch_count() == 0


def test_ctypes_prototypes_match_the_header_signatures():
    """Argument count and argument class (pointer / int32 / int64 / uint64 / size_t / float) of every
    ctypes prototype against the C declaration: a mismatch would not fail to load, it would pass
    garbage to a kernel launch."""
    from nanovllm import _C
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "b200attn.h")).read(), flags=re.S)
    decls = re.findall(r"\b(?:int|int64_t|size_t|void|const char\*)\s+\**(b200_\w+)\s*\(([^;]*?)\)\s*;",
                       header, flags=re.S)
    assert len(decls) >= 30
    scalar = {"float": ctypes.c_float, "int32_t": ctypes.c_int32, "int": ctypes.c_int32,
              "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64, "size_t": ctypes.c_size_t}

    def ctype_of(arg: str):
        if "*" in arg or "b200_stream_t" in arg:
            return ctypes.c_void_p
        return scalar[arg.replace("const", "").split()[0]]
    for name, args in decls:
        args = args.strip()
        want = [] if args in ("", "void") else [ctype_of(a.strip()) for a in args.split(",")]
        got = list(_C.PROTOTYPES[name][1])
        assert len(want) == len(got), (name, len(want), len(got))
        for i, (w, g) in enumerate(zip(want, got)):
            assert w is g or (w is ctypes.c_void_p and g is ctypes.c_char_p), (name, i, w, g)


def test_bad_arguments_are_rejected_without_a_gpu():
    """Argument validation happens before any CUDA call, so it can be exercised on CPU."""
    from nanovllm import _C
    lib = _C.load()
    one = ctypes.c_void_p(16)
    # head_dim 96 is unsupported; null pointers are rejected; bad dtype
    assert lib.b200_qknorm_rope_cache(one, one, one, None, None, 1e-6, None, None, None, 0, 4, 4, 2,
                                      96, 4 * 96 * 2, 128, 0, 0, 0, None) == -3
    assert lib.b200_rmsnorm(None, one, one, 1e-6, 2, 64, 0, None) == -1
    assert lib.b200_rmsnorm(one, one, one, 1e-6, 2, 64, 5, None) == -7
    assert lib.b200_silu_and_mul(one, one, 2, 12, 0, None) == -4          # inter % 8 != 0
    assert lib.b200_paged_decode_attn_workspace_size(4, 24, 2, 128) > 0
    assert lib.b200_paged_decode_attn(one, one, one, one, one, one, 4, 24, 2, 128, 24 * 128, 24 * 128,
                                      8, 16, 4, 0.1, one, 1 << 30, 0, 0, None) == -3   # group 12 > 8
    assert lib.b200_paged_decode_attn(one, one, one, one, one, one, 4, 16, 8, 128, 2048, 2048, 8, 16,
                                      4, 0.1, one, 16, 0, 0, None) == -5               # workspace
    assert lib.b200_reshape_and_cache(one, one, one, one, one, 4, 8, 128, 1028, 1024, 8, 16, 0,
                                      None) == -4                                      # stride % 8
    # fp8 KV-cache entry points: scales must be positive; head_dim 64 needs >= 2 kv heads; q 8-byte aligned
    assert lib.b200_reshape_and_cache_fp8(one, one, one, one, one, 0, 4, 8, 128, 1024, 1024, 8, 16, 0.0,
                                          1.0, 0, None) == -2
    assert lib.b200_paged_decode_attn_fp8(one, one, one, one, one, one, 4, 8, 1, 64, 512, 512, 8, 16, 4,
                                          0.1, 1.0, 1.0, one, 1 << 30, 0, 0, None) == -3
    assert lib.b200_paged_decode_attn_fp8(one, ctypes.c_void_p(20), one, one, one, one, 4, 16, 8, 128,
                                          2048, 2048, 8, 16, 4, 0.1, 1.0, 1.0, one, 1 << 30, 0, 0,
                                          None) == -4
    # two-shot all-reduce: world out of range
    assert lib.b200_allreduce_add_rmsnorm_twoshot(one, one, one, 1e-6, 4, 64, one, one, one, one, 9, 0, 0,
                                                  None) == -2


def test_ops_refuse_cpu_tensors():
    from nanovllm import ops
    x = torch.zeros(2, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="CUDA only"):
        ops.rmsnorm(x, torch.ones(64, dtype=torch.bfloat16), 1e-6)
    with pytest.raises(RuntimeError, match="CUDA only"):
        ops.silu_and_mul(x)


def test_fp8_oracle_and_fragment_map():
    """The e4m3 statement the fp8 kernels are tested against: every finite byte round-trips, ties go
    to even, overflow saturates at +-448; and the lane <-> data mapping of paged_decode_fp8_kernel
    (permuted contraction index, even/odd column split), emulated lane by lane in numpy."""
    from oracle import attention_ref as A
    b = torch.arange(256, dtype=torch.uint8)
    f = A.dequantize_e4m3(b)
    finite = ~torch.isnan(f)
    assert int(finite.sum()) == 254
    assert torch.equal(A.quantize_e4m3(f[finite]), b[finite])
    x = torch.tensor([1000.0, -1000.0, 1.0625, 1.1875, 464.0, 3 * 2.0 ** -11, 2.0 ** -10])  # last: tie -> even (0)
    assert A.dequantize_e4m3(A.quantize_e4m3(x)).tolist() == [448.0, -448.0, 1.0, 1.25, 448.0, 2.0 ** -9, 0.0]
    assert A.dequantize_e4m3(A.quantize_e4m3(x[:4], 2.0), 2.0).tolist() == [896.0, -896.0, 1.0, 1.25]
    kc = torch.zeros(2, 16, 2, 8, dtype=torch.uint8)
    vc = torch.zeros_like(kc)
    k = torch.randn(3, 2, 8)
    A.store_kvcache_fp8(k, k, kc, vc, torch.tensor([17, -1, 3]), 16, 0.5, 1.0)
    assert torch.equal(kc[1, 1], A.quantize_e4m3(k[0], 0.5)) and torch.equal(vc[0, 3], A.quantize_e4m3(k[2]))
    assert int(kc.count_nonzero()) <= 2 * 16
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import emulate_fp8_fragments as E
    for D, heads in ((128, (0, 3)), (64, (0, 1))):
        for h in heads:
            E.run(D, h, G=2, valid=16, seed=1)
            E.run(D, h, G=8, valid=5, seed=2)


def test_sampling_params_and_config():
    from nanovllm import SamplingParams
    from nanovllm.config import Config
    sp = SamplingParams()
    assert (sp.temperature, sp.max_tokens, sp.ignore_eos, sp.greedy) == (1.0, 64, False, False)
    with pytest.raises(AssertionError):
        SamplingParams(temperature=0.0)
    c = Config("qwen3-0.6b", kvcache_block_size=16)
    assert c.hf_config.num_hidden_layers == 28 and c.max_model_len == 4096
    assert c.kvcache_block_size == 16 and c.device == "cuda"
    assert Config("qwen3-0.6b").kvcache_block_size == 256        # reference default
    with pytest.raises(AssertionError):
        Config("qwen3-0.6b", kvcache_block_size=24)
    with pytest.raises(AssertionError):
        Config("no-such-model")


def test_sequence_block_math_and_wire_state():
    import pickle
    from nanovllm import SamplingParams
    from nanovllm.engine.sequence import Sequence
    s = Sequence(list(range(37)), SamplingParams(temperature=0.5, max_tokens=9), block_size=16)
    assert (s.num_blocks, s.last_block_num_tokens, s.block(2)) == (3, 5, list(range(32, 37)))
    s.block_table = [4, 2, 9]
    t = pickle.loads(pickle.dumps(s))
    # the reference drops block_size on the wire; workers need it
    assert t.num_blocks == 3 and t.block_table == [4, 2, 9] and t.max_tokens == 9


# ------------------------------------------------------------------ scheduler vs the reference
@pytest.mark.parametrize("name", ["roomy", "tight"])
def test_scheduler_trace_matches_reference(name):
    """Replay the workload the reference's own Scheduler/BlockManager were driven through
    (tests/golden/scheduler.pt) and require identical decisions at every step: batch
    composition, prefill/decode, block tables, cached-token counts, finish events."""
    from types import SimpleNamespace
    from nanovllm import SamplingParams
    from nanovllm.engine.scheduler import Scheduler
    from nanovllm.engine.sequence import Sequence
    g = torch.load(os.path.join(ROOT, "tests", "golden", "scheduler.pt"), weights_only=False)[name]
    cfg = SimpleNamespace(max_num_seqs=g["max_num_seqs"], max_num_batched_tokens=g["max_num_batched_tokens"],
                          eos=g["eos"], num_kvcache_blocks=g["num_blocks"],
                          kvcache_block_size=g["block_size"], max_model_len=g["max_model_len"])
    sched = Scheduler(cfg)
    seqs = []
    for r in g["requests"]:
        s = Sequence(r["prompt"], SamplingParams(temperature=r["temperature"], max_tokens=r["max_tokens"],
                                                 ignore_eos=r["ignore_eos"]), block_size=g["block_size"])
        seqs.append(s)
        sched.add(s)
    index = {s.seq_id: i for i, s in enumerate(seqs)}
    preempted = 0
    # One documented deviation: the reference never resets num_cached_tokens, so over re-admissions
    # of a preempted sequence its counter is the SUM of every admission's hits (and it never reads
    # it in the runner); ours is per admission. Accumulate ours to compare like with like.
    ref_style_cached = [0] * len(seqs)
    for step, rec in enumerate(g["steps"]):
        assert not sched.is_finished()
        batch, is_prefill = sched.schedule()
        if is_prefill:
            for s in batch:
                ref_style_cached[index[s.seq_id]] += s.num_cached_tokens
        got = [(index[s.seq_id], list(s.block_table), ref_style_cached[index[s.seq_id]], len(s)) for s in batch]
        assert is_prefill == rec["is_prefill"], step
        assert got == rec["seqs"], step
        sched.postprocess(batch, rec["tokens"])
        assert [index[s.seq_id] for s in batch if s.is_finished] == rec["finished"], step
        preempted += sum(1 for s in seqs if s.finish_reason is not None and s.finish_reason.name == "PREEMPTED")
    assert sched.is_finished()
    if name == "tight":
        assert preempted > 0, "the tight pool is meant to exercise preemption"


def test_block_manager_prefix_sharing():
    from nanovllm.engine.block_manager import BlockManager
    from nanovllm.engine.sequence import Sequence
    bm = BlockManager(8, 16)
    a = Sequence(list(range(40)), block_size=16)
    bm.allocate(a)
    b = Sequence(list(range(32)) + [7] * 8, block_size=16)
    bm.allocate(b)
    assert a.block_table[:2] == b.block_table[:2] and a.block_table[2] != b.block_table[2]
    assert b.num_cached_tokens == 32 and bm.blocks[a.block_table[0]].ref_count == 2
    bm.deallocate(a)
    assert bm.blocks[b.block_table[0]].ref_count == 1
    bm.deallocate(b)
    assert len(bm.free_block_ids) == 8 and not bm.used_block_ids
    # freed blocks keep their content hash: a new identical prompt re-uses them
    c = Sequence(list(range(40)), block_size=16)
    bm.allocate(c)
    assert c.num_cached_tokens == 32


def test_chunked_prefill_scheduling():
    """chunked_prefill=True (extension): a prompt longer than the step's token budget is
    computed window by window; windows tile the prompt exactly, the budget is never exceeded, nothing is
    admitted behind a half-computed prompt, only the final window yields a token, and a prompt that
    leaves half-way takes the hashes of its never-computed blocks with it."""
    from types import SimpleNamespace
    from nanovllm import SamplingParams
    from nanovllm.engine.scheduler import Scheduler
    from nanovllm.engine.sequence import Sequence
    cfg = SimpleNamespace(max_num_seqs=8, max_num_batched_tokens=100, eos=-1, num_kvcache_blocks=65,
                          kvcache_block_size=16, max_model_len=4096, chunked_prefill=True)
    sched = Scheduler(cfg)
    sp = SamplingParams(temperature=1.0, max_tokens=3, ignore_eos=True)
    a = Sequence(list(range(1000, 1250)), sp, block_size=16)     # 250 tokens: 3 windows
    b = Sequence(list(range(2000, 2030)), sp, block_size=16)     # 30 tokens
    c = Sequence(list(range(1000, 1064)) + [5] * 20, sp, block_size=16)   # shares a's first 64 tokens
    for s in (a, b, c):
        sched.add(s)
    seen = {id(a): [], id(b): [], id(c): []}
    steps = []
    while sched.waiting:
        batch, is_prefill = sched.schedule()
        assert is_prefill and batch
        used = 0
        for s in batch:
            start, end = s.prefill_window
            seen[id(s)].append((start, end))
            used += end - start
        assert used <= 100
        steps.append([(len(s), s.prefill_window) for s in batch])
        sched.postprocess(batch, [7] * len(batch))
    # a: three windows tiling [0, 250); b rides along with a's last window; c re-uses a's 4 computed blocks
    assert seen[id(a)] == [(0, 100), (100, 200), (200, 250)]
    assert seen[id(b)] == [(0, 30)] and steps[2][1] == (30, (0, 30))
    assert c.num_cached_tokens == 64 and seen[id(c)] == [(64, 84)]
    assert len(steps[0]) == 1 and len(steps[1]) == 1           # nothing admitted behind a half-computed prompt
    assert (len(a), len(b), len(c)) == (251, 31, 85)           # one token each, from the final windows only
    assert all(s.num_computed_tokens == s.num_prompt_tokens for s in (a, b, c))
    batch, is_prefill = sched.schedule()
    assert not is_prefill and len(batch) == 3
    # abort half-way: the uncomputed blocks must not stay in the prefix cache
    d = Sequence(list(range(5000, 5200)), sp, request_id="d", block_size=16)
    sched.add(d)
    batch, _ = sched.schedule()
    assert batch == [d] and d.prefill_window == (0, 100)
    sched.postprocess(batch, [7])
    assert d.num_computed_tokens == 100 and len(d) == 200
    bm = sched.block_manager
    registered = sum(1 for blk in d.block_table if bm.blocks[blk].hash != -1)
    assert registered == 12                                     # allocate() registered all 12 full blocks
    sched.abort_seq_group("d")
    e = Sequence(list(range(5000, 5200)), sp, block_size=16)
    sched.add(e)
    batch, _ = sched.schedule()
    assert batch == [e] and e.num_cached_tokens == 96           # only the 6 computed blocks count as cached
    assert e.prefill_window == (96, 196)


def test_readmission_resets_cached_tokens():
    """A preempted sequence that had prefix hits must not carry its old num_cached_tokens into the
    next admission (the reference accumulates it, block_manager.py:62-88, but never reads it in the
    runner; here prepare_prefill and the chunked scheduler do)."""
    from nanovllm.engine.block_manager import BlockManager
    from nanovllm.engine.sequence import Sequence
    bm = BlockManager(8, 16)
    s = Sequence(list(range(10)), block_size=16)
    for _ in range(3):
        bm.allocate(s)
        assert s.num_cached_tokens == 0
        bm.deallocate(s)
    a = Sequence(list(range(40)), block_size=16)
    bm.allocate(a)
    bm.deallocate(a)
    for _ in range(3):                      # 2 full blocks are found in the prefix cache each time
        bm.allocate(a)
        assert a.num_cached_tokens == 32 < len(a)
        bm.deallocate(a)


def test_abort_mid_reprefill_forgets_uncomputed_blocks():
    """Chunked mode: a preempted sequence WITH completions is re-prefilled over len(seq) tokens; if it
    is aborted between windows, hashes of the not yet recomputed blocks must leave the prefix cache."""
    from types import SimpleNamespace
    from nanovllm import SamplingParams
    from nanovllm.engine.scheduler import Scheduler
    from nanovllm.engine.sequence import Sequence
    cfg = SimpleNamespace(max_num_seqs=4, max_num_batched_tokens=64, eos=-1, num_kvcache_blocks=33,
                          kvcache_block_size=16, max_model_len=4096, chunked_prefill=True)
    sched = Scheduler(cfg)
    sp = SamplingParams(temperature=1.0, max_tokens=50, ignore_eos=True)
    a = Sequence(list(range(100, 228)), sp, request_id="a", block_size=16)     # 128 tokens: 2 windows
    sched.add(a)
    for _ in range(2):
        batch, is_prefill = sched.schedule()
        assert is_prefill
        sched.postprocess(batch, [7])
    for _ in range(20):                     # decode past another full block
        batch, is_prefill = sched.schedule()
        assert not is_prefill
        sched.postprocess(batch, [7])
    assert a.num_completion_tokens == 21
    sched.running.remove(a)
    sched.preempt(a)
    # wipe the prefix cache so the re-admission has no hits and starts at token 0
    bm = sched.block_manager
    bm.hash_to_block_id.clear()
    for blk in bm.blocks:
        blk.hash, blk.token_ids = -1, []
    batch, is_prefill = sched.schedule()
    assert is_prefill and a.prefill_window == (0, 64) and a.num_cached_tokens == 0
    sched.postprocess(batch, [7])
    assert len(bm.hash_to_block_id) == 9    # allocate() registered all 9 full blocks up front
    sched.abort_seq_group("a")
    assert len(bm.hash_to_block_id) == 4    # only the 4 recomputed blocks stay
    b = Sequence(a.token_ids[:], sp, block_size=16)
    sched.add(b)
    batch, _ = sched.schedule()
    assert batch == [b] and b.num_cached_tokens == 64


def test_decode_wire_message_is_small_and_complete():
    """What rank 0 ships to tensor-parallel workers per step (model_runner.call): a decode step must
    not carry the token history (256 x 1100 tokens pickled to 1.19 MB and overflowed the segment)."""
    import pickle
    from nanovllm import SamplingParams
    from nanovllm.engine.sequence import Sequence
    seqs = []
    for i in range(256):
        s = Sequence([151000 + (i + j) % 900 for j in range(1100)], SamplingParams(temperature=0.6),
                     block_size=16)
        s.block_table = list(range(i * 69, i * 69 + 69))
        seqs.append(s)
    Sequence.wire_decode = True
    try:
        blob = pickle.dumps(["run", seqs, False], protocol=pickle.HIGHEST_PROTOCOL)
    finally:
        Sequence.wire_decode = False
    assert len(blob) < 200_000, len(blob)
    _, got, _ = pickle.loads(blob)
    for s, t in zip(seqs, got):
        assert (t.seq_id, len(t), t.last_token, t.block_table, t.temperature, t.greedy, t.num_blocks) == \
               (s.seq_id, len(s), s.last_token, s.block_table, s.temperature, s.greedy, s.num_blocks)
        assert t.token_ids == []
    blob = pickle.dumps(["run", seqs[:16], True], protocol=pickle.HIGHEST_PROTOCOL)
    assert len(blob) < 16 * (4 * 1100 + 1024)
    _, got, _ = pickle.loads(blob)
    assert all(t.token_ids == s.token_ids for s, t in zip(seqs, got))


def test_sampler_noise_key_is_the_device_counter_alone():
    """With a device step counter the host step passed to the kernel is 0 for every launch, so the
    value baked into a CUDA graph at capture cannot alias another bucket's."""
    from nanovllm.layers.sampler import Sampler
    s = Sampler(seed=3)
    assert [s._next_step(object()) for _ in range(4)] == [0, 0, 0, 0] and s.step == 0
    assert [s._next_step(None) for _ in range(3)] == [1, 2, 3]


def test_engine_step_on_an_unschedulable_queue_raises():
    from types import SimpleNamespace
    from nanovllm import SamplingParams
    from nanovllm.engine.llm_engine import LLMEngine
    from nanovllm.engine.scheduler import Scheduler
    from nanovllm.engine.sequence import Sequence
    cfg = SimpleNamespace(max_num_seqs=4, max_num_batched_tokens=4096, eos=-1, num_kvcache_blocks=4,
                          kvcache_block_size=16, max_model_len=4096)
    eng = LLMEngine.__new__(LLMEngine)
    eng.scheduler = Scheduler(cfg)
    eng.model_runner = SimpleNamespace(call=lambda *a: pytest.fail("run() on an empty batch"))
    assert eng.step() == ([], 0)            # idle
    eng.scheduler.add(Sequence(list(range(64)), SamplingParams(), block_size=16))   # 4 blocks, pool has 3
    with pytest.raises(RuntimeError, match="KV blocks"):
        eng.step()


# ------------------------------------------------------------------ N>1: TP shards on gloo
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _tp_worker(rank, world, port, out_q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "nano-vllm-ascend_b200"))
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", world_size=world, rank=rank)
    from nanovllm.layers.linear import (MergedColumnParallelLinear, QKVParallelLinear, RowParallelLinear)
    from nanovllm.layers.embed_head import VocabParallelEmbedding
    torch.manual_seed(0)
    hidden, H, Hkv, D, inter, vocab = 64, 4, 2, 16, 96, 50
    x = torch.randn(5, hidden)
    wq, wk, wv = torch.randn(H * D, hidden), torch.randn(Hkv * D, hidden), torch.randn(Hkv * D, hidden)
    wo = torch.randn(hidden, H * D)
    wg, wu, wd = torch.randn(inter, hidden), torch.randn(inter, hidden), torch.randn(hidden, inter)
    emb = torch.randn(vocab, hidden)
    qkv = QKVParallelLinear(hidden, D, H, Hkv)
    for w, sid in ((wq, "q"), (wk, "k"), (wv, "v")):
        qkv.weight.weight_loader(qkv.weight, w, sid)
    o = RowParallelLinear(H * D, hidden)
    o.weight.weight_loader(o.weight, wo)
    gu = MergedColumnParallelLinear(hidden, [inter, inter])
    gu.weight.weight_loader(gu.weight, wg, 0)
    gu.weight.weight_loader(gu.weight, wu, 1)
    dn = RowParallelLinear(inter, hidden)
    dn.weight.weight_loader(dn.weight, wd)
    e = VocabParallelEmbedding(vocab, hidden)
    e.weight.weight_loader(e.weight, emb)
    with torch.no_grad():
        y = qkv(x)
        hq, hkv = H // world * D, Hkv // world * D
        q_loc, k_loc, v_loc = y.split([hq, hkv, hkv], -1)
        # stand-in for attention: any function that is independent per head
        attn = q_loc * 0.5
        o_out = o(attn)                                   # all-reduce inside
        g_u = gu(x)
        half = inter // world
        mlp = dn(torch.nn.functional.silu(g_u[:, :half]) * g_u[:, half:])
        ids = torch.tensor([0, 3, 24, 25, 49])
        emb_out = e(ids)
        want_o = (x @ wq.T * 0.5) @ wo.T
        want_mlp = (torch.nn.functional.silu(x @ wg.T) * (x @ wu.T)) @ wd.T
        ok = (torch.allclose(o_out, want_o, atol=1e-3, rtol=1e-4) and
              torch.allclose(mlp, want_mlp, atol=1e-3, rtol=1e-4) and
              torch.equal(emb_out, emb[ids]) and
              torch.equal(q_loc, (x @ wq.T)[:, rank * hq:(rank + 1) * hq]) and
              torch.equal(k_loc, (x @ wk.T)[:, rank * hkv:(rank + 1) * hkv]))
    out_q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_tensor_parallel_shards_world_size_2_gloo():
    """Column/row/QKV/merged shards + the row-parallel all-reduce and the vocab-parallel
    embedding reproduce the unsharded layers (reference: layers/linear.py:54-153,
    embed_head.py:34-42). world_size 2 on gloo."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert results == {0: True, 1: True}


def test_seeded_init_matches_oracle_weights():
    """Product and oracle draw identical random-init weights by parameter name (CPU check on the
    tiny preset; no kernels involved)."""
    from oracle import model_ref as M
    from nanovllm.config import preset_hf_config
    from nanovllm.models.qwen3 import Qwen3ForCausalLM
    from nanovllm.utils.loader import seeded_init
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        model = Qwen3ForCausalLM(preset_hf_config("tiny-qwen3"))
    finally:
        torch.set_default_dtype(prev)
    seeded_init(model, seed=7, norm_jitter=0.1)
    want = M.seeded_weights(M.ModelCfg(), seed=7, norm_jitter=0.1)
    got = dict(model.named_parameters())
    assert set(want) <= set(got)
    for name, w in want.items():
        assert torch.equal(got[name].data, w), name
    assert model.lm_head.weight.data_ptr() == model.model.embed_tokens.weight.data_ptr()


# ------------------------------------------------------------------ Qwen3-MoE: host side
def _build_tiny_moe(tp=None, seed=None):
    """Tiny Qwen3-MoE on CPU, optionally as rank tp[0] of tp[1], optionally seeded-initialised
    (the initialiser needs the same tensor-parallel view as the constructor)."""
    from nanovllm.config import preset_hf_config
    from nanovllm.layers.linear import set_tp_world
    from nanovllm.models.qwen3_moe import Qwen3MoeForCausalLM
    from nanovllm.utils.loader import seeded_init
    prev = torch.get_default_dtype()
    try:
        if tp is not None:
            set_tp_world(*tp)
        torch.set_default_dtype(torch.bfloat16)
        try:
            model = Qwen3MoeForCausalLM(preset_hf_config("tiny-qwen3-moe"))
        finally:
            torch.set_default_dtype(prev)       # the initialiser draws in fp32, like the runner
        if seed is not None:
            seeded_init(model, seed=seed, norm_jitter=0.1)
        return model
    finally:
        set_tp_world(None)


def test_moe_model_parameters_and_seeded_init_match_oracle():
    """Reference parameter names (qwen3_moe.py) over stacked expert storage; the seeded initialiser
    fills the stacks exactly as the oracle's per-expert weights."""
    from oracle import model_ref as M
    from nanovllm.utils.loader import seeded_init
    cfg = M.ModelCfg.tiny_qwen3_moe()
    model = _build_tiny_moe()
    want = M.seeded_weights(cfg, seed=5, norm_jitter=0.1)
    got = dict(model.named_parameters())
    assert set(got) == set(want)
    seeded_init(model, seed=5, norm_jitter=0.1)
    for name, w in want.items():
        assert torch.equal(got[name].data, w), name
    moe = model.model.layers[1].mlp
    for e in range(cfg.num_experts):        # the kernels read the stacks, not the per-expert views
        assert torch.equal(moe.w13[e], want[f"model.layers.1.mlp.experts.{e}.gate_up_proj.weight"])
        assert torch.equal(moe.w2[e], want[f"model.layers.1.mlp.experts.{e}.down_proj.weight"])
    # checkpoint-style loading: gate_proj / up_proj halves routed by the packed-module mapping
    p = got["model.layers.0.mlp.experts.3.gate_up_proj.weight"]
    half = torch.full((cfg.moe_intermediate_size, cfg.hidden_size), 2.0, dtype=torch.bfloat16)
    p.weight_loader(p, half, 1)
    blk = model.model.layers[0].mlp
    assert bool((blk.w13[3, cfg.moe_intermediate_size:] == 2).all())
    assert torch.equal(blk.w13[3, :cfg.moe_intermediate_size],
                       want["model.layers.0.mlp.experts.3.gate_up_proj.weight"][:cfg.moe_intermediate_size])


def test_moe_expert_parallel_layout_and_kv_head_replication():
    """tp=4 over the tiny MoE model (8 experts, 4 q heads, 2 kv heads): each rank owns 2 whole experts
    with the weights the single-rank model has for them, and one kv head replicated on 2 ranks."""
    from oracle import model_ref as M
    from nanovllm.utils.loader import seeded_init
    cfg = M.ModelCfg.tiny_qwen3_moe()
    want = M.seeded_weights(cfg, seed=5, norm_jitter=0.1)
    D = cfg.head_dim
    for rank in range(4):
        model = _build_tiny_moe(tp=(rank, 4), seed=5)
        blk = model.model.layers[0].mlp
        assert (blk.num_local, blk.expert_begin) == (2, 2 * rank)
        assert sorted(blk.experts.keys()) == [str(2 * rank), str(2 * rank + 1)]
        virtual = model.virtual_param_shapes()
        assert len(virtual) == 2 * 2 * 6 and not (set(virtual) & set(dict(model.named_parameters())))
        for i in range(2):
            e = 2 * rank + i
            assert torch.equal(blk.w13[i], want[f"model.layers.0.mlp.experts.{e}.gate_up_proj.weight"])
            assert torch.equal(blk.w2[i], want[f"model.layers.0.mlp.experts.{e}.down_proj.weight"])
        assert torch.equal(blk.gate.weight.data, want["model.layers.0.mlp.gate.weight"])
        attn = model.model.layers[0].self_attn
        assert (attn.num_heads, attn.num_kv_heads, attn.qkv_proj.kv_replicas) == (1, 1, 2)
        full = want["model.layers.0.self_attn.qkv_proj.weight"]
        q, k, v = full.split([cfg.num_heads * D, cfg.num_kv_heads * D, cfg.num_kv_heads * D])
        kv = rank // 2
        mine = torch.cat([q[rank * D:(rank + 1) * D], k[kv * D:(kv + 1) * D], v[kv * D:(kv + 1) * D]])
        assert torch.equal(attn.qkv_proj.weight.data, mine)


def test_moe_entry_points_reject_bad_arguments():
    from nanovllm import _C
    lib = _C.load()
    one = ctypes.c_void_p(16)
    assert lib.b200_moe_route(one, one, one, 4, 128, 9, 128, 0, None) == -3          # top_k > 8
    assert lib.b200_moe_route(one, one, one, 4, 512, 8, 512, 0, None) == -3          # > 256 experts
    assert lib.b200_moe_route(one, one, one, 4, 8, 9, 8, 0, None) == -2              # top_k > experts
    assert lib.b200_moe_route(None, one, one, 4, 8, 2, 8, 0, None) == -1
    assert lib.b200_moe_align(one, one, one, one, one, 8, 0, 0, None) == -2
    assert lib.b200_moe_gemm_gate_up(one, one, one, one, one, 100, 64, 4, 2, 100, 0, None) == -3   # K % 64
    assert lib.b200_moe_gemm_gate_up(one, one, one, one, one, 128, 96, 4, 2, 128, 0, None) == -3   # inter % 64
    assert lib.b200_moe_gemm_down(one, one, one, one, one, one, 64, 192, 4, 2, 0, None) == -3      # hidden % 128
    assert lib.b200_moe_gemm_down(one, one, one, one, one, None, 64, 128, 4, 2, 0, None) == -1
    assert lib.b200_moe_combine(one, one, one, one, 4, 2, 100, 100, 0, None) == -4                 # hidden % 8
    assert lib.b200_moe_combine(one, one, one, one, 0, 2, 128, 128, 0, None) == 0                  # empty batch
    assert lib.b200_sample_shard(one, one, None, 4, 1000, 1000, 2 ** 31 - 500, 0, 0, None, 0, None) == -2   # index overflow
    assert lib.b200_sample_shard(None, one, None, 4, 1000, 1000, 0, 0, 0, None, 0, None) == -1
    # tcgen05 grouped GEMMs (moe_umma.cu)
    assert lib.b200_moe_gemm_gate_up_umma(one, one, one, one, 256, 100, 64, 4, 0, None) == -3       # K % 64
    assert lib.b200_moe_gemm_down_umma(one, one, one, one, one, one, 256, 64, 192, 4, 0, None) == -3  # hidden % 128
    assert lib.b200_moe_gemm_down_umma(one, one, one, one, None, one, 256, 64, 128, 4, 0, None) == -1   # scale without pairs
    assert lib.b200_moe_gemm_gate_up_umma(one, one, one, one, 0, 128, 64, 4, 0, None) == 0          # no pairs
    assert lib.b200_moe_gather_rows(one, one, one, one, 16, 4, 2, 100, 100, 0, None) == -4          # hidden % 8


def test_moe_kernel_index_emulation():
    """tools/emulate_moe_kernels.py: thread-level transcription of the align / expert-GEMM / combine
    index arithmetic against the oracle (the part of moe.cu a CPU can check)."""
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emulate_moe_kernels.py")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.count("ok ") >= 10, r.stdout[-2000:] + r.stderr[-2000:]


def test_fused_allreduce_form_selection():
    """comm.PeerAllReduce.mode(): which fused all-reduce form serves a message (pure host logic)."""
    from nanovllm.comm import PeerAllReduce
    c = object.__new__(PeerAllReduce)
    c.world, c.max_bytes, c.multicast_ptr, c.forced = 8, 4 << 20, 0, "auto"
    assert c.mode(256, 1024, 2) == "one"            # Qwen3-0.6B TP=8: 7 x 512 KiB incoming
    assert c.mode(256, 5120, 2) == "two"            # Qwen3-32B TP=8: 18 MB incoming -> row-sliced two-shot
    assert c.mode(4, 5120 * 8, 2) is None           # hidden beyond the kernels' row width
    assert c.mode(4096, 5120, 2) is None            # prefill-sized: beyond the buffer -> NCCL
    assert c.mode(256, 1030, 2) is None             # hidden % 8
    c.world = 2
    assert c.mode(256, 5120, 2) == "one"            # TP=2: 2.6 MB incoming
    c.world = 8
    for forced, big, small, few_rows in (("nccl", None, None, None), ("oneshot", "one", "one", "one"),
                                         ("twoshot", "two", "two", "one"), ("nvls", "nvls", "nvls", "one"),
                                         ("nvls_fused", "nvls_fused", "nvls_fused", "one"),
                                         ("nvls_one", "nvls_one", "nvls_one", "nvls_one")):
        c.forced = forced
        assert (c.mode(256, 5120, 2), c.mode(256, 1024, 2), c.mode(4, 8192, 2)) == (big, small, few_rows), forced
        assert c.mode(4096, 5120, 2) is None        # a forced form never takes what does not fit the region


def _ep_worker(rank, world, port, out_q):
    """Expert-parallel layout on a real process group (gloo): every rank builds its SparseMoeBlock
    shard, fills it through the checkpoint-style weight loaders, evaluates ITS experts' partial sum with
    the oracle on the shard's stacked weights, and the all-reduced result must equal the reference
    block's golden output."""
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "nano-vllm-ascend_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", world_size=world, rank=rank)
    from nanovllm.layers.moe import SparseMoeBlock
    from oracle import moe_ref
    from test_oracle_golden import moe_block_weights
    g = torch.load(os.path.join(ROOT, "tests", "golden", "moe.pt"), weights_only=False)["e32_k4"]
    E, k, hidden, inter = g["experts"], g["top_k"], g["hidden"], g["inter"]
    gate, w13, w2 = moe_block_weights(E, hidden, inter, g["seed"])
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        blk = SparseMoeBlock(hidden, inter, E, k)            # rank / world come from the process group
    finally:
        torch.set_default_dtype(prev)
    ok = (blk.ep_rank, blk.ep_size, blk.num_local, blk.expert_begin) == (rank, world, E // world, rank * E // world)
    blk.gate.weight.weight_loader(blk.gate.weight, gate)
    for name, ex in blk.experts.items():                      # checkpoint order: gate_proj, up_proj, down_proj
        e = int(name)
        ex.gate_up_proj.weight.weight_loader(ex.gate_up_proj.weight, w13[e, :inter], 0)
        ex.gate_up_proj.weight.weight_loader(ex.gate_up_proj.weight, w13[e, inter:], 1)
        ex.down_proj.weight.weight_loader(ex.down_proj.weight, w2[e])
    local = range(blk.expert_begin, blk.expert_begin + blk.num_local)
    full13 = [blk.w13[e - blk.expert_begin] if e in local else None for e in range(E)]
    full2 = [blk.w2[e - blk.expert_begin] if e in local else None for e in range(E)]
    part = moe_ref.sparse_moe_block(g["x"], blk.gate.weight.data, full13, full2, k, experts=set(local)).float()
    dist.all_reduce(part)
    err = (part - g["out"].float()).abs().max().item()
    ok = ok and err <= 0.02 * g["out"].float().abs().max().item()
    out_q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_expert_parallel_layout_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ep_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert results == {0: True, 1: True}


def test_sharded_sampling_key_packing_arithmetic():
    """The int64 key of b200_sample_shard, restated in numpy: MAX over keys == (max score, lowest
    index) and the id decodes back, for negative / zero / infinite scores and exact ties."""
    import numpy as np
    rng = np.random.default_rng(0)

    def key(score: np.float32, idx: int) -> int:
        score = np.float32(0.0) if score == 0.0 else score           # the kernel folds -0.0 into +0.0
        b = int(np.float32(score).view(np.uint32))
        b = (~b & 0xFFFFFFFF) if b & 0x80000000 else (b | 0x80000000)
        return (b << 31) | (0x7FFFFFFF - idx)

    for _ in range(200):
        n = int(rng.integers(2, 50))
        scores = rng.choice(np.array([-np.inf, -3.5, -0.0, 0.0, 1e-30, 2.25, 2.25, 7.0, np.inf], np.float32), n)
        idx = rng.choice(2 ** 20, size=n, replace=False, shuffle=False)
        keys = [key(s, int(i)) for s, i in zip(scores, idx)]
        assert all(0 <= k < 2 ** 63 for k in keys)
        best = max(keys)
        top = scores.max()
        want = min(int(i) for s, i in zip(scores, idx) if s == top)      # float compare: -0.0 == 0.0
        assert 0x7FFFFFFF - (best & 0x7FFFFFFF) == want
    # the product's decoder of all-reduced keys (pure tensor arithmetic, device-agnostic)
    from nanovllm import ops
    ks = torch.tensor([key(np.float32(1.5), 7), key(np.float32(-2.0), 151935), key(np.float32(0.0), 0)])
    assert ops.decode_sample_keys(ks).tolist() == [7, 151935, 0]
    out = torch.zeros(3, dtype=torch.int64)
    assert ops.decode_sample_keys(ks, out=out) is out and out.tolist() == [7, 151935, 0]


def test_decode_step_form_selection():
    """ModelRunner's host-side choices for a decode step: one batch or two micro-batches,
    and the attention grid -- pure arithmetic on the step's context lengths, checked on a stub runner."""
    import types
    from nanovllm.config import Config
    from nanovllm.engine.model_runner import ModelRunner

    def stub(**over):
        cfg = Config("qwen3-0.6b", kvcache_block_size=16, max_num_seqs=256, max_model_len=4096)
        for k in list(over):
            if hasattr(cfg, k):
                setattr(cfg, k, over.pop(k))
        r = types.SimpleNamespace(
            config=cfg, shard_world=1, spec=types.SimpleNamespace(num_experts=0),
            kv_cache=torch.empty(0, dtype=torch.bfloat16), _num_sms=148,
            _kv_token_bytes=2 * 8 * 128 * 2,                 # Qwen3-0.6B: 4 KB of KV per token and layer
            _layer_weight_bytes=31_500_000)
        r.__dict__.update(over)
        for name in ("_overlap_split", "use_split_at", "use_split", "_attn_ctas"):
            setattr(r, name, types.MethodType(getattr(ModelRunner, name), r))
        return r

    r = stub()
    assert r._overlap_split(256) == 128 and r._overlap_split(224) == 112 and r._overlap_split(130) == 72
    assert r._overlap_split(96) == 0                                   # below decode_overlap_min_batch
    assert r.use_split(256, 256 * 1025) and not r.use_split(256, 256 * 900)      # 1.07 GB / 0.94 GB per layer
    assert r.use_split(128, 128 * 2048) and not r.use_split(128, 128 * 1024)     # the same bytes from fewer rows
    assert r._attn_ctas(256) == 104 and r._attn_ctas(128) == 104 and r._attn_ctas(96) == 0
    assert stub(decode_attn_ctas=0)._attn_ctas(256) == 0 and stub(decode_attn_ctas=72)._attn_ctas(8) == 72
    assert stub(decode_overlap_min_batch=0)._overlap_split(256) == 0
    assert stub(shard_world=8)._overlap_split(256) == 0 and stub(shard_world=8)._attn_ctas(256) == 0   # TP ranks
    assert stub(spec=types.SimpleNamespace(num_experts=128))._overlap_split(256) == 0   # experts would stream twice
    assert stub(kv_cache=torch.empty(0, dtype=torch.uint8))._overlap_split(256) == 0    # e4m3 cache
    # weight-heavy layers (Qwen3-32B: ~1 GB per layer): no context length makes the second weight read pay
    assert stub(_layer_weight_bytes=975_000_000)._overlap_split(256) == 0
    # Llama-3.2-1B shapes: 2 KB of KV per token and layer, 122 MB of weights per layer -> a 2.4 GB threshold
    # that 256 full-length (4096-token) contexts = 2.1 GB never reach; longer contexts would
    llama = stub(_kv_token_bytes=2 * 8 * 64 * 2, _layer_weight_bytes=122_000_000)
    assert llama._overlap_split(256) == 0 and llama._attn_ctas(256) == 0
    llama.config.max_model_len = 8192
    assert llama._overlap_split(256) == 128 and not llama.use_split(256, 256 * 4096) and llama.use_split(256, 256 * 5000)
    # tests force the split with a zero threshold
    assert stub(decode_overlap_min_kv_bytes=0, decode_overlap_min_batch=16).use_split(16, 1)


def test_overlap_summary_tool_on_a_synthetic_trace(tmp_path):
    """tools/overlap_summary.py (the script behind profiles/r02b_timeline_decode_*.txt): attention-resident
    time is the UNION of the attention kernels' intervals, whatever the streams."""
    import json
    import subprocess
    ev = []

    def k(name, ts, dur, stream):
        ev.append({"cat": "kernel", "name": name, "ts": ts, "dur": dur, "args": {"stream": stream, "grid": [1, 1, 1]}})
    k("void b200::paged_decode_kernel<bf16, 128, 8>(...)", 0.0, 100.0, 7)
    k("void b200::paged_decode_kernel<bf16, 128, 8>(...)", 60.0, 100.0, 9)        # overlaps the first by 40 us
    k("void b200::moe_umma_gemm_kernel<bf16, false, 6>(...)", 10.0, 20.0, 9)
    k("nvjet_tst_64x32_64x16_2x2_2cta_v_bz_TNT", 160.0, 40.0, 7)
    path = tmp_path / "trace.json"
    path.write_text(json.dumps({"traceEvents": ev + [{"cat": "cpu_op", "name": "x", "ts": 0, "dur": 1}]}))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "overlap_summary.py"), str(path),
                          "--window", "0", "50"], capture_output=True, text=True, check=True).stdout
    assert "4 kernels, step span 200.0 us" in out
    assert "attention resident 160.0 us = 0.800 of the step (sum of its launches 200.0 us)" in out
    assert "| cuBLAS 64x32_64x16_2x2_2cta | 1 | 40.0 | 40.00 |" in out
    assert out.count("  stream ") == 2            # the two kernels starting inside [0, 50)
