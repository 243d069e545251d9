"""Per-rank worker: model, paged KV cache, step metadata, CUDA-graph decode, sampling.

Counterpart of the reference's ``engine/model_runner.py`` with the Ascend pieces replaced:
HCCL -> NCCL (``torch.distributed``), torchair graph compile (:129-154) -> CUDA graphs captured
per batch-size bucket over static input buffers, ``torch.npu`` memory queries -> ``torch.cuda``.
What is kept: the KV layout ``[2, L, nb, bs, Hkv*D]`` bound to every module that has
``k_cache``/``v_cache`` (:218-229), the metadata contract published through ``set_context``
(flat prefill slots, (block, offset) decode slots, -1-padded block tables, padded graph rows
with context_len 0 writing to the dummy last block; :238-366), and the shared-memory RPC that
feeds tensor-parallel workers (:165-193).

Host-side costs are engineered down because a B=256 decode step is only a few ms on a B200:
one packed pinned staging buffer and ONE H2D copy per decode step, block-table rows updated
incrementally (only entries that changed since the previous step), sampling inside the graph,
one small D2H of the sampled ids.
"""
from __future__ import annotations

import os
import pickle
from multiprocessing.shared_memory import SharedMemory

import numpy as np
import torch
import torch.distributed as dist

from ..config import Config, GraphMode
from ..layers.sampler import Sampler
from ..models.models_map import model_dict
from ..utils.context import get_context, reset_context, set_context
from ..utils.loader import device_init, load_model, seeded_init
from .sequence import Sequence

_SHM_PREFIX = "nano_vllm_b200"


def _resolve_dtype(hf_config) -> torch.dtype:
    dt = getattr(hf_config, "torch_dtype", None) or getattr(hf_config, "dtype", None)
    if isinstance(dt, str):
        dt = {"bf16": torch.bfloat16, "bfloat16": torch.bfloat16, "fp16": torch.float16,
              "float16": torch.float16, "half": torch.float16}.get(dt.lower(), getattr(torch, dt, None))
    if dt not in (torch.bfloat16, torch.float16):
        dt = torch.bfloat16 if dt is None or dt == torch.float32 else dt
    return dt


class ModelRunner:

    def __init__(self, config: Config, rank: int, event=None):
        self.config = config
        self.hf_config = config.hf_config
        self.block_size = config.kvcache_block_size
        self.world_size = config.tensor_parallel_size
        self.rank = rank
        self.event = event
        self.spmd = config.distributed_launch == "env"
        self.enforce_eager = config.enforce_eager or config.graph_mode == GraphMode.EAGER.value

        if not torch.cuda.is_available():
            raise RuntimeError("nanovllm (b200): no CUDA device -- the engine has no CPU path")
        if self.world_size > 1 and not dist.is_initialized():
            if self.spmd:
                dist.init_process_group("nccl")
            else:
                dist.init_process_group("nccl", f"tcp://127.0.0.1:{config.hccl_port}",
                                        world_size=self.world_size, rank=rank)
        from ..layers.linear import set_tp_world
        self.shard_world = self.world_size
        if self.world_size == 1 and config.tp_shard_emulation > 1:
            self.shard_world = config.tp_shard_emulation
            set_tp_world(0, self.shard_world, collectives=False)
        else:
            set_tp_world(0, 1) if self.world_size == 1 else set_tp_world(None)
        local = int(os.environ.get("LOCAL_RANK", rank)) if self.spmd else rank
        self.device = torch.device("cuda", local if self.world_size > 1 or self.spmd else
                                   torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        self.dtype = _resolve_dtype(self.hf_config)

        arch = self.hf_config.architectures[0]
        if arch not in model_dict:
            raise ValueError(f"architecture {arch} is outside the hot-path scope (dense Qwen3/Qwen2/Llama)")
        prev = torch.get_default_dtype()
        torch.set_default_dtype(self.dtype)
        try:
            with torch.device(self.device):
                self.model = model_dict[arch](self.hf_config)
        finally:
            torch.set_default_dtype(prev)
        has_ckpt = os.path.isdir(config.model) and any(
            f.endswith(".safetensors") for f in os.listdir(config.model))
        if config.load_weights if config.load_weights is not None else has_ckpt:
            load_model(self.model, config.model)
        else:
            n_params = sum(p.numel() for p in self.model.parameters())
            on_device = config.init_device == "cuda" or (config.init_device == "auto" and n_params > 2e9)
            if on_device:
                device_init(self.model, seed=config.seed, std=config.init_std)
            else:
                seeded_init(self.model, seed=config.seed, norm_jitter=config.norm_jitter, std=config.init_std)
        self.model.eval()
        self.spec = self.model.spec
        self.comm = None
        if self.world_size > 1:
            from ..comm import PeerAllReduce, set_comm
            self.comm = PeerAllReduce(self.device)
            set_comm(self.comm)
        self.sampler = Sampler(seed=config.seed)
        # Tensor parallelism: sample on every rank over its own vocabulary shard and MAX-all-reduce 8
        # bytes per row, instead of gathering [B, vocab] logits on rank 0, sampling there and
        # broadcasting the ids (reference: embed_head.py:62-65 + sampler.py on rank 0). Bit-identical
        # ids (tests/test_kernels_gpu.py, tests/test_tp_gpu.py); B200_SHARDED_SAMPLING=0 restores the
        # gather for A/B timing.
        self.sharded_sampling = self.world_size > 1 and os.environ.get("B200_SHARDED_SAMPLING", "1") != "0"

        self._warmup_prefill()
        self._allocate_kv_cache()
        self._init_decode_buffers()
        self.graphs: dict[int, torch.cuda.CUDAGraph] = {}
        self.split_graphs: dict[int, torch.cuda.CUDAGraph] = {}      # two-micro-batch form of the large buckets
        # bytes one layer's attention reads per cached token, and one layer's weights (split criterion)
        self._kv_token_bytes = 2 * max(1, self.spec.num_kv_heads // self.shard_world) * self.spec.head_dim * \
            self.kv_cache.dtype.itemsize
        self._layer_weight_bytes = sum(p.numel() * p.element_size()
                                       for p in self.model.model.layers[0].parameters())
        self._num_sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        if not self.enforce_eager:
            self._capture_decode_graphs()
        if self.world_size > 1 and not self.spmd:
            self._attach_shared_memory()

    # ------------------------------------------------------------------ TP control plane
    def _attach_shared_memory(self):
        # one segment per engine: the rendezvous port is unique per concurrently running engine on a
        # host, so two tensor-parallel engines never share a control channel
        name = f"{_SHM_PREFIX}_{self.config.hccl_port}"
        # worst case = a prefill batch shipping every token id of max_num_seqs full-length prompts
        # (4 B each) plus their block tables and ~300 B of per-sequence state
        cfg = self.config
        shm_size = max(2 ** 20, cfg.max_num_seqs * (5 * cfg.max_model_len + 1024))
        if self.rank == 0:
            try:
                self.shm = SharedMemory(name=name, create=True, size=shm_size)
            except FileExistsError:                   # stale segment from a crashed run
                stale = SharedMemory(name=name)
                stale.close()
                stale.unlink()
                self.shm = SharedMemory(name=name, create=True, size=shm_size)
            dist.barrier()
        else:
            dist.barrier()
            self.shm = SharedMemory(name=name)
            self.loop()

    def loop(self):
        while True:
            self.event.wait()
            n = int.from_bytes(self.shm.buf[0:4], "little")
            method, *args = pickle.loads(self.shm.buf[4:4 + n])
            self.event.clear()
            self.call(method, *args)
            if method == "exit":
                break

    def call(self, method, *args):
        if self.world_size > 1 and self.rank == 0 and not self.spmd:
            Sequence.wire_decode = method == "run" and not args[1]
            try:
                blob = pickle.dumps([method, *args], protocol=pickle.HIGHEST_PROTOCOL)
            finally:
                Sequence.wire_decode = False
            if len(blob) + 4 > self.shm.size:
                raise RuntimeError(f"step message of {len(blob)} bytes exceeds the {self.shm.size}-byte "
                                   f"control segment (max_num_seqs x max_model_len grew after start-up?)")
            self.shm.buf[0:4] = len(blob).to_bytes(4, "little")
            self.shm.buf[4:4 + len(blob)] = blob
            for ev in self.event:
                ev.set()
        return getattr(self, method)(*args)

    def exit(self):
        self.graphs.clear()
        self.split_graphs.clear()
        if self.comm is not None:
            from ..comm import set_comm
            torch.cuda.synchronize()
            if dist.is_initialized():
                dist.barrier()
            set_comm(None)
            self.comm.close()
            self.comm = None
        if self.world_size > 1 and not self.spmd:
            self.shm.close()
            dist.barrier()
            if self.rank == 0:
                self.shm.unlink()
        torch.cuda.synchronize()
        if self.world_size > 1 and not self.spmd and dist.is_initialized():
            dist.destroy_process_group()

    # ------------------------------------------------------------------ memory
    @torch.inference_mode()
    def _warmup_prefill(self):
        """One max-size prefill without a KV cache: loads cuBLAS kernels and records the peak
        activation footprint before the cache is sized (model_runner.py:195-216)."""
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        cfg = self.config
        n_seqs = max(1, min(cfg.max_num_batched_tokens // cfg.max_model_len, cfg.max_num_seqs))
        seqs = [Sequence([0] * cfg.max_model_len, block_size=self.block_size) for _ in range(n_seqs)]
        ids, pos = self.prepare_prefill(seqs)
        self.model.compute_logits(self.model(ids, pos))
        reset_context()
        torch.cuda.synchronize()
        torch.cuda.empty_cache()

    def _allocate_kv_cache(self):
        cfg, spec = self.config, self.spec
        kv_heads = max(1, spec.num_kv_heads // self.shard_world)    # < 1: kv heads replicated (linear.py)
        # e4m3 caches are kept as raw bytes (uint8); the attention layers dispatch on that dtype
        cache_dtype = self.dtype if cfg.kv_cache_dtype == "auto" else torch.uint8
        block_bytes = 2 * spec.num_layers * self.block_size * kv_heads * spec.head_dim * cache_dtype.itemsize
        if cfg.num_kvcache_blocks <= 0:
            free, total = torch.cuda.mem_get_info()
            stats = torch.cuda.memory_stats()
            used = total - free
            peak, current = stats["allocated_bytes.all.peak"], stats["allocated_bytes.all.current"]
            available = total * cfg.gpu_memory_utilization - used - peak + current
            cfg.num_kvcache_blocks = int(available) // block_bytes
        if self.world_size > 1:
            # every rank must hold every block id the (rank-0) scheduler can hand out
            n = torch.tensor([cfg.num_kvcache_blocks], dtype=torch.int64, device=self.device)
            dist.all_reduce(n, op=dist.ReduceOp.MIN)
            cfg.num_kvcache_blocks = int(n.item())
        assert cfg.num_kvcache_blocks > 1, "not enough memory for the KV cache"
        self.kv_cache = torch.zeros(2, spec.num_layers, cfg.num_kvcache_blocks, self.block_size,
                                    kv_heads * spec.head_dim, dtype=cache_dtype, device=self.device)
        layer = 0
        for module in self.model.modules():
            if hasattr(module, "k_cache") and hasattr(module, "v_cache"):
                module.k_cache = self.kv_cache[0, layer]
                module.v_cache = self.kv_cache[1, layer]
                module.layer_id = layer
                layer += 1

    # ------------------------------------------------------------------ prefill metadata
    def prepare_prefill(self, seqs: list[Sequence]):
        """Packed prompt tokens (model_runner.py:238-290). By default every prompt token is fed
        (the reference recomputes cached prefixes, :247-261); with ``prefix_hit_prefill`` only
        the uncached suffix is fed and attention reads K/V through the block tables
        (attention_ori.py:65-66 semantics); with chunked prefill the scheduler hands every sequence
        a ``prefill_window`` and later windows read the earlier ones the same way."""
        bs = self.block_size
        windows = [s.prefill_window for s in seqs]
        hit = (self.config.prefix_hit_prefill and any(s.num_cached_tokens for s in seqs) or
               any(w is not None and w[0] > 0 for w in windows)) and all(s.block_table for s in seqs)
        ids, pos, slots, cu_q, cu_k = [], [], [], [0], [0]
        max_q = max_k = 0
        for seq, window in zip(seqs, windows):
            # [start, n): the prompt tokens this step computes; keys visible to them: [0, n)
            if window is not None:
                start, n = window            # chunked prefill (scheduler._schedule_prefill_chunks)
            else:
                n = len(seq)
                start = min(seq.num_cached_tokens, n - 1) if hit else 0
            ids.append(np.asarray(seq.token_ids[start:n], dtype=np.int64))
            t = np.arange(start, n, dtype=np.int64)
            pos.append(t)
            cu_q.append(cu_q[-1] + n - start)
            cu_k.append(cu_k[-1] + n)
            max_q, max_k = max(max_q, n - start), max(max_k, n)
            if seq.block_table:
                table = np.asarray(seq.block_table, dtype=np.int64)
                slots.append(table[t // bs] * bs + t % bs)
        ids = np.concatenate(ids) if ids else np.zeros(0, np.int64)
        pos = np.concatenate(pos) if pos else np.zeros(0, np.int64)
        slots = np.concatenate(slots).astype(np.int32) if slots else None
        dev = self.device
        block_tables = self._dense_block_tables(seqs) if any(s.block_table for s in seqs) else None

        def as_dev(data, dt):
            host = torch.from_numpy(data) if isinstance(data, np.ndarray) else torch.tensor(data, dtype=dt)
            return host.to(dt).pin_memory().to(dev, non_blocking=True)
        set_context(True, cu_seqlens_q=as_dev(cu_q, torch.int32), cu_seqlens_k=as_dev(cu_k, torch.int32),
                    max_seqlen_q=max_q, max_seqlen_k=max_k,
                    slot_mapping=as_dev(slots, torch.int32) if slots is not None else None,
                    context_lens=None, block_tables=block_tables, is_enforce_eager=True,
                    real_bs=len(seqs), block_size=bs, paged_prefill=hit)
        return as_dev(ids, torch.int64), as_dev(pos, torch.int64)

    def _dense_block_tables(self, seqs):
        width = max(len(s.block_table) for s in seqs)
        rows = [s.block_table + [-1] * (width - len(s.block_table)) for s in seqs]
        return torch.tensor(rows, dtype=torch.int32, pin_memory=True).to(self.device, non_blocking=True)

    # ------------------------------------------------------------------ decode metadata
    def _init_decode_buffers(self):
        """One pinned host buffer + one device buffer holding every decode input:
        [ids i64 | positions i64 | slots i32x2 | context_lens i32 | temperatures f32 | block_tables i32]."""
        cfg = self.config
        B = cfg.max_num_seqs
        self.max_blocks = (cfg.max_model_len + self.block_size - 1) // self.block_size
        sizes = [B * 8, B * 8, B * 8, B * 4, B * 4, B * self.max_blocks * 4]
        offs = np.concatenate([[0], np.cumsum([(s + 255) // 256 * 256 for s in sizes])])
        self._stage_host = torch.zeros(int(offs[-1]), dtype=torch.uint8, pin_memory=True)
        self._stage_dev = torch.zeros(int(offs[-1]), dtype=torch.uint8, device=self.device)

        def views(buf):
            seg = lambda i, dt, shape: buf[int(offs[i]):int(offs[i]) + sizes[i]].view(dt).view(shape)
            return (seg(0, torch.int64, (B,)), seg(1, torch.int64, (B,)), seg(2, torch.int32, (B, 2)),
                    seg(3, torch.int32, (B,)), seg(4, torch.float32, (B,)),
                    seg(5, torch.int32, (B, self.max_blocks)))
        self.d_ids, self.d_pos, self.d_slots, self.d_ctx, self.d_temp, self.d_tables = views(self._stage_dev)
        h = views(self._stage_host)
        self.h_ids, self.h_pos, self.h_slots, self.h_ctx, self.h_temp, self.h_tables = [t.numpy() for t in h]
        self.h_tables.fill(-1)
        self._stage_dev.copy_(self._stage_host)
        self._head_bytes = int(offs[5])          # everything before the block tables
        self._row_owner = [-1] * B               # seq_id whose table sits in row i
        self._row_len = [0] * B                  # how many entries of that table are staged
        self.d_out = torch.zeros(B, dtype=torch.int64, device=self.device)
        self.h_out = torch.zeros(B, dtype=torch.int64, pin_memory=True)
        self.d_rng = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.d_keys = torch.zeros(B, dtype=torch.int64, device=self.device)     # sharded sampling keys
        self.h2d_bytes = self.d2h_bytes = 0        # decode-step traffic counters (bench.py e2e)
        self.launches_per_graph: dict = {}             # bucket (or (bucket, "split")) -> launches per replay
        # two-micro-batch decode (_forward_overlapped): second stream + the joined hidden states
        self._side_stream = torch.cuda.Stream(self.device)
        self.d_hidden = torch.zeros(B, self.spec.hidden_size, dtype=self.dtype, device=self.device)

    def _stage_decode(self, seqs: list[Sequence], padded: int) -> int:
        """Fill the pinned staging buffer for a decode step and upload it with one copy
        (model_runner.py:292-366). Rows >= len(seqs) are padding: token 0, position 0,
        context_len 0, slot = (last block, 0), table -1. Returns the step's total context tokens."""
        n = len(seqs)
        bs = self.block_size
        self.h_ids[:n] = [s.last_token for s in seqs]
        lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=n)
        self.h_pos[:n] = lens - 1
        self.h_ctx[:n] = lens
        self.h_slots[:n, 0] = [s.block_table[-1] for s in seqs]
        self.h_slots[:n, 1] = (lens - 1) % bs
        self.h_temp[:n] = [0.0 if s.greedy else s.temperature for s in seqs]
        if padded > n:
            self.h_ids[n:padded] = 0
            self.h_pos[n:padded] = 0
            self.h_ctx[n:padded] = 0
            self.h_slots[n:padded, 0] = self.config.num_kvcache_blocks - 1
            self.h_slots[n:padded, 1] = 0
            self.h_temp[n:padded] = 1.0
        # block tables: touch only what changed since the previous step
        dirty_hi = 0
        owner, rlen, tables = self._row_owner, self._row_len, self.h_tables
        for i, s in enumerate(seqs):
            t = s.block_table
            if owner[i] == s.seq_id and rlen[i] <= len(t):
                if rlen[i] < len(t):
                    tables[i, rlen[i]:len(t)] = t[rlen[i]:]
                    rlen[i] = len(t)
                    dirty_hi = i + 1
            else:
                tables[i, :len(t)] = t
                if rlen[i] > len(t):
                    tables[i, len(t):rlen[i]] = -1
                owner[i], rlen[i] = s.seq_id, len(t)
                dirty_hi = i + 1
        for i in range(n, padded):
            if owner[i] != -1:
                tables[i, :rlen[i]] = -1
                owner[i], rlen[i] = -1, 0
                dirty_hi = i + 1
        nbytes = self._head_bytes + dirty_hi * self.max_blocks * 4
        self._stage_dev[:nbytes].copy_(self._stage_host[:nbytes], non_blocking=True)
        self.h2d_bytes += nbytes
        return int(lens.sum())

    def _attn_ctas(self, bs: int) -> int:
        """Persistent grid of the paged decode kernel for a `bs`-row step (Config.decode_attn_ctas)."""
        n = self.config.decode_attn_ctas
        if n >= 0:
            return n
        if bs < 128 or self.shard_world > 1 or self._kv_token_bytes < 4096:
            return 0                    # small launches are latency-bound: every SM helps
        return max(8, int(self._num_sms * 0.7 + 4) // 8 * 8)

    def _set_decode_context(self, bs: int, real_bs: int, lo: int = 0, own_gemm: bool = False, ctas: int | None = None):
        """Publish rows [lo, bs) of the static decode buffers as the step context."""
        set_context(False, slot_mapping=self.d_slots[lo:bs], context_lens=self.d_ctx[lo:bs],
                    block_tables=self.d_tables[lo:bs], is_enforce_eager=self.enforce_eager,
                    real_bs=real_bs, block_size=self.block_size,
                    decode_attn_ctas=self._attn_ctas(bs) if ctas is None else ctas, own_gemm=own_gemm)

    # ------------------------------------------------------------------ the decode step
    def _overlap_split(self, bs: int) -> int:
        """Rows of the first micro-batch if a `bs`-row decode step MAY run as two micro-batches, else 0:
        single GPU (the chain has no collective), dense MLPs (a split would stream every expert twice)."""
        least = self.config.decode_overlap_min_batch
        if least <= 0 or bs < max(least, 16) or self.shard_world > 1 or self.spec.num_experts > 0:
            return 0
        if self.kv_cache.dtype == torch.uint8:
            return 0                   # e4m3 cache: half the KV bytes per step; the split form is not validated there
        if not self.use_split_at(bs * self.config.max_model_len):
            return 0                   # not even full-length contexts would reach the threshold
        return (bs // 2 + 7) // 8 * 8

    def use_split_at(self, kv_tokens: int) -> bool:
        if self.config.decode_overlap_min_kv_bytes <= 0:       # tests: every eligible step splits
            return True
        return kv_tokens * self._kv_token_bytes >= max(self.config.decode_overlap_min_kv_bytes,
                                                       20 * self._layer_weight_bytes)

    def use_split(self, bs: int, kv_tokens: int) -> bool:
        """Does a `bs`-row step over `kv_tokens` cached tokens run as two micro-batches? Only when the
        attention of one half is long enough to cover the other half's chain (which runs ~2.5x slower
        next to a kernel that saturates HBM: measured ~100 us against ~37 us per layer for Qwen3-0.6B),
        and the second read of the layer weights is small against the KV stream."""
        return bool(self._overlap_split(bs)) and self.use_split_at(kv_tokens)

    def decode_graph(self, bs: int, kv_tokens: int | None = None):
        """(graph, launches of this library per replay) that run() replays for a `bs`-row bucket over
        `kv_tokens` cached tokens (default: the context lengths staged by the last step)."""
        if kv_tokens is None:
            kv_tokens = int(self.h_ctx[:bs].sum())
        if bs in self.split_graphs and self.use_split(bs, kv_tokens):
            return self.split_graphs[bs], self.launches_per_graph[(bs, "split")]
        return self.graphs[bs], self.launches_per_graph[bs]

    def _forward_overlapped(self, bs: int, split: int) -> torch.Tensor:
        """The layer stack over rows [0, split) on the current stream and over rows [split, bs) on a
        second stream. Rows of a decode step are independent sequences, so the two halves share nothing
        but the weights; what the split buys is overlap of two different bottlenecks -- the paged
        attention of one half streams KV at the HBM roofline while the other half runs its GEMM / norm /
        RoPE chain, ~8 launches of a few microseconds each that are latency-bound and would otherwise
        leave HBM idle. Under CUDA-graph capture the second stream forks from and joins
        the capturing stream, so the step is still ONE graph with two parallel branches. Every kernel is
        row-wise (GEMMs included), so the result equals the one-batch step up to cuBLAS' choice of
        algorithm for M = bs / 2."""
        real_bs = get_context().real_bs
        own, ctas = self.config.decode_overlap_own_gemm, self._attn_ctas(bs)
        main, side = torch.cuda.current_stream(), self._side_stream
        out = self.d_hidden[:bs]
        side.wait_stream(main)
        self._set_decode_context(split, real_bs, 0, own, ctas)
        out[:split].copy_(self.model(self.d_ids[:split], self.d_pos[:split]))
        with torch.cuda.stream(side):
            self._set_decode_context(bs, real_bs, split, own, ctas)
            out[split:].copy_(self.model(self.d_ids[split:bs], self.d_pos[split:bs]))
        main.wait_stream(side)
        self._set_decode_context(bs, real_bs)
        return out

    def _decode_step(self, bs: int, two_halves: bool = False):
        """Forward + logits + sampling for `bs` rows of the static buffers (graph body)."""
        split = self._overlap_split(bs) if two_halves else 0
        hidden = self._forward_overlapped(bs, split) if split else self.model(self.d_ids[:bs], self.d_pos[:bs])
        if self.sharded_sampling:
            self.d_rng.add_(1)
            self._sample_sharded(hidden, self.d_temp[:bs], self.d_keys[:bs], self.d_out[:bs])
            return
        logits = self.model.compute_logits(hidden)
        self.d_rng.add_(1)
        if self.rank == 0:
            self.sampler(logits, self.d_temp[:bs], out=self.d_out[:bs], step_offset=self.d_rng)
        if self.world_size > 1:
            dist.broadcast(self.d_out[:bs], 0)

    def _sample_sharded(self, hidden, temps, keys, out):
        head = self.model.lm_head
        self.sampler.shard_keys(head.local_logits(hidden), temps, head.vocab_start_idx, out=keys,
                                step_offset=self.d_rng)
        dist.all_reduce(keys, op=dist.ReduceOp.MAX)
        from .. import ops
        ops.decode_sample_keys(keys, out=out)

    def _buckets(self) -> list[int]:
        top = self.config.max_num_seqs
        ladder = self.config.graph_batch_sizes or (1, 2, 4, 8, 16, 32, 64, 96, 128, 160, 192, 224, 256,
                                                   320, 384, 448, 512)
        return sorted({b for b in ladder if 0 < b < top}) + [top]

    @torch.inference_mode()
    def _capture_decode_graphs(self):
        from .. import ops
        pool = None
        for bs in reversed(self._buckets()):
            for two_halves in ([False, True] if self._overlap_split(bs) else [False]):
                self._set_decode_context(bs, bs)
                self._decode_step(bs, two_halves)      # warm-up (allocator, cuBLAS heuristics)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                before = ops.launch_count()
                with torch.cuda.graph(graph, pool=pool):
                    self._decode_step(bs, two_halves)
                self.launches_per_graph[(bs, "split") if two_halves else bs] = ops.launch_count() - before
                pool = pool or graph.pool()
                (self.split_graphs if two_halves else self.graphs)[bs] = graph
        reset_context()
        torch.cuda.synchronize()
        self.kv_cache[:, :, -1].zero_()                # the dummy block soaked up warm-up writes

    # ------------------------------------------------------------------ run
    @torch.inference_mode()
    def run(self, seqs: list[Sequence], is_prefill: bool) -> list[int] | None:
        n = len(seqs)
        if is_prefill:
            ids, pos = self.prepare_prefill(seqs)
            out = self.d_out[:n] if n <= self.d_out.numel() else torch.empty(
                n, dtype=torch.int64, device=self.device)
            if self.sharded_sampling:
                hidden = self.model(ids, pos)
                temps = torch.tensor([0.0 if s.greedy else s.temperature for s in seqs],
                                     dtype=torch.float32, pin_memory=True).to(self.device, non_blocking=True)
                self.d_rng.add_(1)
                self._sample_sharded(hidden, temps, torch.empty_like(out), out)
                reset_context()
                return out.tolist()
            logits = self.model.compute_logits(self.model(ids, pos))
            if self.rank == 0:
                temps = torch.tensor([0.0 if s.greedy else s.temperature for s in seqs],
                                     dtype=torch.float32, pin_memory=True).to(self.device, non_blocking=True)
                self.d_rng.add_(1)
                self.sampler(logits, temps, out=out, step_offset=self.d_rng)
            if self.world_size > 1:
                dist.broadcast(out, 0)
            reset_context()
            return out.tolist()
        use_graph = not self.enforce_eager and n <= self.config.max_num_seqs
        bs = next(b for b in self._buckets() if b >= n) if use_graph else n
        kv_tokens = self._stage_decode(seqs, bs)
        if use_graph:
            self.decode_graph(bs, kv_tokens)[0].replay()
        else:
            self._set_decode_context(bs, n)
            self._decode_step(bs, self.use_split(bs, kv_tokens))
            reset_context()
        self.h_out[:n].copy_(self.d_out[:n], non_blocking=True)
        self.d2h_bytes += n * 8
        torch.cuda.current_stream().synchronize()
        return self.h_out[:n].tolist()
