"""Engine configuration (reference: config.py:14-60). Field names are kept; ``device`` is
"cuda", ``hccl_port`` is the TCP rendezvous port of the NCCL process group, ``graph_mode``
selects eager vs CUDA-graph decode. ``model`` may be a checkpoint directory (HF layout), or
the name of a built-in architecture preset (random-init weights, for synthetic benchmarks
where no checkpoint can be downloaded)."""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from enum import Enum


class GraphMode(Enum):
    EAGER = "eager"
    CUDA_GRAPH = "cuda-graph"
    # accepted aliases of the reference's Ascend modes: both mean "replay a captured step"
    MAX_AUTOTUNE = "max-autotune"
    REDUCE_OVERHEAD = "reduce-overhead"


def preset_hf_config(name: str):
    """Architecture presets built offline from the public model cards."""
    from transformers import LlamaConfig, Qwen3Config, Qwen3MoeConfig
    key = name.lower().rstrip("/").split("/")[-1]
    if key == "qwen3-0.6b":
        c = Qwen3Config(vocab_size=151936, hidden_size=1024, num_hidden_layers=28,
                        num_attention_heads=16, num_key_value_heads=8, head_dim=128,
                        intermediate_size=3072, rms_norm_eps=1e-6, max_position_embeddings=40960,
                        tie_word_embeddings=True, attention_bias=False, rope_theta=1000000.0)
    elif key == "qwen3-32b":
        c = Qwen3Config(vocab_size=151936, hidden_size=5120, num_hidden_layers=64,
                        num_attention_heads=64, num_key_value_heads=8, head_dim=128,
                        intermediate_size=25600, rms_norm_eps=1e-6, max_position_embeddings=40960,
                        tie_word_embeddings=False, attention_bias=False, rope_theta=1000000.0)
    elif key == "llama-3.2-1b":
        c = LlamaConfig(vocab_size=128256, hidden_size=2048, num_hidden_layers=16,
                        num_attention_heads=32, num_key_value_heads=8, head_dim=64,
                        intermediate_size=8192, rms_norm_eps=1e-5, max_position_embeddings=131072,
                        tie_word_embeddings=True, rope_theta=500000.0)
    elif key == "tiny-qwen3":
        c = Qwen3Config(vocab_size=512, hidden_size=128, num_hidden_layers=2,
                        num_attention_heads=4, num_key_value_heads=2, head_dim=64,
                        intermediate_size=256, rms_norm_eps=1e-6, max_position_embeddings=512,
                        tie_word_embeddings=True, attention_bias=False, rope_theta=1000000.0)
    elif key == "tiny-llama":
        c = LlamaConfig(vocab_size=512, hidden_size=256, num_hidden_layers=2,
                        num_attention_heads=4, num_key_value_heads=2, head_dim=64,
                        intermediate_size=512, rms_norm_eps=1e-5, max_position_embeddings=512,
                        tie_word_embeddings=True, rope_theta=10000.0)
    elif key == "qwen3-30b-a3b":
        c = Qwen3MoeConfig(vocab_size=151936, hidden_size=2048, num_hidden_layers=48,
                           num_attention_heads=32, num_key_value_heads=4, head_dim=128,
                           intermediate_size=6144, moe_intermediate_size=768, num_experts=128,
                           num_experts_per_tok=8, decoder_sparse_step=1, mlp_only_layers=[],
                           norm_topk_prob=True, rms_norm_eps=1e-6, max_position_embeddings=40960,
                           tie_word_embeddings=False, attention_bias=False, rope_theta=1000000.0)
    elif key == "tiny-qwen3-moe":
        c = Qwen3MoeConfig(vocab_size=512, hidden_size=128, num_hidden_layers=2,
                           num_attention_heads=4, num_key_value_heads=2, head_dim=64,
                           intermediate_size=256, moe_intermediate_size=64, num_experts=8,
                           num_experts_per_tok=2, decoder_sparse_step=1, mlp_only_layers=[],
                           rms_norm_eps=1e-6, max_position_embeddings=512,
                           tie_word_embeddings=False, attention_bias=False, rope_theta=1000000.0)
    else:
        return None
    c.architectures = ["LlamaForCausalLM" if "llama" in key else
                       "Qwen3MoeForCausalLM" if isinstance(c, Qwen3MoeConfig) else "Qwen3ForCausalLM"]
    c.dtype = "bfloat16"
    c.eos_token_id = -1 if key.startswith("tiny-") else c.eos_token_id
    return c


@dataclass
class Config:
    model: str
    max_num_batched_tokens: int = 16384
    max_num_seqs: int = 256
    max_model_len: int = 4096
    gpu_memory_utilization: float = 0.7
    tensor_parallel_size: int = 1
    enforce_eager: bool = False
    hf_config: object | None = None
    eos: int = -1
    kvcache_block_size: int = 256
    num_kvcache_blocks: int = -1
    use_graph_cache: bool = False
    hccl_port: int = 28000
    graph_mode: str = GraphMode.CUDA_GRAPH.value
    is_multimodal: bool = False
    trust_remote_code: bool = False
    # additions (all optional; defaults keep the reference behaviour)
    seed: int = 0                      # random-init / sampler seed
    norm_jitter: float = 0.0           # random-init only
    init_std: float = 0.02             # random-init only: std of the matrices
    init_device: str = "auto"          # random-init: "cpu" (name-seeded, == oracle), "cuda", or "auto"
                                       # (cuda above 2e9 parameters)
    load_weights: bool | None = None   # None: load iff *.safetensors exist under `model`
    distributed_launch: str = "spawn"  # "spawn": reference style (rank 0 spawns workers);
                                       # "env": every rank is already running (torchrun/SPMD)
    prefix_hit_prefill: bool = False   # feed only uncached prompt tokens (attention_ori semantics)
    chunked_prefill: bool = False      # split prompts over several prefill steps when the token budget
                                       # (max_num_batched_tokens) runs out; later chunks attend to the
                                       # earlier ones through the paged cache
    graph_batch_sizes: tuple | None = None   # decode batch sizes to capture CUDA graphs for (None: the
                                       # default ladder 1, 2, 4, ..., max_num_seqs; max_num_seqs is always added)
    tp_shard_emulation: int = 0        # profiling aid: N > 1 builds rank 0's shard of an N-way tensor-parallel
                                       # model in ONE process, collectives skipped (per-rank kernel timeline of a
                                       # TP=N step on one GPU; outputs are meaningless). tensor_parallel_size stays 1
    kv_cache_dtype: str = "auto"       # "auto": the model dtype (reference behaviour); "fp8" /
                                       # "fp8_e4m3": e4m3 bytes, per-tensor scale 1.0
    # Two-micro-batch decode (engine/model_runner.py::_forward_overlapped): a decode step whose attention
    # streams at least decode_overlap_min_kv_bytes of KV per layer (and at least 20x the layer's weight
    # bytes) runs as TWO half batches on two streams inside one CUDA graph, so that the HBM-bound attention
    # of one half overlaps the latency-bound GEMM / norm / RoPE chain of the other (measured, Qwen3-0.6B,
    # 256 x 1027 tokens: 5.31 ms/step against 5.75-5.87). Below the threshold the one-batch graph of the same
    # bucket is replayed. Single GPU, dense models.
    decode_overlap_min_batch: int = 128     # buckets below this never split; 0 disables the feature
    decode_overlap_min_kv_bytes: float = 1.0e9
    decode_overlap_own_gemm: bool = True    # the split step's GEMMs on the library's single-CTA tcgen05 kernel
                                            # (cuBLAS' clustered 2-CTA kernels cannot be placed next to a
                                            # resident attention kernel: measured, no overlap at all)
    decode_attn_ctas: int = -1         # persistent grid of the paged decode kernel in decode steps: 0 = one CTA
                                       # per SM; -1 = auto: 0.7 x #SMs (a multiple of 8) from 128 rows on -- the
                                       # HBM-saturated regime, where 104 CTAs stream 6.7 TB/s and 148 CTAs 6.4
    device = "cuda"

    def __post_init__(self):
        assert self.kvcache_block_size % 16 == 0
        assert 1 <= self.tensor_parallel_size <= 8
        assert self.kv_cache_dtype in ("auto", "fp8", "fp8_e4m3"), self.kv_cache_dtype
        if self.hf_config is None:
            if os.path.isdir(self.model):
                from transformers import AutoConfig
                self.hf_config = AutoConfig.from_pretrained(self.model,
                                                            trust_remote_code=self.trust_remote_code)
            else:
                self.hf_config = preset_hf_config(self.model)
                assert self.hf_config is not None, (
                    f"{self.model!r} is neither a checkpoint directory nor a known preset")
        text_config = getattr(self.hf_config, "text_config", self.hf_config)
        max_pos = getattr(text_config, "max_position_embeddings", None)
        if max_pos is not None:
            self.max_model_len = min(self.max_model_len, max_pos)
        eos = getattr(text_config, "eos_token_id", None)
        if isinstance(eos, (list, tuple)):
            eos = eos[0] if eos else None
        if eos is not None:
            self.eos = eos
        # without chunking a prompt must fit one prefill step (reference config.py)
        assert self.chunked_prefill or self.max_num_batched_tokens >= self.max_model_len
