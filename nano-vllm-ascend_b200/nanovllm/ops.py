"""Tensor-level wrappers over the C ABI: torch owns memory and streams, CUDA does the work.

Each function validates what the C side cannot know (device, dtype, contiguity of the inner
dimensions), then passes raw pointers + the current stream. All of them enqueue only; they
are safe inside CUDA-graph capture.
"""
from __future__ import annotations

import os

import torch

from . import _C

_DT = {torch.bfloat16: _C.BF16, torch.float16: _C.F16}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"b200attn kernels take bf16/fp16 tensors, got {t.dtype}") from None


def _need_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "b200attn: the hot path runs on CUDA only (got a CPU tensor); there is no "
                "CPU fallback -- use oracle/ for CPU reference results")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _cache_dims(cache: torch.Tensor, row_elems: int) -> tuple[int, int]:
    """(num_blocks, block_size) of a cache given as [nb, bs, Hkv*D] or [nb, bs, Hkv, D]."""
    if cache.dim() not in (3, 4) or not cache.is_contiguous():
        raise ValueError("KV cache must be a contiguous [nb, bs, Hkv*D] (or [nb, bs, Hkv, D]) tensor")
    if cache[0, 0].numel() != row_elems:
        raise ValueError(f"KV cache row has {cache[0, 0].numel()} elements, expected {row_elems}")
    return cache.shape[0], cache.shape[1]


def launch_count() -> int:
    return _C.load().b200_launch_count()


def reset_launch_count() -> None:
    _C.load().b200_reset_launch_count()


def reshape_and_cache(key, value, k_cache, v_cache, slot_mapping):
    """K1/K2. key/value [T, Hkv, D] (token-strided views allowed), slot_mapping [T] flat or
    [T, 2] (block, offset); int32."""
    _need_cuda(key, value, k_cache, v_cache, slot_mapping)
    T, Hkv, D = key.shape
    if value.shape != key.shape:
        raise ValueError("key/value shape mismatch")
    if key.stride(2) != 1 or key.stride(1) != D or value.stride(2) != 1 or value.stride(1) != D:
        raise ValueError("key/value must be contiguous in (heads, head_dim)")
    if slot_mapping.dtype != torch.int32 or not slot_mapping.is_contiguous():
        raise TypeError("slot_mapping must be a contiguous int32 tensor")
    nb, bs = _cache_dims(k_cache, Hkv * D)
    lib = _C.load()
    two_d = slot_mapping.dim() == 2
    if slot_mapping.shape[0] != T or (two_d and slot_mapping.shape[1] != 2):
        raise ValueError(f"slot_mapping shape {tuple(slot_mapping.shape)} does not match {T} tokens")
    fn = lib.b200_reshape_and_cache_2d if two_d else lib.b200_reshape_and_cache
    _C.check(fn(key.data_ptr(), value.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(),
                slot_mapping.data_ptr(), T, Hkv, D, key.stride(0) if T > 1 else Hkv * D,
                value.stride(0) if T > 1 else Hkv * D, nb, bs, _dt(key), _stream()),
             "b200_reshape_and_cache")


def _is_fp8_cache(cache: torch.Tensor) -> bool:
    return cache.dtype in (torch.uint8, torch.float8_e4m3fn)


def reshape_and_cache_fp8(key, value, k_cache, v_cache, slot_mapping, k_scale=1.0, v_scale=1.0):
    """K1/K2 into an e4m3 cache (uint8 / float8_e4m3fn tensors [nb, bs, Hkv*D]):
    cache <- e4m3(x * (1/scale)), saturating. key/value [T, Hkv, D] bf16/fp16."""
    _need_cuda(key, value, k_cache, v_cache, slot_mapping)
    T, Hkv, D = key.shape
    if value.shape != key.shape:
        raise ValueError("key/value shape mismatch")
    if key.stride(2) != 1 or key.stride(1) != D or value.stride(2) != 1 or value.stride(1) != D:
        raise ValueError("key/value must be contiguous in (heads, head_dim)")
    if slot_mapping.dtype != torch.int32 or not slot_mapping.is_contiguous():
        raise TypeError("slot_mapping must be a contiguous int32 tensor")
    if not (_is_fp8_cache(k_cache) and _is_fp8_cache(v_cache)):
        raise TypeError("fp8 KV caches must be uint8 or float8_e4m3fn tensors")
    nb, bs = _cache_dims(k_cache, Hkv * D)
    two_d = slot_mapping.dim() == 2
    if slot_mapping.shape[0] != T or (two_d and slot_mapping.shape[1] != 2):
        raise ValueError(f"slot_mapping shape {tuple(slot_mapping.shape)} does not match {T} tokens")
    _C.check(_C.load().b200_reshape_and_cache_fp8(
        key.data_ptr(), value.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(),
        slot_mapping.data_ptr(), 1 if two_d else 0, T, Hkv, D,
        key.stride(0) if T > 1 else Hkv * D, value.stride(0) if T > 1 else Hkv * D, nb, bs,
        float(k_scale), float(v_scale), _dt(key), _stream()), "b200_reshape_and_cache_fp8")


_decode_ws: dict = {}


def decode_workspace(batch, num_heads, num_kv_heads, head_dim, device) -> torch.Tensor:
    need = _C.load().b200_paged_decode_attn_workspace_size(batch, num_heads, num_kv_heads, head_dim)
    if need == 0:
        raise _C.B200Error("b200_paged_decode_attn: unsupported head configuration "
                           f"(H={num_heads}, Hkv={num_kv_heads}, D={head_dim})")
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _decode_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(need, dtype=torch.uint8, device=device)   # arrival counters start at 0
        _decode_ws[key] = ws
    return ws


def paged_decode_attn(q, k_cache, v_cache, block_tables, context_lens, scale, num_kv_heads,
                      out=None, workspace=None, num_ctas=0):
    """K6. q [B, H, D] (token-strided view allowed) -> out [B, H*D]."""
    _need_cuda(q, k_cache, v_cache, block_tables, context_lens)
    B, H, D = q.shape
    if q.stride(2) != 1 or q.stride(1) != D:
        raise ValueError("q must be contiguous in (heads, head_dim)")
    if block_tables.dtype != torch.int32 or context_lens.dtype != torch.int32:
        raise TypeError("block_tables/context_lens must be int32")
    if block_tables.dim() != 2 or block_tables.stride(1) != 1 or not context_lens.is_contiguous():
        raise ValueError("block_tables must be [B, max_blocks] row-major, context_lens contiguous")
    nb, bs = _cache_dims(k_cache, num_kv_heads * D)
    if out is None:
        out = torch.empty(B, H * D, dtype=q.dtype, device=q.device)
    if workspace is None:
        workspace = decode_workspace(B, H, num_kv_heads, D, q.device)
    _C.check(_C.load().b200_paged_decode_attn(
        out.data_ptr(), q.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(),
        block_tables.data_ptr(), context_lens.data_ptr(), B, H, num_kv_heads, D,
        q.stride(0) if B > 1 else H * D, out.stride(0) if B > 1 else H * D, nb, bs,
        block_tables.stride(0), float(scale), workspace.data_ptr(), workspace.numel(),
        int(num_ctas), _dt(q), _stream()), "b200_paged_decode_attn")
    return out


def paged_decode_attn_fp8(q, k_cache, v_cache, block_tables, context_lens, scale, num_kv_heads,
                          k_scale=1.0, v_scale=1.0, out=None, workspace=None, num_ctas=0):
    """K6 over an e4m3 cache. q [B, H, D] bf16/fp16 -> out [B, H*D] in q's dtype."""
    _need_cuda(q, k_cache, v_cache, block_tables, context_lens)
    B, H, D = q.shape
    if q.stride(2) != 1 or q.stride(1) != D:
        raise ValueError("q must be contiguous in (heads, head_dim)")
    if block_tables.dtype != torch.int32 or context_lens.dtype != torch.int32:
        raise TypeError("block_tables/context_lens must be int32")
    if block_tables.dim() != 2 or block_tables.stride(1) != 1 or not context_lens.is_contiguous():
        raise ValueError("block_tables must be [B, max_blocks] row-major, context_lens contiguous")
    if not (_is_fp8_cache(k_cache) and _is_fp8_cache(v_cache)):
        raise TypeError("fp8 KV caches must be uint8 or float8_e4m3fn tensors")
    nb, bs = _cache_dims(k_cache, num_kv_heads * D)
    if out is None:
        out = torch.empty(B, H * D, dtype=q.dtype, device=q.device)
    if workspace is None:
        workspace = decode_workspace(B, H, num_kv_heads, D, q.device)
    _C.check(_C.load().b200_paged_decode_attn_fp8(
        out.data_ptr(), q.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(),
        block_tables.data_ptr(), context_lens.data_ptr(), B, H, num_kv_heads, D,
        q.stride(0) if B > 1 else H * D, out.stride(0) if B > 1 else H * D, nb, bs,
        block_tables.stride(0), float(scale), float(k_scale), float(v_scale), workspace.data_ptr(),
        workspace.numel(), int(num_ctas), _dt(q), _stream()), "b200_paged_decode_attn_fp8")
    return out


def gather_dequant_fp8(k_cache, v_cache, block_tables, cu_seqlens_k, max_seqlen_k, num_kv_heads, head_dim,
                       k_scale=1.0, v_scale=1.0, dtype=torch.bfloat16, total_k=None):
    """Paged e4m3 caches -> packed [total_k, Hkv, D] K and V in `dtype` (prefix-hit / chunked prefill over
    an fp8 cache). total_k: number of packed rows (= cu_seqlens_k[-1]); when None it is bounded by
    num_seqs * max_seqlen_k so that no device value has to be read back."""
    _need_cuda(k_cache, v_cache, block_tables, cu_seqlens_k)
    if not (_is_fp8_cache(k_cache) and _is_fp8_cache(v_cache)):
        raise TypeError("fp8 KV caches must be uint8 or float8_e4m3fn tensors")
    if block_tables.dtype != torch.int32 or block_tables.stride(1) != 1 or cu_seqlens_k.dtype != torch.int32:
        raise TypeError("block_tables must be row-major int32, cu_seqlens_k int32")
    num_seqs = cu_seqlens_k.numel() - 1
    nb, bs = _cache_dims(k_cache, num_kv_heads * head_dim)
    rows = int(total_k) if total_k is not None else num_seqs * int(max_seqlen_k)
    # zeros, not empty: rows past the last sequence are inside the prefill kernel's TMA bounds, and a
    # masked key's probability 0 times a stale NaN pattern in V would poison the output row
    k = torch.zeros(rows, num_kv_heads, head_dim, dtype=dtype, device=k_cache.device)
    v = torch.zeros_like(k)
    _C.check(_C.load().b200_gather_dequant_fp8(
        k.data_ptr(), v.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), block_tables.data_ptr(),
        cu_seqlens_k.data_ptr(), num_seqs, int(max_seqlen_k), block_tables.stride(0), nb, bs, num_kv_heads,
        head_dim, float(k_scale), float(v_scale), _dt(k), _stream()), "b200_gather_dequant_fp8")
    return k, v


_prefill_ws: dict = {}


def prefill_workspace(device, head_dim: int) -> torch.Tensor:
    """The split-KV workspace of the prefill kernel (b200_prefill_workspace_size bytes, a few MB),
    one per device and stream, allocated on first use -- outside CUDA-graph capture; prefill runs eagerly."""
    key = (device, head_dim, torch.cuda.current_stream(device).cuda_stream)   # concurrent launches must not share it
    ws = _prefill_ws.get(key)
    if ws is None:
        n = int(_C.load().b200_prefill_workspace_size(int(head_dim)))
        ws = _prefill_ws[key] = torch.empty(n, dtype=torch.uint8, device=device)
    return ws


def prefill_varlen_attn(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, scale,
                        num_kv_heads, block_tables=None, out=None, workspace="auto"):
    """K4/K5. q [Tq, H, D]; packed mode: k/v [Tk, Hkv, D]; paged mode (block_tables given):
    k/v are the caches [nb, bs, Hkv*D]. Returns out [Tq, H*D]. workspace: "auto" (the per-device
    split-KV workspace), None (never split the keys across CTAs) or a uint8 CUDA tensor."""
    _need_cuda(q, k, v, cu_seqlens_q, cu_seqlens_k, block_tables)
    Tq, H, D = q.shape
    if q.stride(2) != 1 or q.stride(1) != D:
        raise ValueError("q must be contiguous in (heads, head_dim)")
    if cu_seqlens_q.dtype != torch.int32 or cu_seqlens_k.dtype != torch.int32:
        raise TypeError("cu_seqlens must be int32")
    num_seqs = cu_seqlens_q.numel() - 1
    if out is None:
        out = torch.empty(Tq, H * D, dtype=q.dtype, device=q.device)
    if block_tables is not None:
        if block_tables.dtype != torch.int32 or block_tables.stride(1) != 1:
            raise TypeError("block_tables must be row-major int32")
        nb, bs = _cache_dims(k, num_kv_heads * D)
        ks = vs = num_kv_heads * D
        bt_ptr, bt_stride = block_tables.data_ptr(), block_tables.stride(0)
    else:
        if k.stride(2) != 1 or k.stride(1) != D or v.stride(2) != 1 or v.stride(1) != D:
            raise ValueError("k/v must be contiguous in (heads, head_dim)")
        nb = bs = 0
        ks = k.stride(0) if k.shape[0] > 1 else num_kv_heads * D
        vs = v.stride(0) if v.shape[0] > 1 else num_kv_heads * D
        bt_ptr, bt_stride = None, 0
    if isinstance(workspace, str):
        workspace = prefill_workspace(q.device, D)
    ws_ptr, ws_bytes = (None, 0) if workspace is None else (workspace.data_ptr(), workspace.numel())
    _C.check(_C.load().b200_prefill_varlen_attn(
        out.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), cu_seqlens_q.data_ptr(),
        cu_seqlens_k.data_ptr(), num_seqs, Tq, 0 if block_tables is not None else k.shape[0],
        int(max_seqlen_q), int(max_seqlen_k), H, num_kv_heads,
        D, q.stride(0) if Tq > 1 else H * D, ks, vs, out.stride(0) if Tq > 1 else H * D, bt_ptr,
        bt_stride, nb, bs, float(scale), _dt(q), ws_ptr, ws_bytes, _stream()), "b200_prefill_varlen_attn")
    return out


def qknorm_rope_cache(qkv, positions, cos_sin_cache, q_norm_weight, k_norm_weight, eps,
                      k_cache, v_cache, slots, num_heads, num_kv_heads, head_dim):
    """K8, in place on qkv [T, (H+2Hkv)*D]. k_cache/v_cache/slots may be None (no cache bound)."""
    _need_cuda(qkv, positions, cos_sin_cache, q_norm_weight, k_norm_weight, k_cache, v_cache, slots)
    T = qkv.shape[0]
    if qkv.dim() != 2 or qkv.stride(1) != 1:
        raise ValueError("qkv must be [T, (H+2Hkv)*D] with unit inner stride")
    if positions.dtype != torch.int64 or not positions.is_contiguous():
        raise TypeError("positions must be contiguous int64")
    if cos_sin_cache.dtype != torch.float32 or not cos_sin_cache.is_contiguous():
        raise TypeError("cos_sin_cache must be contiguous fp32")
    if cos_sin_cache.shape[-1] != head_dim:
        raise ValueError("cos_sin_cache last dim must equal head_dim")
    have_cache = k_cache is not None and k_cache.numel() > 0
    nb = bs = 0
    if have_cache:
        nb, bs = _cache_dims(k_cache, num_kv_heads * head_dim)
        if slots.dtype != torch.int32 or not slots.is_contiguous():
            raise TypeError("slots must be contiguous int32")
    _C.check(_C.load().b200_qknorm_rope_cache(
        qkv.data_ptr(), positions.data_ptr(), cos_sin_cache.data_ptr(), _ptr(q_norm_weight),
        _ptr(k_norm_weight), float(eps), k_cache.data_ptr() if have_cache else None,
        v_cache.data_ptr() if have_cache else None, slots.data_ptr() if have_cache else None,
        1 if (have_cache and slots.dim() == 2) else 0, T, num_heads, num_kv_heads, head_dim,
        qkv.stride(0) if T > 1 else qkv.shape[1], cos_sin_cache.shape[0], nb, bs, _dt(qkv),
        _stream()), "b200_qknorm_rope_cache")
    return qkv


def rmsnorm(x, weight, eps, out=None):
    _need_cuda(x, weight)
    if not x.is_contiguous():
        raise ValueError("rmsnorm input must be contiguous")
    hidden = x.shape[-1]
    rows = x.numel() // hidden if hidden else 0
    if out is None:
        out = torch.empty_like(x)
    _C.check(_C.load().b200_rmsnorm(out.data_ptr(), x.data_ptr(), weight.data_ptr(), float(eps),
                                    rows, hidden, _dt(x), _stream()), "b200_rmsnorm")
    return out


def add_rmsnorm_(x, residual, weight, eps):
    """In place: residual <- x + residual; x <- RMSNorm(x + residual) * weight."""
    _need_cuda(x, residual, weight)
    if not (x.is_contiguous() and residual.is_contiguous()):
        raise ValueError("add_rmsnorm tensors must be contiguous")
    hidden = x.shape[-1]
    rows = x.numel() // hidden if hidden else 0
    _C.check(_C.load().b200_add_rmsnorm(x.data_ptr(), residual.data_ptr(), weight.data_ptr(),
                                        float(eps), rows, hidden, _dt(x), _stream()),
             "b200_add_rmsnorm")
    return x, residual


def silu_and_mul(x, out=None):
    _need_cuda(x)
    if not x.is_contiguous():
        raise ValueError("silu_and_mul input must be contiguous")
    inter = x.shape[-1] // 2
    rows = x.numel() // (2 * inter) if inter else 0
    if out is None:
        out = torch.empty(*x.shape[:-1], inter, dtype=x.dtype, device=x.device)
    _C.check(_C.load().b200_silu_and_mul(out.data_ptr(), x.data_ptr(), rows, inter, _dt(x),
                                         _stream()), "b200_silu_and_mul")
    return out


def sample(logits, temperatures=None, seed=0, step=0, out=None, step_offset=None):
    """argmax (temperatures=None or <= 0) or Gumbel-max sampling over [rows, vocab] bf16/fp16
    logits. step_offset: optional device int64 scalar added to `step` when the kernel runs."""
    _need_cuda(logits, temperatures, step_offset)
    rows, vocab = logits.shape
    if logits.stride(1) != 1:
        raise ValueError("logits must have unit inner stride")
    if temperatures is not None and (temperatures.dtype != torch.float32 or
                                     not temperatures.is_contiguous()):
        raise TypeError("temperatures must be contiguous fp32")
    if out is None:
        out = torch.empty(rows, dtype=torch.int64, device=logits.device)
    _C.check(_C.load().b200_sample(out.data_ptr(), logits.data_ptr(), _ptr(temperatures), rows,
                                   vocab, logits.stride(0) if rows > 1 else vocab, int(seed),
                                   int(step), _ptr(step_offset), _dt(logits), _stream()),
             "b200_sample")
    return out


def sample_shard(logits, temperatures=None, seed=0, step=0, index_offset=0, out=None, step_offset=None):
    """Vocab-sharded sampling (tensor parallelism): packed (score, global index) keys of this rank's
    shard, to be MAX-all-reduced over the ranks; see include/b200attn.h::b200_sample_shard."""
    _need_cuda(logits, temperatures, step_offset)
    rows, vocab = logits.shape
    if logits.stride(1) != 1:
        raise ValueError("logits must have unit inner stride")
    if temperatures is not None and (temperatures.dtype != torch.float32 or
                                     not temperatures.is_contiguous()):
        raise TypeError("temperatures must be contiguous fp32")
    if out is None:
        out = torch.empty(rows, dtype=torch.int64, device=logits.device)
    _C.check(_C.load().b200_sample_shard(out.data_ptr(), logits.data_ptr(), _ptr(temperatures), rows, vocab,
                                         logits.stride(0) if rows > 1 else vocab, int(index_offset),
                                         int(seed), int(step), _ptr(step_offset), _dt(logits), _stream()),
             "b200_sample_shard")
    return out


def decode_sample_keys(keys: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Token ids from (all-reduced) sample_shard keys: id = 0x7fffffff - (key & 0x7fffffff)."""
    ids = 0x7FFFFFFF - (keys & 0x7FFFFFFF)
    if out is None:
        return ids
    out.copy_(ids)
    return out


# ------------------------------------------------------------------ dense layers on the grouped tcgen05 kernel
_dense_offsets: dict = {}


def _dense_plan(rows: int, device) -> torch.Tensor:
    """offsets = {0, rows}: one group holding every row (csrc/moe_umma.cu serves an ordinary linear
    layer as a one-expert group). Cached per (rows, device) so that graph capture never allocates it."""
    key = (rows, device)
    t = _dense_offsets.get(key)
    if t is None:
        t = _dense_offsets[key] = torch.tensor([0, rows], dtype=torch.int32, device=device)
    return t


def gate_up_silu_umma(x: torch.Tensor, w_gate_up: torch.Tensor) -> torch.Tensor:
    """SiluAndMul(x @ w_gate_up.T) in ONE kernel: the gate|up GEMM of a dense MLP on the grouped tcgen05
    kernel (one group) with the activation as its TMEM epilogue. Not on the decode path: at M = 256
    it measured 41.87 k tok/s against 41.93 k for cuBLAS + silu_and_mul (no gain, r02).
    x [M, K] contiguous; w_gate_up [2*I, K] (gate rows, then up rows); K % 64 == 0, I % 64 == 0."""
    _need_cuda(x, w_gate_up)
    M, K = x.shape
    inter = w_gate_up.shape[0] // 2
    if not x.is_contiguous() or not w_gate_up.is_contiguous() or w_gate_up.shape[1] != K:
        raise ValueError("gate_up_silu_umma takes contiguous x [M, K] and w [2*I, K]")
    h = torch.empty(M, inter, dtype=x.dtype, device=x.device)
    _C.check(_C.load().b200_moe_gemm_gate_up_umma(h.data_ptr(), x.data_ptr(), w_gate_up.data_ptr(),
                                                  _dense_plan(M, x.device).data_ptr(), M, K, inter, 1,
                                                  _dt(x), _stream()), "b200_moe_gemm_gate_up_umma (dense)")
    return h


def linear_umma(x: torch.Tensor, w: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """x @ w.T on the same kernel (plain epilogue). x [M, K]; w [N, K]; K % 64 == 0, N % 128 == 0."""
    _need_cuda(x, w)
    M, K = x.shape
    N = w.shape[0]
    if not x.is_contiguous() or not w.is_contiguous() or w.shape[1] != K:
        raise ValueError("linear_umma takes contiguous x [M, K] and w [N, K]")
    if out is None:
        out = torch.empty(M, N, dtype=x.dtype, device=x.device)
    elif out.shape != (M, N) or not out.is_contiguous() or out.dtype != x.dtype:
        raise ValueError("out must be a contiguous [M, N] tensor of x's dtype")
    _C.check(_C.load().b200_moe_gemm_down_umma(out.data_ptr(), x.data_ptr(), w.data_ptr(),
                                               _dense_plan(M, x.device).data_ptr(), None, None, M, K, N, 1,
                                               _dt(x), _stream()), "b200_moe_gemm_down_umma (dense)")
    return out


# ------------------------------------------------------------------ sparse MoE (csrc/moe.cu)
class MoePlan:
    """Device-side dispatch plan of one MoE layer call (b200_moe_align): fixed-capacity buffers,
    nothing is read back by the host."""
    __slots__ = ("offsets", "sorted_pair", "pos_of_pair", "counts", "num_local", "expert_begin", "top_k")

    def __init__(self, num_pairs: int, num_local: int, expert_begin: int, top_k: int, device):
        i32 = dict(dtype=torch.int32, device=device)
        self.offsets = torch.empty(num_local + 1, **i32)
        self.sorted_pair = torch.zeros(num_pairs, **i32)     # tail beyond offsets[-1] stays 0 (a valid pair index)
        self.pos_of_pair = torch.empty(num_pairs, **i32)
        self.counts = torch.empty(num_local, **i32)
        self.num_local, self.expert_begin, self.top_k = num_local, expert_begin, top_k


def moe_route(router_logits: torch.Tensor, top_k: int):
    """softmax (fp32) over all experts + top-k + renormalise (qwen3_moe.py:150-158).
    router_logits [T, E] bf16/fp16 -> (topk_ids int32 [T, k], topk_weights dtype [T, k])."""
    _need_cuda(router_logits)
    T, E = router_logits.shape
    if router_logits.stride(1) != 1:
        raise ValueError("router_logits must be contiguous along the expert dimension")
    ids = torch.empty(T, top_k, dtype=torch.int32, device=router_logits.device)
    w = torch.empty(T, top_k, dtype=router_logits.dtype, device=router_logits.device)
    _C.check(_C.load().b200_moe_route(ids.data_ptr(), w.data_ptr(), router_logits.data_ptr(), T, E, top_k,
                                      router_logits.stride(0) if T > 1 else E, _dt(router_logits),
                                      _stream()), "b200_moe_route")
    return ids, w


def moe_align(topk_ids: torch.Tensor, expert_begin: int, num_local: int) -> MoePlan:
    """Pairs (t, j) stably sorted by local expert; see include/b200attn.h."""
    _need_cuda(topk_ids)
    if topk_ids.dtype != torch.int32 or not topk_ids.is_contiguous():
        raise TypeError("topk_ids must be a contiguous int32 tensor")
    T, k = topk_ids.shape
    plan = MoePlan(T * k, num_local, expert_begin, k, topk_ids.device)
    _C.check(_C.load().b200_moe_align(plan.offsets.data_ptr(), plan.sorted_pair.data_ptr(),
                                      plan.pos_of_pair.data_ptr(), plan.counts.data_ptr(),
                                      topk_ids.data_ptr(), T * k, expert_begin, num_local, _stream()),
             "b200_moe_align")
    return plan


def moe_experts(x: torch.Tensor, w13: torch.Tensor, w2: torch.Tensor, topk_ids: torch.Tensor,
                topk_weights: torch.Tensor, plan: MoePlan, out: torch.Tensor | None = None) -> torch.Tensor:
    """The expert MLPs of the local experts and the per-token combine.
    x [T, hidden]; w13 [num_local, 2*inter, hidden]; w2 [num_local, hidden, inter].
    Returns out [T, hidden]: this rank's partial sum (the full result when it owns every expert)."""
    _need_cuda(x, w13, w2, topk_ids, topk_weights)
    T, hidden = x.shape
    n_local, two_i, h2 = w13.shape
    inter = two_i // 2
    if h2 != hidden or w2.shape != (n_local, hidden, inter) or n_local != plan.num_local:
        raise ValueError("expert weight shapes do not match")
    if not (w13.is_contiguous() and w2.is_contiguous() and topk_weights.is_contiguous()):
        raise ValueError("expert weights / routing weights must be contiguous")
    if x.stride(1) != 1 or x.dtype != w13.dtype or x.dtype != w2.dtype or x.dtype != topk_weights.dtype:
        raise ValueError("x must be row-major and share its dtype with the weights")
    k = plan.top_k
    pairs = T * k
    lib = _C.load()
    if _MOE_BIG_GEMM != "stream" and pairs >= _MOE_BIG_GEMM_ROWS * n_local:
        big = _moe_expert_gemms_umma if _MOE_BIG_GEMM == "umma" else _moe_expert_gemms_library
        y = big(x, w13, w2, topk_weights, plan)
        return _moe_combine(lib, y, topk_ids, plan, T, hidden, out, x)
    h = torch.empty(pairs, inter, dtype=x.dtype, device=x.device)
    y = torch.empty(pairs, hidden, dtype=x.dtype, device=x.device)
    _C.check(lib.b200_moe_gemm_gate_up(h.data_ptr(), x.data_ptr(), w13.data_ptr(), plan.offsets.data_ptr(),
                                       plan.sorted_pair.data_ptr(), hidden, inter, n_local, k,
                                       x.stride(0) if T > 1 else hidden, _dt(x), _stream()),
             "b200_moe_gemm_gate_up")
    _C.check(lib.b200_moe_gemm_down(y.data_ptr(), h.data_ptr(), w2.data_ptr(), plan.offsets.data_ptr(),
                                    plan.sorted_pair.data_ptr(), topk_weights.data_ptr(), inter, hidden,
                                    n_local, k, _dt(x), _stream()), "b200_moe_gemm_down")
    return _moe_combine(lib, y, topk_ids, plan, T, hidden, out, x)


# Prefill-sized batches (hundreds of rows per expert) are tensor-bound, not weight-streaming: from
# _MOE_BIG_GEMM_ROWS average rows per local expert on they go through the tcgen05 grouped GEMM of
# csrc/moe_umma.cu (measured on Qwen3-30B-A3B, 4096-token prefill layer: 721 us vs 1224 us for the
# weight-streaming kernel; whole-model prefill 99.9 k vs 51.8 k tok/s). Module attributes, not
# environment switches: "stream" forces the weight-streaming kernel for every size, "library" routes
# through torch._grouped_mm (comparator for tools/moe_diag.py and the tests).
_MOE_BIG_GEMM = "umma"
_MOE_BIG_GEMM_ROWS = 96


def _moe_expert_gemms_umma(x, w13, w2, topk_weights, plan: MoePlan) -> torch.Tensor:
    T, hidden = x.shape
    n_local, two_i, _ = w13.shape
    inter, k = two_i // 2, plan.top_k
    pairs = T * k
    lib = _C.load()
    xs = torch.empty(pairs, hidden, dtype=x.dtype, device=x.device)
    h = torch.empty(pairs, inter, dtype=x.dtype, device=x.device)
    y = torch.empty(pairs, hidden, dtype=x.dtype, device=x.device)
    _C.check(lib.b200_moe_gather_rows(xs.data_ptr(), x.data_ptr(), plan.offsets.data_ptr(),
                                      plan.sorted_pair.data_ptr(), pairs, n_local, k, hidden,
                                      x.stride(0) if T > 1 else hidden, _dt(x), _stream()),
             "b200_moe_gather_rows")
    _C.check(lib.b200_moe_gemm_gate_up_umma(h.data_ptr(), xs.data_ptr(), w13.data_ptr(),
                                            plan.offsets.data_ptr(), pairs, hidden, inter, n_local,
                                            _dt(x), _stream()), "b200_moe_gemm_gate_up_umma")
    _C.check(lib.b200_moe_gemm_down_umma(y.data_ptr(), h.data_ptr(), w2.data_ptr(), plan.offsets.data_ptr(),
                                         plan.sorted_pair.data_ptr(), topk_weights.data_ptr(), pairs, inter,
                                         hidden, n_local, _dt(x), _stream()), "b200_moe_gemm_down_umma")
    return y


def _moe_expert_gemms_library(x, w13, w2, topk_weights, plan: MoePlan) -> torch.Tensor:
    k = plan.top_k
    pair = plan.sorted_pair.long()                    # tail entries are 0: harmless rows, never combined
    xs = x.index_select(0, pair // k)                 # [pairs, hidden] in expert order
    offs = plan.offsets[1:]
    gu = torch._grouped_mm(xs, w13.transpose(1, 2), offs=offs)
    h = silu_and_mul(gu)
    y = torch._grouped_mm(h, w2.transpose(1, 2), offs=offs)
    return y * topk_weights.reshape(-1)[pair].unsqueeze(1)      # dtype multiply, as qwen3_moe.py:180-182


def _moe_combine(lib, y, topk_ids, plan: MoePlan, T, hidden, out, x):
    k = plan.top_k
    if out is None:
        out = torch.empty(T, hidden, dtype=x.dtype, device=x.device)
    elif out.shape != (T, hidden) or out.dtype != x.dtype or out.stride(1) != 1:
        raise ValueError("out must be a row-major [T, hidden] tensor of x's dtype")
    _C.check(lib.b200_moe_combine(out.data_ptr(), y.data_ptr(), topk_ids.data_ptr(),
                                  plan.pos_of_pair.data_ptr(), T, k, hidden,
                                  out.stride(0) if T > 1 else hidden, _dt(x), _stream()),
             "b200_moe_combine")
    return out
