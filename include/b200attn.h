/*
 * b200attn.h -- C ABI of libb200attn.so: the paged-KV decode/attention hot path
 * of nano-vllm (reference: linzm1007/nano-vllm-ascend) as hand-written sm_100a CUDA.
 *
 * Boundary contract:
 *   - plain C, raw device pointers + sizes, no torch types;
 *   - the caller owns every buffer (q/k/v, paged caches, tables, outputs, workspaces);
 *   - every entry point only ENQUEUES work on `stream` (a cudaStream_t): no allocation,
 *     no synchronisation, no host<->device copies  => CUDA-graph capturable;
 *   - return value: 0 on success, a negative B200_ERR_* for rejected arguments, a
 *     positive cudaError_t value if the launch itself failed;
 *   - KV caches are written in place (same as the reference's
 *     torch_npu._npu_reshape_and_cache / scatter_update_, layers/attention.py:25-35).
 *
 * The reference has no native FFI of its own: its "FFI" for this path is the vendor op
 * table reached through torch dispatch (torch_npu.* / torchair / flash_attn / triton).
 * Each entry point below cites the reference call site(s) it replaces; paths are
 * relative to the reference checkout, line numbers as of commit ec1ffe9a.
 *
 * dtype: element type of q/k/v/caches/activations: B200_DTYPE_BF16 or B200_DTYPE_F16.
 * All index tensors are int32 unless stated (positions: int64, as the reference's
 * model_runner.py:274-275 builds them).
 */
#ifndef B200ATTN_H_
#define B200ATTN_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200ATTN_VERSION 100 /* major*10000 + minor*100 + patch */

typedef void* b200_stream_t; /* cudaStream_t */

enum {
  B200_DTYPE_BF16 = 0,
  B200_DTYPE_F16 = 1,
};

enum {
  B200_OK = 0,
  B200_ERR_NULL_POINTER = -1,
  B200_ERR_BAD_SHAPE = -2,       /* zero/negative dims, H % Hkv != 0, ... */
  B200_ERR_UNSUPPORTED = -3,     /* head_dim not in {64,128}, group size > 8, Hkv not 1/2/4/8k ... */
  B200_ERR_MISALIGNED = -4,      /* pointer / stride not 16-byte aligned */
  B200_ERR_WORKSPACE = -5,       /* workspace too small */
  B200_ERR_DRIVER = -6,          /* cuTensorMapEncodeTiled unavailable / failed */
  B200_ERR_BAD_DTYPE = -7,
};

/* Version of the loaded library (== B200ATTN_VERSION it was built with). */
int b200_version(void);
/* Human-readable text for a return code of any function below. Never NULL. */
const char* b200_error_string(int code);
/* Number of kernels this library has enqueued since load / since the last reset
 * (host-side counter; lets bench.py report `gpu_launches`). */
int64_t b200_launch_count(void);
void b200_reset_launch_count(void);

/* ---------------------------------------------------------------------------------
 * K1  reshape_and_cache, flat slots.
 * Replaces: torch_npu._npu_reshape_and_cache        layers/attention.py:25-30
 *           triton store_kvcache_kernel              layers/attention_ori.py:10-40
 *           store_kvcache (torch native, the oracle) layers/attention_torch_native.py:13-51
 * key/value: [num_tokens, num_kv_heads, head_dim], token stride in ELEMENTS
 *            (key_stride/value_stride; heads and head_dim contiguous).
 * k_cache/v_cache: [num_blocks, block_size, num_kv_heads*head_dim] contiguous.
 * slot_mapping[i] = block*block_size + offset; slot == -1 or out of range => skipped
 * (attention_ori.py:23, attention_torch_native.py:27-35).
 * ------------------------------------------------------------------------------- */
int b200_reshape_and_cache(const void* key, const void* value, void* k_cache, void* v_cache,
                           const int32_t* slot_mapping, int32_t num_tokens,
                           int32_t num_kv_heads, int32_t head_dim, int64_t key_stride,
                           int64_t value_stride, int32_t num_blocks, int32_t block_size,
                           int32_t dtype, b200_stream_t stream);

/* K2  reshape_and_cache, 2-D slots [num_tokens, 2] = (block_id, offset), the decode form.
 * Replaces: torch_npu.scatter_update_ x2             layers/attention.py:32-35
 *           (index built at engine/model_runner.py:301,315,353). */
int b200_reshape_and_cache_2d(const void* key, const void* value, void* k_cache,
                              void* v_cache, const int32_t* slot_pairs, int32_t num_tokens,
                              int32_t num_kv_heads, int32_t head_dim, int64_t key_stride,
                              int64_t value_stride, int32_t num_blocks, int32_t block_size,
                              int32_t dtype, b200_stream_t stream);

/* ---------------------------------------------------------------------------------
 * K6  paged decode attention (one query token per sequence).
 * Replaces: npu_fused_infer_attention_score_v2 BNSD + block_table  layers/attention.py:64-76
 *           tng.ops.npu_fused_infer_attention_score (graph mode)   layers/attention.py:80-92
 *           flash_attn_with_kvcache                                layers/attention_ori.py:72-74
 *           Attention._decode_attn (oracle)                layers/attention_torch_native.py:147-196
 * q:   [batch, num_heads, head_dim], token stride q_stride (elements).
 * out: [batch, num_heads*head_dim],  token stride out_stride (elements).
 * k_cache/v_cache: [num_blocks, block_size, num_kv_heads*head_dim] contiguous.
 * block_tables: [batch, block_table_stride] int32, -1 padded; context_lens: [batch] int32
 * (tokens to attend INCLUDING the one just stored). context_len == 0 => that row of
 * `out` is written as zeros and nothing else is touched (padded graph rows,
 * engine/model_runner.py:303-311).
 * workspace: >= b200_paged_decode_attn_workspace_size(...) bytes, 16-byte aligned.
 * num_ctas: 0 = auto (one persistent CTA per SM), otherwise the persistent grid size.
 * Supported (anything else returns B200_ERR_UNSUPPORTED, and the workspace-size query returns 0):
 *   head_dim 64 or 128; num_kv_heads 1, 2, 4 or a multiple of 8; GQA group num_heads/num_kv_heads <= 8
 *   (the group is the M dimension of one m16n8k16 tile; every BASELINE model fits -- the reference's
 *   native path takes any grouping with num_heads <= 32, attention_torch_native.py:53-58);
 *   block_size a multiple of 16; batch * ceil(num_kv_heads/8) <= 2048 per launch.
 * ------------------------------------------------------------------------------- */
size_t b200_paged_decode_attn_workspace_size(int32_t batch, int32_t num_heads,
                                             int32_t num_kv_heads, int32_t head_dim);
int b200_paged_decode_attn(void* out, const void* q, const void* k_cache, const void* v_cache,
                           const int32_t* block_tables, const int32_t* context_lens,
                           int32_t batch, int32_t num_heads, int32_t num_kv_heads,
                           int32_t head_dim, int64_t q_stride, int64_t out_stride,
                           int32_t num_blocks, int32_t block_size,
                           int32_t block_table_stride, float scale, void* workspace,
                           size_t workspace_bytes, int32_t num_ctas, int32_t dtype,
                           b200_stream_t stream);

/* ---------------------------------------------------------------------------------
 * K1/K2 + K6 over an FP8 (e4m3) KV cache -- SURVEY.md 8f row f3. An EXTENSION of the path: the
 * reference keeps its caches in the model dtype (engine/model_runner.py:218-229); nothing in it is
 * replaced by these two entry points, they halve the bytes K6 streams.
 * Cache element = e4m3(x * (1/scale)), round-to-nearest-even, saturating at +-448; k_scale and
 * v_scale are per-tensor fp32 scales (> 0). Caches: [num_blocks, block_size, num_kv_heads*head_dim]
 * BYTES, contiguous. Everything else (layouts, slots, block tables, context_len == 0 rows,
 * workspace, return codes) is as for b200_reshape_and_cache[_2d] / b200_paged_decode_attn;
 * `dtype` is the element type of key/value/q/out. Attention is evaluated over the de-quantised
 * values k8*k_scale, v8*v_scale with fp32 softmax and accumulation.
 * Extra limits: q/out 8-byte aligned with strides % 4 == 0; head_dim 64 needs >= 2 kv heads.
 * ------------------------------------------------------------------------------- */
int b200_reshape_and_cache_fp8(const void* key, const void* value, void* k_cache, void* v_cache,
                               const int32_t* slots, int32_t slot_is_2d, int32_t num_tokens,
                               int32_t num_kv_heads, int32_t head_dim, int64_t key_stride,
                               int64_t value_stride, int32_t num_blocks, int32_t block_size,
                               float k_scale, float v_scale, int32_t dtype, b200_stream_t stream);
int b200_paged_decode_attn_fp8(void* out, const void* q, const void* k_cache, const void* v_cache,
                               const int32_t* block_tables, const int32_t* context_lens,
                               int32_t batch, int32_t num_heads, int32_t num_kv_heads,
                               int32_t head_dim, int64_t q_stride, int64_t out_stride,
                               int32_t num_blocks, int32_t block_size, int32_t block_table_stride,
                               float scale, float k_scale, float v_scale, void* workspace,
                               size_t workspace_bytes, int32_t num_ctas, int32_t dtype,
                               b200_stream_t stream);

/* Paged e4m3 cache -> packed 16-bit K/V for the prefill kernel (prefix-cache-hit / chunked prefill
 * over an fp8 cache; extension, the reference's caches are always the model dtype). For sequence i,
 * rows cu_seqlens_k[i] .. cu_seqlens_k[i+1]-1 of k_out / v_out [total_k, num_kv_heads*head_dim]
 * receive its first cached tokens, gathered through block_tables[i] and multiplied by k_scale /
 * v_scale (fp32, one rounding to dtype). Table entries outside [0, num_blocks) give zero rows. */
int b200_gather_dequant_fp8(void* k_out, void* v_out, const void* k_cache, const void* v_cache,
                            const int32_t* block_tables, const int32_t* cu_seqlens_k, int32_t num_seqs,
                            int32_t max_seqlen_k, int32_t block_table_stride, int32_t num_blocks,
                            int32_t block_size, int32_t num_kv_heads, int32_t head_dim, float k_scale,
                            float v_scale, int32_t dtype, b200_stream_t stream);

/* ---------------------------------------------------------------------------------
 * K4/K5  varlen causal prefill flash-attention.
 * Replaces: npu_fused_infer_attention_score_v2 TND, sparse_mode=3  layers/attention.py:49-59
 *           flash_attn_varlen_func(causal=True[, block_table])     layers/attention_ori.py:63-70
 *           Attention._prefill_attn (oracle)               layers/attention_torch_native.py:103-145
 * q: [total_q, num_heads, head_dim] (token stride q_stride), out: [total_q, num_heads*head_dim].
 * Contiguous K/V mode (block_tables == NULL): k/v: [total_k, num_kv_heads, head_dim] with
 *   token strides k_stride/v_stride; cu_seqlens_k indexes them.
 * Paged K/V mode (block_tables != NULL, the prefix-cache-hit form of attention_ori.py:65-66):
 *   k/v are the paged caches [num_blocks, block_size, num_kv_heads*head_dim]; sequence i's
 *   keys are its first (cu_seqlens_k[i+1]-cu_seqlens_k[i]) cached tokens via block_tables[i].
 * total_q / total_k: number of packed rows in q / in k,v (= cu_seqlens[num_seqs]; bounds of the
 * TMA tensor maps; total_k is ignored in paged mode).
 * workspace (optional, 16-byte aligned device memory owned by the caller, may be NULL): when the
 *   (sequence, head, 256-row) grid would leave most SMs idle -- the prefix-cache-hit form, a few
 *   hundred query rows against thousands of cached keys -- up to 8 CTAs split one row block along
 *   the keys, write un-normalised fp32 partials here and a second small kernel merges them. Needs
 *   total_q * num_heads * splits * (head_dim + 2) * 4 bytes; see b200_prefill_workspace_size.
 * Causal mask is bottom-right aligned (query row r of a sequence with Lq queries and Lk keys
 * sees keys 0 .. r + Lk - Lq), == NPU sparse_mode=3 and flash-attn causal; identical to the
 * oracle's top-left mask whenever Lq == Lk, which is the only case the reference runner
 * produces (engine/model_runner.py:252-256).
 * ------------------------------------------------------------------------------- */
int b200_prefill_varlen_attn(void* out, const void* q, const void* k, const void* v,
                             const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k,
                             int32_t num_seqs, int32_t total_q, int32_t total_k,
                             int32_t max_seqlen_q, int32_t max_seqlen_k,
                             int32_t num_heads, int32_t num_kv_heads, int32_t head_dim,
                             int64_t q_stride, int64_t k_stride, int64_t v_stride,
                             int64_t out_stride, const int32_t* block_tables,
                             int32_t block_table_stride, int32_t num_blocks,
                             int32_t block_size, float scale, int32_t dtype,
                             void* workspace, int64_t workspace_bytes, b200_stream_t stream);

/* Bytes of split-KV workspace that let b200_prefill_varlen_attn use every SM for ANY launch it would
 * split (few query rows against many keys): the largest such launch has 148 CTAs x 256 rows of
 * un-normalised fp32 partials. A smaller (or NULL) workspace is legal: fewer (or no) key splits. */
int64_t b200_prefill_workspace_size(int32_t head_dim);

/* Tuning / test hook, process-wide: the tcgen05 prefill kernel has two CTA layouts -- 1 query tile of
 * 128 rows per CTA with two CTAs per SM, or 2 ping-pong tiles per CTA with one CTA per SM -- and picks
 * per launch (short sequences and head_dim 64 -> 1; long head_dim-128 sequences and key-split launches
 * -> 2; csrc/prefill_umma.cu::prefill_query_tiles). 1 or 2 forces a layout, 0 restores the choice.
 * Results are the same up to fp32 summation order across key splits. Not thread-safe against
 * concurrent b200_prefill_varlen_attn calls. */
int b200_prefill_set_query_tiles(int32_t query_tiles_per_cta);

/* ---------------------------------------------------------------------------------
 * K8  fused qkv epilogue: split -> per-head RMSNorm(q,k) -> NeoX RoPE(q,k) -> KV scatter.
 * Replaces the op chain models/qwen3.py:79-88 (q_norm/k_norm: layers/layernorm.py:16-25;
 * rotary: layers/rotary_embedding.py:6-14,37-47; scatter: K1/K2 above).
 * qkv: [num_tokens, (H + 2*Hkv)*D] contiguous rows (row stride qkv_stride elements);
 *      q and k are normalised + rotated IN PLACE (v untouched), so q = qkv[:, :H*D] etc.
 *      stay valid views for the attention kernels.
 * positions: [num_tokens] int64. cos_sin_cache: [max_position, D] fp32 (cos | sin halves),
 *      rotary_embedding.py:26-35.
 * q_norm_weight/k_norm_weight: [D] in dtype, or NULL for models without QK-norm
 *      (qkv_bias models, models/qwen3.py:70-72; Llama).
 * k_cache/v_cache may be NULL (no cache bound yet: warm-up, attention_ori.py:62).
 * slots: flat [num_tokens] when slot_is_2d == 0, else [num_tokens,2] (block, offset).
 * ------------------------------------------------------------------------------- */
int b200_qknorm_rope_cache(void* qkv, const int64_t* positions, const float* cos_sin_cache,
                           const void* q_norm_weight, const void* k_norm_weight, float eps,
                           void* k_cache, void* v_cache, const int32_t* slots,
                           int32_t slot_is_2d, int32_t num_tokens, int32_t num_heads,
                           int32_t num_kv_heads, int32_t head_dim, int64_t qkv_stride,
                           int32_t max_position, int32_t num_blocks, int32_t block_size,
                           int32_t dtype, b200_stream_t stream);

/* F1  RMSNorm: out = cast(x_fp32 * rsqrt(mean(x^2)+eps)) * weight   layers/layernorm.py:16-25
 * x, out: [rows, hidden] contiguous (out may alias x). */
int b200_rmsnorm(void* out, const void* x, const void* weight, float eps, int32_t rows,
                 int32_t hidden, int32_t dtype, b200_stream_t stream);

/* F1' fused residual-add + RMSNorm                                   layers/layernorm.py:27-38
 * residual <- cast(x + residual); x <- RMSNorm(fp32(x + residual)) * weight. Both in place. */
int b200_add_rmsnorm(void* x, void* residual, const void* weight, float eps, int32_t rows,
                     int32_t hidden, int32_t dtype, b200_stream_t stream);

/* F3  SiluAndMul: out[r, i] = silu(x[r, i]) * x[r, inter + i]         layers/activation.py:11-13
 * x: [rows, 2*inter] contiguous, out: [rows, inter] contiguous. */
int b200_silu_and_mul(void* out, const void* x, int32_t rows, int32_t inter, int32_t dtype,
                      b200_stream_t stream);

/* f2 (SURVEY 8f "next"): greedy/temperature sampling tail without the fp32 logits
 * round trip. Replaces layers/sampler.py:9-17 for the greedy parity path (argmax) and
 * for temperature sampling via the Gumbel-max trick with a counter-based RNG.
 * logits: [rows, vocab] in dtype (row stride logits_stride elements);
 * temperatures: [rows] fp32 or NULL; a row whose temperature is NULL/<= 0 takes the argmax.
 * out_tokens: [rows] int64. step_offset: optional DEVICE int64 added to `step` at run time, so
 * a captured CUDA graph draws fresh randomness on every replay (NULL = none).
 * Gumbel sampling draws the same distribution as softmax+multinomial but not the same
 * random stream as torch.multinomial; only the argmax form is bit-comparable. */
int b200_sample(int64_t* out_tokens, const void* logits, const float* temperatures,
                int32_t rows, int32_t vocab, int64_t logits_stride, uint64_t seed,
                uint64_t step, const int64_t* step_offset, int32_t dtype,
                b200_stream_t stream);

/* f2, vocab-sharded (tensor parallelism). Replaces the logits gather to rank 0 (layers/embed_head.py:62-65)
 * + the full-vocabulary sampling pass there: each rank scores its own vocabulary shard
 * logits [rows, vocab_local] (global index = index_offset + local index; the Gumbel draw is keyed by
 * the GLOBAL index, so the shards see exactly the unsharded scores) and writes one int64 key per row,
 *   key = order_preserving_u32(best score) << 31 | (0x7fffffff - global index of the winner).
 * A MAX all-reduce of the keys over the ranks (8 bytes per row) is the global arg-max with
 * b200_sample's tie rule; token id = 0x7fffffff - (key & 0x7fffffff). Same seed/step/step_offset
 * contract as b200_sample; every rank must pass the same values. */
int b200_sample_shard(int64_t* out_keys, const void* logits, const float* temperatures, int32_t rows,
                      int32_t vocab_local, int64_t logits_stride, int32_t index_offset, uint64_t seed,
                      uint64_t step, const int64_t* step_offset, int32_t dtype, b200_stream_t stream);

/* ---------------------------------------------------------------------------------
 * MoE  sparse mixture-of-experts MLP: routing, dispatch plan, expert GEMMs, combine.
 * Replaces: Qwen3MoeSparseMoeBlock.forward                            models/qwen3_moe.py:146-186
 *           (softmax/top-k :150-158, per-expert gather + Qwen3MoeMLP :95-122 + index_add_ :160-186).
 * Expert parallelism: the caller owns experts [expert_begin, expert_begin + num_local) whole and
 * sees every token; pairs routed elsewhere are skipped (pos_of_pair = -1) and the per-rank partial
 * outputs are summed by the row-parallel all-reduce of the decoder layer. All buffers have fixed
 * capacity (tokens * top_k pairs), nothing is read back by the host.
 *
 * b200_moe_route: router_logits [tokens, num_experts] (dtype; row stride logits_stride elements) ->
 *   topk_ids int32 [tokens, top_k] (descending probability, ties to the lower index) and
 *   topk_weights (dtype) = renormalised fp32 softmax probabilities. top_k <= 8, num_experts <= 256.
 * b200_moe_align: pairs (t, j) = index t*top_k + j, stably sorted by local expert ->
 *   offsets int32 [num_local + 1], sorted_pair int32 [num_pairs] (first offsets[num_local] valid),
 *   pos_of_pair int32 [num_pairs] (-1 = not local); counts int32 [num_local] is scratch.
 * b200_moe_gemm_gate_up: h[pos, :inter] = SiluAndMul(x[sorted_pair[pos] / top_k] . w13[e]^T) for the
 *   rows pos of every local expert e; w13 [num_local, 2*inter, hidden] (gate rows, then up rows);
 *   hidden % 64 == 0, inter % 64 == 0.
 * b200_moe_gemm_down: y[pos, :hidden] = (h[pos] . w2[e]^T in dtype) * topk_weights[sorted_pair[pos]];
 *   w2 [num_local, hidden, inter]; inter % 64 == 0, hidden % 128 == 0.
 * b200_moe_combine: out[t] = sum of y[pos_of_pair[t, j]] over t's local pairs in ascending expert
 *   order, rounded to dtype after every add (index_add_ semantics); zeros when none is local.
 * ------------------------------------------------------------------------------- */
int b200_moe_route(int32_t* topk_ids, void* topk_weights, const void* router_logits, int32_t tokens,
                   int32_t num_experts, int32_t top_k, int64_t logits_stride, int32_t dtype,
                   b200_stream_t stream);
int b200_moe_align(int32_t* offsets, int32_t* sorted_pair, int32_t* pos_of_pair, int32_t* counts,
                   const int32_t* topk_ids, int32_t num_pairs, int32_t expert_begin,
                   int32_t num_local, b200_stream_t stream);
int b200_moe_gemm_gate_up(void* h, const void* x, const void* w13, const int32_t* offsets,
                          const int32_t* sorted_pair, int32_t hidden, int32_t inter,
                          int32_t num_local, int32_t top_k, int64_t x_stride, int32_t dtype,
                          b200_stream_t stream);
int b200_moe_gemm_down(void* y, const void* h, const void* w2, const int32_t* offsets,
                       const int32_t* sorted_pair, const void* topk_weights, int32_t inter,
                       int32_t hidden, int32_t num_local, int32_t top_k, int32_t dtype,
                       b200_stream_t stream);
int b200_moe_combine(void* out, const void* y, const int32_t* topk_ids, const int32_t* pos_of_pair,
                     int32_t tokens, int32_t top_k, int32_t hidden, int64_t out_stride,
                     int32_t dtype, b200_stream_t stream);

/* MoE expert GEMMs for prefill-sized batches on tcgen05 / TMEM (csrc/moe_umma.cu; staged). Same plan,
 * rounding points and outputs as b200_moe_gemm_gate_up / _down, for hundreds of rows per expert:
 * 128-row x 128-weight-row tiles of one expert, TMA-staged SW128 panels, accumulators in TMEM.
 * b200_moe_gather_rows writes xs[pos, :hidden] = x[sorted_pair[pos] / top_k] for pos < offsets[num_local]
 * (TMA needs contiguous A rows); xs / h / y have `pairs` = tokens * top_k rows of capacity.
 * hidden % 64 == 0, inter % 64 == 0; down additionally hidden % 128 == 0.
 * Dense use: with num_local = 1 and offsets = {0, M} the two GEMMs are ordinary linear layers over M rows
 * -- gate_up_umma = SiluAndMul(x . W_gate_up^T) in one kernel (layers/activation.py:11-13 fused into the
 * GEMM of models/qwen3.py:115-117), down_umma with sorted_pair = topk_weights = NULL = x . W^T. */
int b200_moe_gather_rows(void* xs, const void* x, const int32_t* offsets, const int32_t* sorted_pair,
                         int32_t pairs, int32_t num_local, int32_t top_k, int32_t hidden,
                         int64_t x_stride, int32_t dtype, b200_stream_t stream);
int b200_moe_gemm_gate_up_umma(void* h, const void* xs, const void* w13, const int32_t* offsets,
                               int32_t pairs, int32_t hidden, int32_t inter, int32_t num_local,
                               int32_t dtype, b200_stream_t stream);
int b200_moe_gemm_down_umma(void* y, const void* h, const void* w2, const int32_t* offsets,
                            const int32_t* sorted_pair, const void* topk_weights, int32_t pairs,
                            int32_t inter, int32_t hidden, int32_t num_local, int32_t dtype,
                            b200_stream_t stream);

/* ---------------------------------------------------------------------------------
 * C1 (fused)  tensor-parallel all-reduce + residual add + RMSNorm in one kernel over NVLink peer
 * memory, for decode-sized messages.
 * Replaces: dist.all_reduce(y) in RowParallelLinear.forward            layers/linear.py:151-152
 *           followed by RMSNorm.add_rms_forward                        layers/layernorm.py:27-38
 *           (call chain models/qwen3.py:89,118 -> :155-159).
 * Every rank's row-parallel GEMM writes its partial [rows, hidden] into a buffer mapped by all
 * ranks (b200_comm_* below). peer_bufs[r] / peer_flags[r] (HOST arrays of `world` device
 * pointers) are rank r's partial buffer and flag array (>= 8 uint32, zero-initialised) as
 * addressable from this rank; `state` is a local zero-initialised uint32[4]. All ranks must issue
 * the same sequence of calls; consecutive calls must alternate between two partial buffers.
 * Result: residual <- cast(sum_r partial_r) + residual ; out <- RMSNorm(that sum) * weight, with
 * the rounding points of the reference (sum materialised in dtype, add in fp32).
 * ------------------------------------------------------------------------------- */
int b200_allreduce_add_rmsnorm(void* out, void* residual, const void* weight, float eps,
                               int32_t rows, int32_t hidden, const void* const* peer_bufs,
                               void* const* peer_flags, void* state, int32_t world, int32_t rank,
                               int32_t dtype, b200_stream_t stream);

/* C1 (fused), two-shot form for messages whose one-shot traffic ((world-1) full copies pulled by
 * every rank) is bandwidth-bound. Same contract and result bits as b200_allreduce_add_rmsnorm; in
 * addition peer_gather[r] is rank r's gather buffer [rows, hidden] (dtype) as addressable from this
 * rank, peer_flags[r] must hold >= 16 uint32 (the first 8 are shared with the one-shot form, so both
 * forms may be mixed call by call as long as every rank makes the same choice), `state` is the same
 * local uint32[4]. Rank r reduces rows [r*ceil(rows/world), ...) and pushes the sums to every rank;
 * NVLink bytes per rank: 2*(world-1)/world * rows*hidden*sizeof(dtype). The grid is capped at one
 * CTA per SM (every CTA spins on peer flags and must be co-resident). */
int b200_allreduce_add_rmsnorm_twoshot(void* out, void* residual, const void* weight, float eps,
                                       int32_t rows, int32_t hidden, const void* const* peer_bufs,
                                       void* const* peer_gather, void* const* peer_flags,
                                       void* state, int32_t world, int32_t rank, int32_t dtype,
                                       b200_stream_t stream);

/* C1 (fused), NVLS / multicast form of the two-shot kernel: the partial and gather buffers of all
 * ranks are bound to one NVSwitch multicast object (set up by the caller -- the Python side uses
 * torch.distributed._symmetric_memory for the allocation and the cuMulticast* binding). `mc_partial`
 * and `mc_gather` are the MULTICAST addresses of the partial buffer this call reduces and of the
 * gather buffer; `local_gather` is this rank's own (unicast) gather buffer. Rank r issues one
 * multimem.ld_reduce (in-switch sum, fp32 accumulation, rounded once to dtype) and one multimem.st
 * (in-switch broadcast) per 16 bytes of its row slice, then every rank runs the residual add +
 * RMSNorm locally. Flags/state: as the two-shot form (same epoch sequence; forms may be mixed call
 * by call). The in-switch summation order is unspecified: results agree with the other forms to
 * one rounding of dtype, and are identical on every rank. */
int b200_allreduce_add_rmsnorm_nvls(void* out, void* residual, const void* weight, float eps,
                                    int32_t rows, int32_t hidden, const void* mc_partial,
                                    void* mc_gather, const void* local_gather,
                                    void* const* peer_flags, void* state, int32_t world,
                                    int32_t rank, int32_t dtype, b200_stream_t stream);

/* C1 (fused), NVLS with a row-sliced ("sequence-parallel") residual: rank r owns rows
 * [r*ceil(rows/world), ...). For its rows it takes the in-switch sum of the partials, adds ITS
 * residual rows, stores the new residual rows locally and multicasts the normalised rows into every
 * rank's `out` region (`mc_out` = multicast address of a [rows, hidden] region of the symmetric
 * allocation; read it afterwards through the local unicast address). One pass, two flag barriers, the
 * residual never crosses a link. After the call `residual` is current ONLY in the caller's own row
 * slice -- sufficient for a chain of such calls (each needs a residual row only on its owner), so all
 * add+RMSNorm steps of one forward pass must use this form together. Flags/state as the other forms. */
int b200_allreduce_add_rmsnorm_nvls_fused(void* residual, const void* weight, float eps, int32_t rows,
                                          int32_t hidden, const void* mc_partial, void* mc_out,
                                          void* const* peer_flags, void* state, int32_t world,
                                          int32_t rank, int32_t dtype, b200_stream_t stream);

/* NVLS one-shot form: every rank pulls the in-switch sum of EVERY row (one multimem.ld_reduce per 16
 * bytes; this rank's links carry 1 x the message instead of (world-1) x), adds the residual and
 * normalises locally. One flag barrier per call, no row slicing (any rows <= 1184). */
int b200_allreduce_add_rmsnorm_nvls_oneshot(void* out, void* residual, const void* weight, float eps,
                                            int32_t rows, int32_t hidden, const void* mc_partial,
                                            void* const* peer_flags, void* state, int32_t world,
                                            int32_t rank, int32_t dtype, b200_stream_t stream);

/* Set-up helpers for the peer buffers (NOT hot path: they allocate / map memory, synchronously).
 * alloc: cudaMalloc + zero; get/open_handle: CUDA IPC (64-byte handle, peer access enabled
 * lazily); all return 0 or a cudaError_t value. */
int b200_comm_alloc(size_t bytes, void** ptr);
int b200_comm_free(void* ptr);
int b200_comm_get_handle(const void* ptr, void* handle64);
int b200_comm_open_handle(const void* handle64, void** ptr);
int b200_comm_close_handle(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* B200ATTN_H_ */
