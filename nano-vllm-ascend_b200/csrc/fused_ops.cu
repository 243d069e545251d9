// fused_ops.cu -- the elementwise/normalisation epilogues around the attention path.
//
//   rmsnorm / add_rmsnorm   <- layers/layernorm.py:16-38 of the reference
//   silu_and_mul            <- layers/activation.py:11-13
//   qknorm_rope_cache (K8)  <- the chain models/qwen3.py:79-88: split, q_norm/k_norm
//                              (layernorm.py:16-25), NeoX RoPE (rotary_embedding.py:6-14),
//                              KV scatter (attention.py:22-35)
//   sample                  <- layers/sampler.py:9-17 (argmax / Gumbel-max tail)
//
// All HBM-bound, one pass over the data, 16-byte vector accesses. The arithmetic follows
// the reference's eager op sequence step by step (fp32 internal math, every intermediate
// that the reference materialises in bf16/fp16 is rounded at the same point, products and
// sums that torch evaluates as separate ops are kept un-fused with __fmul_rn/__fadd_rn).
#include "common.cuh"

namespace b200 {

// fp32 1/sqrt(x): torch's CPU rsqrt is an exact sqrt followed by a division
__device__ __forceinline__ float rsqrt_exact(float x) { return __fdiv_rn(1.0f, __fsqrt_rn(x)); }

__device__ __forceinline__ float block_sum(float v, float* red /*[32]*/) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (nw == 1) return v;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  __syncthreads();
  return t;
}

// ------------------------------------------------------------------ RMSNorm (+ residual add)
// One CTA per row. Each thread keeps its slice of the row in registers (<= kMaxVec uint4).
constexpr int kMaxVec = 4;

template <typename T, bool kAdd>
__global__ void __launch_bounds__(256)
rmsnorm_kernel(T* __restrict__ out, const T* __restrict__ x, T* __restrict__ residual,
               const T* __restrict__ weight, float eps, int hidden) {
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const int nvec = hidden / 8;
  const uint4* xrow = reinterpret_cast<const uint4*>(x + row * hidden);
  uint4* rrow = kAdd ? reinterpret_cast<uint4*>(residual + row * hidden) : nullptr;
  float vals[kMaxVec][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int v = threadIdx.x + i * blockDim.x;
    if (v < nvec) {
      uint4 a = xrow[v];
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
      uint32_t rw[4] = {0, 0, 0, 0};
      if (kAdd) {
        uint4 r = rrow[v];
        rw[0] = r.x, rw[1] = r.y, rw[2] = r.z, rw[3] = r.w;
      }
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = unpack2<T>(aw[j]);
        if (kAdd) {
          float2 g = unpack2<T>(rw[j]);
          f.x = __fadd_rn(f.x, g.x);
          f.y = __fadd_rn(f.y, g.y);
          ow[j] = pack2<T>(f.x, f.y);  // residual' = cast(x + residual)
        }
        vals[i][2 * j] = f.x;
        vals[i][2 * j + 1] = f.y;
        ss = fmaf(f.x, f.x, ss);
        ss = fmaf(f.y, f.y, ss);
      }
      if (kAdd) rrow[v] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
  }
  ss = block_sum(ss, red);
  const float inv = rsqrt_exact(__fadd_rn(__fdiv_rn(ss, (float)hidden), eps));
  uint4* orow = reinterpret_cast<uint4*>(out + row * hidden);
  const uint4* wrow = reinterpret_cast<const uint4*>(weight);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int v = threadIdx.x + i * blockDim.x;
    if (v < nvec) {
      uint4 w = wrow[v];
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 wf = unpack2<T>(ww[j]);
        // x.to(dtype).mul_(weight): round the normalised value first, then a dtype multiply
        float a = to_f32<T>(from_f32<T>(__fmul_rn(vals[i][2 * j], inv)));
        float b = to_f32<T>(from_f32<T>(__fmul_rn(vals[i][2 * j + 1], inv)));
        ow[j] = pack2<T>(__fmul_rn(a, wf.x), __fmul_rn(b, wf.y));
      }
      orow[v] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
  }
}

template <bool kAdd>
static int rmsnorm_launch(void* out, const void* x, void* residual, const void* weight, float eps,
                          int32_t rows, int32_t hidden, int32_t dtype, cudaStream_t stream) {
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(rows < 0 || hidden <= 0, B200_ERR_BAD_SHAPE);
  if (rows == 0) return B200_OK;
  B200_RETURN_IF(!out || !x || !weight || (kAdd && !residual), B200_ERR_NULL_POINTER);
  B200_RETURN_IF(hidden % 8 != 0, B200_ERR_MISALIGNED);
  B200_RETURN_IF(hidden > 8 * 256 * kMaxVec, B200_ERR_UNSUPPORTED);
  B200_RETURN_IF(!aligned16(out) || !aligned16(x) || !aligned16(weight) ||
                     (kAdd && !aligned16(residual)),
                 B200_ERR_MISALIGNED);
  const int nvec = hidden / 8;
  int threads = (nvec + kMaxVec - 1) / kMaxVec;
  threads = ((threads + 31) / 32) * 32;
  if (threads < 128 && nvec >= 128) threads = 128;
  if (threads > 256) threads = 256;
  if (dtype == B200_DTYPE_BF16)
    rmsnorm_kernel<__nv_bfloat16, kAdd><<<rows, threads, 0, stream>>>(
        (__nv_bfloat16*)out, (const __nv_bfloat16*)x, (__nv_bfloat16*)residual,
        (const __nv_bfloat16*)weight, eps, hidden);
  else
    rmsnorm_kernel<__half, kAdd><<<rows, threads, 0, stream>>>(
        (__half*)out, (const __half*)x, (__half*)residual, (const __half*)weight, eps, hidden);
  return finish_launch();
}

// ------------------------------------------------------------------ SiluAndMul
template <typename T>
__global__ void __launch_bounds__(256)
silu_and_mul_kernel(uint4* __restrict__ out, const uint4* __restrict__ x, int64_t rows,
                    int inter_vec) {
  const int64_t total = rows * inter_vec;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / inter_vec;
    const int c = (int)(idx - r * inter_vec);
    uint4 g = ld_nc_v4(x + r * 2 * inter_vec + c);
    uint4 u = ld_nc_v4(x + r * 2 * inter_vec + inter_vec + c);
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, uw[4] = {u.x, u.y, u.z, u.w};
    uint32_t ow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 a = unpack2<T>(gw[j]), b = unpack2<T>(uw[j]);
      // F.silu in dtype: x / (1 + exp(-x)) evaluated in fp32, rounded to dtype, then a
      // dtype multiply with the up projection.
      float s0 = to_f32<T>(from_f32<T>(__fdiv_rn(a.x, __fadd_rn(1.0f, expf(-a.x)))));
      float s1 = to_f32<T>(from_f32<T>(__fdiv_rn(a.y, __fadd_rn(1.0f, expf(-a.y)))));
      ow[j] = pack2<T>(__fmul_rn(s0, b.x), __fmul_rn(s1, b.y));
    }
    out[idx] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

// ------------------------------------------------------------------ K8 qk-norm + RoPE + scatter
// grid = (num_tokens, ceil((H + 2*Hkv) / 8)), 8 warps per CTA, one warp per head row.
// Lane l owns elements {EPL*l .. EPL*l+EPL-1} of the first half and the same offsets in the
// second half (EPL = D/64), i.e. exactly the (x1[i], x2[i]) pairs NeoX rotation couples.
template <typename T, int D>
__global__ void __launch_bounds__(256)
qknorm_rope_cache_kernel(T* __restrict__ qkv, const int64_t* __restrict__ positions,
                         const float* __restrict__ cos_sin, const T* __restrict__ qw,
                         const T* __restrict__ kw, float eps, T* __restrict__ k_cache,
                         T* __restrict__ v_cache, const int32_t* __restrict__ slots,
                         int slot_is_2d, int H, int Hkv, int64_t qkv_stride, int max_position,
                         int num_blocks, int block_size) {
  constexpr int EPL = D / 64;  // elements per lane per half: 2 (D=128) or 1 (D=64)
  constexpr int HALF = D / 2;
  const int token = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hidx = blockIdx.y * 8 + warp;
  if (hidx >= H + 2 * Hkv) return;
  T* row = qkv + (int64_t)token * qkv_stride + (int64_t)hidx * D;

  // destination slot in the paged cache (k and v heads only)
  int64_t slot = -1;
  if (hidx >= H && k_cache != nullptr) {
    if (slot_is_2d) {
      const int blk = slots[2 * token], off = slots[2 * token + 1];
      if (blk >= 0 && blk < num_blocks && off >= 0 && off < block_size)
        slot = (int64_t)blk * block_size + off;
    } else {
      const int s = slots[token];
      if (s >= 0 && s < (int64_t)num_blocks * block_size) slot = s;
    }
  }

  if (hidx >= H + Hkv) {  // ---- V head: plain copy into the cache
    if (slot >= 0) {
      const int vh = hidx - H - Hkv;
      T* dst = v_cache + (slot * Hkv + vh) * D;
      if (EPL == 2) {
        reinterpret_cast<uint2*>(dst)[lane] = reinterpret_cast<const uint2*>(row)[lane];
      } else {
        reinterpret_cast<uint32_t*>(dst)[lane] = reinterpret_cast<const uint32_t*>(row)[lane];
      }
    }
    return;
  }

  const bool is_q = hidx < H;
  const T* w = is_q ? qw : kw;
  float x1[EPL], x2[EPL];
  if (EPL == 2) {
    float2 a = unpack2<T>(reinterpret_cast<const uint32_t*>(row)[lane]);
    float2 b = unpack2<T>(reinterpret_cast<const uint32_t*>(row + HALF)[lane]);
    x1[0] = a.x, x1[EPL - 1] = a.y, x2[0] = b.x, x2[EPL - 1] = b.y;
  } else {
    x1[0] = to_f32<T>(row[lane]);
    x2[0] = to_f32<T>(row[HALF + lane]);
  }
  if (w != nullptr) {  // per-head RMSNorm (layers/layernorm.py:16-25 applied to [T,H,D])
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) ss = fmaf(x1[i], x1[i], fmaf(x2[i], x2[i], ss));
    ss = warp_sum(ss);
    const float inv = rsqrt_exact(__fadd_rn(__fdiv_rn(ss, (float)D), eps));
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const float w1 = to_f32<T>(w[EPL * lane + i]), w2 = to_f32<T>(w[HALF + EPL * lane + i]);
      const float n1 = to_f32<T>(from_f32<T>(__fmul_rn(x1[i], inv)));
      const float n2 = to_f32<T>(from_f32<T>(__fmul_rn(x2[i], inv)));
      x1[i] = to_f32<T>(from_f32<T>(__fmul_rn(n1, w1)));
      x2[i] = to_f32<T>(from_f32<T>(__fmul_rn(n2, w2)));
    }
  }
  // NeoX rotation in fp32 (rotary_embedding.py:6-14): y1 = x1*cos - x2*sin, y2 = x2*cos + x1*sin
  int64_t pos = positions[token];
  if (pos < 0) pos = 0;
  if (pos >= max_position) pos = max_position - 1;
  const float* cs = cos_sin + pos * D;
  float y1[EPL], y2[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    const float c = cs[EPL * lane + i], s = cs[HALF + EPL * lane + i];
    y1[i] = __fsub_rn(__fmul_rn(x1[i], c), __fmul_rn(x2[i], s));
    y2[i] = __fadd_rn(__fmul_rn(x2[i], c), __fmul_rn(x1[i], s));
  }
  T* kdst = nullptr;
  if (!is_q && slot >= 0) kdst = k_cache + (slot * Hkv + (hidx - H)) * D;
  if (EPL == 2) {
    const uint32_t o1 = pack2<T>(y1[0], y1[EPL - 1]), o2 = pack2<T>(y2[0], y2[EPL - 1]);
    reinterpret_cast<uint32_t*>(row)[lane] = o1;
    reinterpret_cast<uint32_t*>(row + HALF)[lane] = o2;
    if (kdst) {
      reinterpret_cast<uint32_t*>(kdst)[lane] = o1;
      reinterpret_cast<uint32_t*>(kdst + HALF)[lane] = o2;
    }
  } else {
    const T o1 = from_f32<T>(y1[0]), o2 = from_f32<T>(y2[0]);
    row[lane] = o1;
    row[HALF + lane] = o2;
    if (kdst) {
      kdst[lane] = o1;
      kdst[HALF + lane] = o2;
    }
  }
}

template <typename T>
static int qknorm_launch(void* qkv, const int64_t* positions, const float* cos_sin, const void* qw,
                         const void* kw, float eps, void* k_cache, void* v_cache,
                         const int32_t* slots, int slot_is_2d, int T_, int H, int Hkv, int D,
                         int64_t qkv_stride, int max_position, int num_blocks, int block_size,
                         cudaStream_t stream) {
  dim3 grid(T_, (H + 2 * Hkv + 7) / 8);
  if (D == 128)
    qknorm_rope_cache_kernel<T, 128><<<grid, 256, 0, stream>>>(
        (T*)qkv, positions, cos_sin, (const T*)qw, (const T*)kw, eps, (T*)k_cache, (T*)v_cache,
        slots, slot_is_2d, H, Hkv, qkv_stride, max_position, num_blocks, block_size);
  else
    qknorm_rope_cache_kernel<T, 64><<<grid, 256, 0, stream>>>(
        (T*)qkv, positions, cos_sin, (const T*)qw, (const T*)kw, eps, (T*)k_cache, (T*)v_cache,
        slots, slot_is_2d, H, Hkv, qkv_stride, max_position, num_blocks, block_size);
  return finish_launch();
}

// ------------------------------------------------------------------ sampling tail (f2)
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// One CTA per row. score_i = logit_i / T + Gumbel_i (T given) or logit_i (greedy);
// winner = argmax score, ties -> lowest index (torch.argmax semantics on a first-max basis).
template <typename T>
__global__ void __launch_bounds__(1024)
sample_kernel(int64_t* __restrict__ out, const T* __restrict__ logits,
              const float* __restrict__ temperatures, int vocab, int64_t stride, uint64_t seed,
              uint64_t step, const int64_t* __restrict__ step_offset) {
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  const int row = blockIdx.x;
  const T* lrow = logits + (int64_t)row * stride;
  const float temp = temperatures == nullptr ? 0.f : temperatures[row];
  const bool greedy = !(temp > 0.f);   // no temperatures, or a non-positive one => argmax
  const float inv_t = greedy ? 1.f : 1.f / temp;
  if (step_offset != nullptr) step += (uint64_t)*step_offset;
  const uint64_t key = splitmix64(seed ^ splitmix64(step * 0x100000001B3ull + row));
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  for (int i = threadIdx.x; i < vocab; i += blockDim.x) {
    float v = to_f32<T>(lrow[i]);
    if (!greedy) {
      const uint64_t r = splitmix64(key + (uint64_t)i);
      const float u = ((float)(r >> 40) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
      v = v * inv_t - __logf(-__logf(u));
    }
    if (v > best || (v == best && i < best_i)) best = v, best_i = i;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ov > best || (ov == best && oi < best_i)) best = ov, best_i = oi;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_val[warp] = best, s_idx[warp] = best_i;
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    best = lane < nw ? s_val[lane] : -INFINITY;
    best_i = lane < nw ? s_idx[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ov > best || (ov == best && oi < best_i)) best = ov, best_i = oi;
    }
    if (lane == 0) out[row] = best_i == 0x7fffffff ? 0 : best_i;
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_rmsnorm(void* out, const void* x, const void* weight, float eps, int32_t rows,
                            int32_t hidden, int32_t dtype, b200_stream_t stream) {
  return rmsnorm_launch<false>(out, x, nullptr, weight, eps, rows, hidden, dtype,
                               (cudaStream_t)stream);
}

extern "C" int b200_add_rmsnorm(void* x, void* residual, const void* weight, float eps,
                                int32_t rows, int32_t hidden, int32_t dtype,
                                b200_stream_t stream) {
  return rmsnorm_launch<true>(x, x, residual, weight, eps, rows, hidden, dtype,
                              (cudaStream_t)stream);
}

extern "C" int b200_silu_and_mul(void* out, const void* x, int32_t rows, int32_t inter,
                                 int32_t dtype, b200_stream_t stream) {
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(rows < 0 || inter <= 0, B200_ERR_BAD_SHAPE);
  if (rows == 0) return B200_OK;
  B200_RETURN_IF(!out || !x, B200_ERR_NULL_POINTER);
  B200_RETURN_IF(inter % 8 != 0 || !aligned16(out) || !aligned16(x), B200_ERR_MISALIGNED);
  const int64_t total = (int64_t)rows * (inter / 8);
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (dtype == B200_DTYPE_BF16)
    silu_and_mul_kernel<__nv_bfloat16>
        <<<grid, 256, 0, (cudaStream_t)stream>>>((uint4*)out, (const uint4*)x, rows, inter / 8);
  else
    silu_and_mul_kernel<__half>
        <<<grid, 256, 0, (cudaStream_t)stream>>>((uint4*)out, (const uint4*)x, rows, inter / 8);
  return finish_launch();
}

extern "C" int b200_qknorm_rope_cache(void* qkv, const int64_t* positions,
                                      const float* cos_sin_cache, const void* q_norm_weight,
                                      const void* k_norm_weight, float eps, void* k_cache,
                                      void* v_cache, const int32_t* slots, int32_t slot_is_2d,
                                      int32_t num_tokens, int32_t num_heads,
                                      int32_t num_kv_heads, int32_t head_dim, int64_t qkv_stride,
                                      int32_t max_position, int32_t num_blocks,
                                      int32_t block_size, int32_t dtype, b200_stream_t stream) {
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(num_tokens < 0 || num_heads <= 0 || num_kv_heads < 0 || max_position <= 0,
                 B200_ERR_BAD_SHAPE);
  B200_RETURN_IF(head_dim != 64 && head_dim != 128, B200_ERR_UNSUPPORTED);
  if (num_tokens == 0) return B200_OK;
  B200_RETURN_IF(!qkv || !positions || !cos_sin_cache, B200_ERR_NULL_POINTER);
  B200_RETURN_IF((q_norm_weight == nullptr) != (k_norm_weight == nullptr), B200_ERR_NULL_POINTER);
  B200_RETURN_IF((k_cache == nullptr) != (v_cache == nullptr), B200_ERR_NULL_POINTER);
  B200_RETURN_IF(k_cache && (!slots || num_blocks <= 0 || block_size <= 0), B200_ERR_BAD_SHAPE);
  B200_RETURN_IF(qkv_stride < (int64_t)(num_heads + 2 * num_kv_heads) * head_dim,
                 B200_ERR_BAD_SHAPE);
  B200_RETURN_IF(qkv_stride % 8 != 0 || !aligned16(qkv) || (k_cache && !aligned16(k_cache)) ||
                     (v_cache && !aligned16(v_cache)),
                 B200_ERR_MISALIGNED);
  if (dtype == B200_DTYPE_BF16)
    return qknorm_launch<__nv_bfloat16>(qkv, positions, cos_sin_cache, q_norm_weight,
                                        k_norm_weight, eps, k_cache, v_cache, slots, slot_is_2d,
                                        num_tokens, num_heads, num_kv_heads, head_dim, qkv_stride,
                                        max_position, num_blocks, block_size,
                                        (cudaStream_t)stream);
  return qknorm_launch<__half>(qkv, positions, cos_sin_cache, q_norm_weight, k_norm_weight, eps,
                               k_cache, v_cache, slots, slot_is_2d, num_tokens, num_heads,
                               num_kv_heads, head_dim, qkv_stride, max_position, num_blocks,
                               block_size, (cudaStream_t)stream);
}

extern "C" int b200_sample(int64_t* out_tokens, const void* logits, const float* temperatures,
                           int32_t rows, int32_t vocab, int64_t logits_stride, uint64_t seed,
                           uint64_t step, const int64_t* step_offset, int32_t dtype,
                           b200_stream_t stream) {
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(rows < 0 || vocab <= 0 || logits_stride < vocab, B200_ERR_BAD_SHAPE);
  if (rows == 0) return B200_OK;
  B200_RETURN_IF(!out_tokens || !logits, B200_ERR_NULL_POINTER);
  if (dtype == B200_DTYPE_BF16)
    sample_kernel<__nv_bfloat16><<<rows, 1024, 0, (cudaStream_t)stream>>>(
        out_tokens, (const __nv_bfloat16*)logits, temperatures, vocab, logits_stride, seed, step,
        step_offset);
  else
    sample_kernel<__half><<<rows, 1024, 0, (cudaStream_t)stream>>>(
        out_tokens, (const __half*)logits, temperatures, vocab, logits_stride, seed, step, step_offset);
  return finish_launch();
}
