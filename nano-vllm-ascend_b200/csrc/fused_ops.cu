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
  pdl_wait();
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
    launch_pdl(rmsnorm_kernel<__nv_bfloat16, kAdd>, dim3(rows), dim3(threads), 0, stream,
               (__nv_bfloat16*)out, (const __nv_bfloat16*)x, (__nv_bfloat16*)residual,
               (const __nv_bfloat16*)weight, eps, hidden);
  else
    launch_pdl(rmsnorm_kernel<__half, kAdd>, dim3(rows), dim3(threads), 0, stream, (__half*)out,
               (const __half*)x, (__half*)residual, (const __half*)weight, eps, hidden);
  return finish_launch();
}

// ------------------------------------------------------------------ SiluAndMul
template <typename T>
__global__ void __launch_bounds__(256)
silu_and_mul_kernel(uint4* __restrict__ out, const uint4* __restrict__ x, int64_t rows,
                    int inter_vec) {
  pdl_wait();
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
      // dtype multiply with the up projection. The fp32 part uses the fast exp / divide
      // intrinsics (~2 ulp fp32, invisible after the rounding to 8/11 mantissa bits except for
      // a ~1e-4 fraction of 1-ulp flips): with IEEE expf + division the kernel was issue-bound
      // at 4.2 TB/s.
      float s0 = to_f32<T>(from_f32<T>(__fdividef(a.x, 1.0f + __expf(-a.x))));
      float s1 = to_f32<T>(from_f32<T>(__fdividef(a.y, 1.0f + __expf(-a.y))));
      ow[j] = pack2<T>(__fmul_rn(s0, b.x), __fmul_rn(s1, b.y));
    }
    out[idx] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

// ------------------------------------------------------------------ K8 qk-norm + RoPE + scatter
// One CTA (8 warps) per token and per group of 8*HPW head rows; LPH = D/16 lanes share a head row:
// lane `sub` owns elements [8 sub, 8 sub + 8) of the first half and the same offsets of the
// second half -- exactly the (x1[i], x2[i]) pairs the NeoX rotation couples -- so every access is
// a 16-byte vector and a warp moves 4 (D=128) or 8 (D=64) head rows per pass.
template <typename T, int D>
__global__ void __launch_bounds__(256)
qknorm_rope_cache_kernel(T* __restrict__ qkv, const int64_t* __restrict__ positions,
                         const float* __restrict__ cos_sin, const T* __restrict__ qw,
                         const T* __restrict__ kw, float eps, T* __restrict__ k_cache,
                         T* __restrict__ v_cache, const int32_t* __restrict__ slots,
                         int slot_is_2d, int H, int Hkv, int64_t qkv_stride, int max_position,
                         int num_blocks, int block_size) {
  constexpr int HALF = D / 2;
  constexpr int LPH = D / 16;     // lanes per head row
  constexpr int HPW = 32 / LPH;   // head rows per warp
  pdl_wait();
  const int token = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane % LPH;
  const int hidx = (blockIdx.y * 8 + warp) * HPW + lane / LPH;
  const bool active = hidx < H + 2 * Hkv;
  T* row = qkv + (int64_t)token * qkv_stride + (int64_t)(active ? hidx : 0) * D;

  // destination slot in the paged cache (k and v heads only)
  int64_t slot = -1;
  if (active && hidx >= H && k_cache != nullptr) {
    if (slot_is_2d) {
      const int blk = slots[2 * token], off = slots[2 * token + 1];
      if (blk >= 0 && blk < num_blocks && off >= 0 && off < block_size)
        slot = (int64_t)blk * block_size + off;
    } else {
      const int s = slots[token];
      if (s >= 0 && s < (int64_t)num_blocks * block_size) slot = s;
    }
  }
  uint4 a = make_uint4(0, 0, 0, 0), b = make_uint4(0, 0, 0, 0);
  if (active) {
    a = *reinterpret_cast<const uint4*>(row + 8 * sub);
    b = *reinterpret_cast<const uint4*>(row + HALF + 8 * sub);
  }
  const bool is_v = hidx >= H + Hkv;
  if (active && is_v) {  // ---- V head: plain copy into the cache
    if (slot >= 0) {
      T* dst = v_cache + (slot * Hkv + (hidx - H - Hkv)) * D;
      *reinterpret_cast<uint4*>(dst + 8 * sub) = a;
      *reinterpret_cast<uint4*>(dst + HALF + 8 * sub) = b;
    }
  }
  const bool is_q = hidx < H;
  const T* w = is_q ? qw : kw;
  float x1[8], x2[8];
  {
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack2<T>(aw[j]), g = unpack2<T>(bw[j]);
      x1[2 * j] = f.x, x1[2 * j + 1] = f.y, x2[2 * j] = g.x, x2[2 * j + 1] = g.y;
    }
  }
  if (w != nullptr) {  // per-head RMSNorm (layers/layernorm.py:16-25 applied to [T,H,D]); warp-uniform
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) ss = fmaf(x1[i], x1[i], fmaf(x2[i], x2[i], ss));
#pragma unroll
    for (int o = 1; o < LPH; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = rsqrt_exact(__fadd_rn(__fdiv_rn(ss, (float)D), eps));
    const uint4 wa = *reinterpret_cast<const uint4*>(w + 8 * sub);
    const uint4 wb = *reinterpret_cast<const uint4*>(w + HALF + 8 * sub);
    const uint32_t waw[4] = {wa.x, wa.y, wa.z, wa.w}, wbw[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 w1 = unpack2<T>(waw[j]), w2 = unpack2<T>(wbw[j]);
      const float w1v[2] = {w1.x, w1.y}, w2v[2] = {w2.x, w2.y};
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float n1 = to_f32<T>(from_f32<T>(__fmul_rn(x1[2 * j + e], inv)));
        const float n2 = to_f32<T>(from_f32<T>(__fmul_rn(x2[2 * j + e], inv)));
        x1[2 * j + e] = to_f32<T>(from_f32<T>(__fmul_rn(n1, w1v[e])));
        x2[2 * j + e] = to_f32<T>(from_f32<T>(__fmul_rn(n2, w2v[e])));
      }
    }
  }
  if (!active || is_v) return;
  // NeoX rotation in fp32 (rotary_embedding.py:6-14): y1 = x1*cos - x2*sin, y2 = x2*cos + x1*sin
  int64_t pos = positions[token];
  if (pos < 0) pos = 0;
  if (pos >= max_position) pos = max_position - 1;
  const float* cs = cos_sin + pos * D + 8 * sub;
  const float4 c0 = *reinterpret_cast<const float4*>(cs), c1 = *reinterpret_cast<const float4*>(cs + 4);
  const float4 s0 = *reinterpret_cast<const float4*>(cs + HALF),
               s1 = *reinterpret_cast<const float4*>(cs + HALF + 4);
  const float cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
  const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
  uint32_t o1[4], o2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float y1[2], y2[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int i = 2 * j + e;
      y1[e] = __fsub_rn(__fmul_rn(x1[i], cv[i]), __fmul_rn(x2[i], sv[i]));
      y2[e] = __fadd_rn(__fmul_rn(x2[i], cv[i]), __fmul_rn(x1[i], sv[i]));
    }
    o1[j] = pack2<T>(y1[0], y1[1]);
    o2[j] = pack2<T>(y2[0], y2[1]);
  }
  const uint4 r1 = make_uint4(o1[0], o1[1], o1[2], o1[3]), r2 = make_uint4(o2[0], o2[1], o2[2], o2[3]);
  *reinterpret_cast<uint4*>(row + 8 * sub) = r1;
  *reinterpret_cast<uint4*>(row + HALF + 8 * sub) = r2;
  if (!is_q && slot >= 0) {
    T* kdst = k_cache + (slot * Hkv + (hidx - H)) * D;
    *reinterpret_cast<uint4*>(kdst + 8 * sub) = r1;
    *reinterpret_cast<uint4*>(kdst + HALF + 8 * sub) = r2;
  }
}

template <typename T>
static int qknorm_launch(void* qkv, const int64_t* positions, const float* cos_sin, const void* qw,
                         const void* kw, float eps, void* k_cache, void* v_cache,
                         const int32_t* slots, int slot_is_2d, int T_, int H, int Hkv, int D,
                         int64_t qkv_stride, int max_position, int num_blocks, int block_size,
                         cudaStream_t stream) {
  const int rows_per_cta = 8 * (32 / (D / 16));
  dim3 grid(T_, (H + 2 * Hkv + rows_per_cta - 1) / rows_per_cta);
  if (D == 128)
    launch_pdl(qknorm_rope_cache_kernel<T, 128>, grid, dim3(256), 0, stream, (T*)qkv, positions,
               cos_sin, (const T*)qw, (const T*)kw, eps, (T*)k_cache, (T*)v_cache, slots, slot_is_2d,
               H, Hkv, qkv_stride, max_position, num_blocks, block_size);
  else
    launch_pdl(qknorm_rope_cache_kernel<T, 64>, grid, dim3(256), 0, stream, (T*)qkv, positions,
               cos_sin, (const T*)qw, (const T*)kw, eps, (T*)k_cache, (T*)v_cache, slots, slot_is_2d,
               H, Hkv, qkv_stride, max_position, num_blocks, block_size);
  return finish_launch();
}

// ------------------------------------------------------------------ sampling tail (f2)
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// 32-bit finaliser (murmur3 fmix32): one draw per vocabulary entry
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}

// One CTA per row, 16-byte loads. score_i = logit_i / T + Gumbel_i (T > 0) or logit_i (greedy);
// winner = argmax score, ties -> lowest index (torch.argmax semantics on a first-max basis).
template <typename T>
__global__ void __launch_bounds__(1024)
sample_kernel(int64_t* __restrict__ out, const T* __restrict__ logits,
              const float* __restrict__ temperatures, int vocab, int64_t stride, uint64_t seed,
              uint64_t step, const int64_t* __restrict__ step_offset) {
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  pdl_wait();
  const int row = blockIdx.x;
  const T* lrow = logits + (int64_t)row * stride;
  const float temp = temperatures == nullptr ? 0.f : temperatures[row];
  const bool greedy = !(temp > 0.f);   // no temperatures, or a non-positive one => argmax
  const float inv_t = greedy ? 1.f : 1.f / temp;
  if (step_offset != nullptr) step += (uint64_t)*step_offset;
  const uint64_t key64 = splitmix64(seed ^ splitmix64(step * 0x100000001B3ull + row));
  const uint32_t key = (uint32_t)key64 ^ (uint32_t)(key64 >> 32);
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  auto consider = [&](float v, int i) {
    if (!greedy) {
      const uint32_t r = mix32(key + (uint32_t)i * 0x9E3779B1u);
      const float u = ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0,1)
      // Gumbel(0,1) = -ln(-ln u)
      v = fmaf(v, inv_t, -0.6931471806f * __log2f(-0.6931471806f * __log2f(u)));
    }
    if (v > best || (v == best && i < best_i)) best = v, best_i = i;
  };
  const bool vec_ok = (stride % 8 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  const int nvec = vec_ok ? vocab / 8 : 0;
  for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
    const uint4 q = ld_nc_v4(reinterpret_cast<const uint4*>(lrow) + vi);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack2<T>(w[j]);
      consider(f.x, vi * 8 + 2 * j);
      consider(f.y, vi * 8 + 2 * j + 1);
    }
  }
  for (int i = nvec * 8 + threadIdx.x; i < vocab; i += blockDim.x) consider(to_f32<T>(lrow[i]), i);
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

// Vocab-sharded form of sample_kernel for tensor parallelism (replaces the logits gather of
// layers/embed_head.py:62-65 + a full-vocabulary pass on rank 0): every rank scores ITS shard of
// the vocabulary -- the Gumbel draw is a pure function of (seed, step, row, GLOBAL index), so the
// shards together see exactly the scores of the unsharded kernel -- and emits per row one int64 key
//   key = order_preserving_u32(best score) << 31  |  (0x7fffffff - global index of the winner)
// so that a MAX all-reduce over the ranks (8 bytes per row) yields the global arg-max with the same
// tie rule (lowest index); b200_sample_shard's caller decodes id = 0x7fffffff - (key & 0x7fffffff).
template <typename T>
__global__ void __launch_bounds__(1024)
sample_shard_kernel(int64_t* __restrict__ out_keys, const T* __restrict__ logits,
                    const float* __restrict__ temperatures, int vocab, int64_t stride, int index_offset,
                    uint64_t seed, uint64_t step, const int64_t* __restrict__ step_offset) {
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  pdl_wait();
  const int row = blockIdx.x;
  const T* lrow = logits + (int64_t)row * stride;
  const float temp = temperatures == nullptr ? 0.f : temperatures[row];
  const bool greedy = !(temp > 0.f);
  const float inv_t = greedy ? 1.f : 1.f / temp;
  if (step_offset != nullptr) step += (uint64_t)*step_offset;
  const uint64_t key64 = splitmix64(seed ^ splitmix64(step * 0x100000001B3ull + row));
  const uint32_t key = (uint32_t)key64 ^ (uint32_t)(key64 >> 32);
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  auto consider = [&](float v, int local) {
    const int i = local + index_offset;                  // global vocabulary index
    if (!greedy) {
      const uint32_t r = mix32(key + (uint32_t)i * 0x9E3779B1u);
      const float u = ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f);
      v = fmaf(v, inv_t, -0.6931471806f * __log2f(-0.6931471806f * __log2f(u)));
    }
    if (v > best || (v == best && i < best_i)) best = v, best_i = i;
  };
  const bool vec_ok = (stride % 8 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  const int nvec = vec_ok ? vocab / 8 : 0;
  for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
    const uint4 q = ld_nc_v4(reinterpret_cast<const uint4*>(lrow) + vi);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack2<T>(w[j]);
      consider(f.x, vi * 8 + 2 * j);
      consider(f.y, vi * 8 + 2 * j + 1);
    }
  }
  for (int i = nvec * 8 + threadIdx.x; i < vocab; i += blockDim.x) consider(to_f32<T>(lrow[i]), i);
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
    if (lane == 0) {
      if (best == 0.f) best = 0.f;                        // -0.0 and +0.0 compare equal: one key for both
      uint32_t b = __float_as_uint(best);
      b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);    // order-preserving map float -> u32
      if (best_i == 0x7fffffff) best_i = index_offset;    // empty / all -inf shard: a valid id, lowest key
      out_keys[row] = ((int64_t)b << 31) | (int64_t)(0x7fffffff - best_i);
    }
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
    launch_pdl(silu_and_mul_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream,
               (uint4*)out, (const uint4*)x, (int64_t)rows, inter / 8);
  else
    launch_pdl(silu_and_mul_kernel<__half>, dim3(grid), dim3(256), 0, (cudaStream_t)stream,
               (uint4*)out, (const uint4*)x, (int64_t)rows, inter / 8);
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
                     (v_cache && !aligned16(v_cache)) || !aligned16(cos_sin_cache) ||
                     (q_norm_weight && (!aligned16(q_norm_weight) || !aligned16(k_norm_weight))),
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
    launch_pdl(sample_kernel<__nv_bfloat16>, dim3(rows), dim3(1024), 0, (cudaStream_t)stream,
               out_tokens, (const __nv_bfloat16*)logits, temperatures, vocab, logits_stride, seed, step,
               step_offset);
  else
    launch_pdl(sample_kernel<__half>, dim3(rows), dim3(1024), 0, (cudaStream_t)stream, out_tokens,
               (const __half*)logits, temperatures, vocab, logits_stride, seed, step, step_offset);
  return finish_launch();
}

extern "C" int b200_sample_shard(int64_t* out_keys, const void* logits, const float* temperatures,
                                 int32_t rows, int32_t vocab_local, int64_t logits_stride,
                                 int32_t index_offset, uint64_t seed, uint64_t step,
                                 const int64_t* step_offset, int32_t dtype, b200_stream_t stream) {
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(rows < 0 || vocab_local <= 0 || logits_stride < vocab_local || index_offset < 0 ||
                     (int64_t)index_offset + vocab_local > 0x7fffffff,
                 B200_ERR_BAD_SHAPE);
  if (rows == 0) return B200_OK;
  B200_RETURN_IF(!out_keys || !logits, B200_ERR_NULL_POINTER);
  if (dtype == B200_DTYPE_BF16)
    launch_pdl(sample_shard_kernel<__nv_bfloat16>, dim3(rows), dim3(1024), 0, (cudaStream_t)stream, out_keys,
               (const __nv_bfloat16*)logits, temperatures, vocab_local, logits_stride, index_offset, seed,
               step, step_offset);
  else
    launch_pdl(sample_shard_kernel<__half>, dim3(rows), dim3(1024), 0, (cudaStream_t)stream, out_keys,
               (const __half*)logits, temperatures, vocab_local, logits_stride, index_offset, seed, step,
               step_offset);
  return finish_launch();
}
