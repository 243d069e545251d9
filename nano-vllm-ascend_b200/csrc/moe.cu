// moe.cu -- sparse mixture-of-experts MLP: routing, dispatch plan, expert GEMMs, combine.
//
// Replaces Qwen3MoeSparseMoeBlock.forward (models/qwen3_moe.py:146-186 of the reference): a Python
// loop over the experts that were hit, each iteration gathering its tokens (torch.where + index),
// running three dense GEMMs through TP-sharded linears (one all-reduce per expert) and
// index_add_-ing the weighted result -- ~10 launches and one collective per expert per layer.
//
// Here a layer's MoE is 6 launches, no host sync, fixed-size buffers (CUDA-graph capturable):
//   moe_route_kernel    softmax over all experts (fp32) + top-k + renormalise      :150-158
//   moe_count/place     dispatch plan: (token, slot) pairs stably sorted by local expert
//   moe_gemm<gate_up>   H[pos] = SiluAndMul(X[token(pos)] . W13[e]^T)   rows gathered on the fly
//   moe_gemm<down>      Y[pos] = (H[pos] . W2[e]^T) * routing_weight     rounded as :180-182
//   moe_combine_kernel  out[t] = sum over t's experts in ascending expert order, one rounding per
//                       add (the reference's index_add_ into a model-dtype tensor, :184-186)
// Expert parallelism: a rank owns experts [expert_begin, expert_begin + num_local) WHOLE; tokens
// are replicated on every rank (they leave the attention all-reduce that way), so "dispatch" is a
// local selection and "combine" is the row-parallel SUM all-reduce the decoder layer already
// performs (fused with the next add+RMSNorm over NVLink, allreduce.cu) -- one collective per
// layer instead of one per expert, and no all-to-all.
//
// The expert GEMMs at decode are weight streaming: B tokens x top_k / E rows per expert (16 at
// B=256, k=8, E=128) against 9.4 MB of expert weights -> HBM-bound by a factor > 10, so they run
// on mma.sync with a 4-stage cp.async ring (32/64-row x 128-column tiles, K blocked by 64), sized
// to keep ~140 KB of loads in flight per SM, not on tcgen05. Algorithmic bytes per layer:
// sum over hit experts of 3 * hidden * moe_intermediate * 2 B  (+ activations, negligible).
#include "common.cuh"

namespace b200 {

constexpr int kMoeMaxTopK = 8;
constexpr int kMoeMaxExperts = 256;

__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------ routing
// One warp per token; lane l holds experts l, l+32, ... Ties go to the lower expert index.
template <typename T, int VPL>
__global__ void __launch_bounds__(256)
moe_route_kernel(int32_t* __restrict__ topk_ids, T* __restrict__ topk_w, const T* __restrict__ logits,
                 int tokens, int E, int k, int64_t stride) {
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 8 + warp;
  if (t >= tokens) return;
  const T* row = logits + (int64_t)t * stride;
  float p[VPL];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int e = lane + 32 * i;
    p[i] = e < E ? to_f32<T>(row[e]) : -INFINITY;
    mx = fmaxf(mx, p[i]);
  }
  mx = warp_max(mx);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int e = lane + 32 * i;
    p[i] = e < E ? expf(p[i] - mx) : -1.f;      // -1: never selected
    s += e < E ? p[i] : 0.f;
  }
  s = warp_sum(s);
#pragma unroll
  for (int i = 0; i < VPL; ++i)
    if (lane + 32 * i < E) p[i] = __fdiv_rn(p[i], s);
  float my_v = 0.f;
  int my_i = 0;
  for (int j = 0; j < k; ++j) {
    float bv = -1.f;
    int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      if (p[i] > bv) bv = p[i], bi = lane + 32 * i;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) bv = ov, bi = oi;
    }
    if ((bi & 31) == lane) {
#pragma unroll
      for (int i = 0; i < VPL; ++i)
        if (i == (bi >> 5)) p[i] = -1.f;
    }
    if (lane == j) my_v = bv, my_i = bi;
  }
  const float tot = warp_sum(lane < k ? my_v : 0.f);
  if (lane < k) {
    topk_ids[(int64_t)t * k + lane] = my_i;
    topk_w[(int64_t)t * k + lane] = from_f32<T>(__fdiv_rn(my_v, tot));
  }
}

// ------------------------------------------------------------------------------------ dispatch plan
// CTA e counts the pairs routed to local expert e (and resets one slice of pos_of_pair to -1).
__global__ void __launch_bounds__(256)
moe_count_kernel(int32_t* __restrict__ counts, int32_t* __restrict__ pos_of_pair,
                 const int32_t* __restrict__ ids, int P, int expert_begin) {
  __shared__ int red[8];
  pdl_wait();
  const int e = blockIdx.x, target = expert_begin + e;
  const int per = (P + gridDim.x - 1) / gridDim.x;
  const int hi = min(P, (e + 1) * per);
  for (int i = e * per + threadIdx.x; i < hi; i += blockDim.x) pos_of_pair[i] = -1;
  int c = 0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) c += ids[i] == target;
  c = warp_sum_i(c);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < 8; ++w) tot += red[w];
    counts[e] = tot;
  }
}

// CTA e places its pairs at offsets[e].. in ascending pair order (a stable sort by expert without
// atomics: each warp owns a contiguous chunk of the pair list, counts it, then ranks by ballot).
__global__ void __launch_bounds__(256)
moe_place_kernel(int32_t* __restrict__ offsets, int32_t* __restrict__ sorted_pair,
                 int32_t* __restrict__ pos_of_pair, const int32_t* __restrict__ counts,
                 const int32_t* __restrict__ ids, int P, int expert_begin, int num_local) {
  __shared__ int warp_cnt[8];
  __shared__ int s_base;
  pdl_wait();
  const int e = blockIdx.x, target = expert_begin + e;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    int b = 0;
    for (int i = lane; i < e; i += 32) b += counts[i];
    b = warp_sum_i(b);
    if (lane == 0) s_base = b;
  }
  const int chunk = (((P + 7) / 8) + 31) / 32 * 32;
  const int lo = min(P, warp * chunk), hi = min(P, lo + chunk);
  int c = 0;
  for (int i = lo + lane; i < hi; i += 32) c += ids[i] == target;
  c = warp_sum_i(c);
  if (lane == 0) warp_cnt[warp] = c;
  __syncthreads();
  int before = s_base;
  for (int w = 0; w < warp; ++w) before += warp_cnt[w];
  for (int i0 = lo; i0 < hi; i0 += 32) {
    const int i = i0 + lane;
    const bool m = i < hi && ids[i] == target;
    const unsigned b = __ballot_sync(0xffffffffu, m);
    if (m) {
      const int pos = before + __popc(b & ((1u << lane) - 1u));
      sorted_pair[pos] = i;
      pos_of_pair[i] = pos;
    }
    before += __popc(b);
  }
  if (threadIdx.x == 0) {
    offsets[e] = s_base;
    if (e == num_local - 1) {
      int tot = s_base;
      for (int w = 0; w < 8; ++w) tot += warp_cnt[w];
      offsets[num_local] = tot;
    }
  }
}

// ------------------------------------------------------------------------------------ expert GEMMs
struct MoeGemmParams {
  const void* a;               // gate_up: X [tokens, K] ; down: H [pairs, K]
  const void* w;               // gate_up: [num_local, 2*n_out, K] ; down: [num_local, n_out, K]
  void* out;                   // gate_up: H [pairs, n_out] ; down: Y [pairs, n_out]
  const int32_t* offsets;      // [num_local + 1] rows of expert e: [offsets[e], offsets[e+1])
  const int32_t* sorted_pair;  // [pairs] pair index t*top_k + j at each sorted position
  const void* pair_scale;      // down: routing weight of every pair, [tokens * top_k]
  int64_t a_stride;            // elements between rows of a
  int K, n_out, top_k;
};

constexpr int kMoeStages = 4;
constexpr int kMoeBK = 64;                      // elements: one 128-byte row per tile row
constexpr int kMoeBRows = 128;                  // weight rows per CTA tile
constexpr int kMoeARows = 64;                   // token rows per pass (2 x 32)
constexpr int kMoeStageBytes = (kMoeARows + kMoeBRows) * 128;
constexpr int kMoeSmemBytes = kMoeStages * kMoeStageBytes + 1024;

// swizzled byte offset of 16-byte chunk c16 (0..7) of a 128-byte tile row
__device__ __forceinline__ int moe_tile_off(int row, int c16) {
  return row * 128 + ((c16 ^ (row & 7)) << 4);
}

template <typename T>
__device__ __forceinline__ float silu_round(float g) {
  // F.silu evaluated in fp32, rounded to the model dtype (same formula as silu_and_mul_kernel)
  return to_f32<T>(from_f32<T>(__fdividef(g, 1.0f + __expf(-g))));
}

// One pass: rows [m0, m0 + 32*MI) of expert e against the CTA's 128 weight rows, full K.
template <typename T, bool kGateUp, int MI>
__device__ __forceinline__ void moe_gemm_pass(const MoeGemmParams& p, uint8_t* smem, const T* wbase,
                                              int row_begin, int cnt, int m0, int n0) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c16 = tid & 7, rowbase = tid >> 3;
  const int KB = p.K / kMoeBK;
  const T* ag = reinterpret_cast<const T*>(p.a);

  const T* a_src[MI];
  bool a_ok[MI];
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int r = m0 + rowbase + 32 * i;
    a_ok[i] = r < cnt;
    const int pos = row_begin + (a_ok[i] ? r : 0);
    const int64_t src_row = kGateUp ? (int64_t)(p.sorted_pair[pos] / p.top_k) : (int64_t)pos;
    a_src[i] = ag + src_row * p.a_stride + c16 * 8;
  }
  const T* b_src[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = rowbase + 32 * i;
    const int wrow = kGateUp ? (r < 64 ? n0 + r : p.n_out + n0 + (r - 64)) : n0 + r;
    b_src[i] = wbase + (int64_t)wrow * p.K + c16 * 8;
  }
  auto load_stage = [&](int stage, int kb) {
    uint8_t* sa = smem + stage * kMoeStageBytes;
    uint8_t* sb = sa + kMoeARows * 128;
    const int koff = kb * kMoeBK;
#pragma unroll
    for (int i = 0; i < MI; ++i)
      cp_async_16(smem_u32(sa + moe_tile_off(rowbase + 32 * i, c16)), a_src[i] + koff, a_ok[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      cp_async_16(smem_u32(sb + moe_tile_off(rowbase + 32 * i, c16)), b_src[i] + koff, true);
  };

  float acc[2 * MI][2][4];
#pragma unroll
  for (int mf = 0; mf < 2 * MI; ++mf)
#pragma unroll
    for (int nf = 0; nf < 2; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = acc[mf][nf][2] = acc[mf][nf][3] = 0.f;

#pragma unroll
  for (int s = 0; s < kMoeStages - 1; ++s) {
    if (s < KB) load_stage(s, s);
    cp_async_commit();
  }
  // per-lane ldmatrix coordinates (same conventions as prefill_attn.cu)
  const int a_row = (((lane >> 3) & 1) << 3) + (lane & 7), a_half = lane >> 4;
  const int b_row = kGateUp ? ((lane >> 4) ? 64 : 0) + warp * 8 + (lane & 7)
                            : warp * 16 + ((lane >> 4) << 3) + (lane & 7);
  const int b_half = (lane >> 3) & 1;

  for (int kb = 0; kb < KB; ++kb) {
    cp_async_wait<kMoeStages - 2>();
    __syncthreads();
    const int nk = kb + kMoeStages - 1;
    if (nk < KB) load_stage(nk % kMoeStages, nk);
    cp_async_commit();
    const uint32_t sa = smem_u32(smem + (kb % kMoeStages) * kMoeStageBytes);
    const uint32_t sb = sa + kMoeARows * 128;
#pragma unroll
    for (int ks = 0; ks < kMoeBK / 16; ++ks) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(sb + moe_tile_off(b_row, ks * 2 + b_half), b0, b1, b2, b3);
#pragma unroll
      for (int mf = 0; mf < 2 * MI; ++mf) {
        uint32_t a0, a1, a2, a3;
        ldsm_x4(sa + moe_tile_off(mf * 16 + a_row, ks * 2 + a_half), a0, a1, a2, a3);
        mma_16816<T>(acc[mf][0], a0, a1, a2, a3, b0, b1);
        mma_16816<T>(acc[mf][1], a0, a1, a2, a3, b2, b3);
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();               // the ring is free for the next pass

  // ---- epilogue (C fragment: rows g / g+8, columns 2*t4, 2*t4+1 of each 8-wide n fragment)
  const int g = lane >> 2, t4 = lane & 3;
  T* og = reinterpret_cast<T*>(p.out);
#pragma unroll
  for (int mf = 0; mf < 2 * MI; ++mf) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = m0 + mf * 16 + g + 8 * r;
      if (row >= cnt) continue;
      const int64_t pos = row_begin + row;
      if (kGateUp) {
        // gate_up in the model dtype (F.linear), silu in fp32 -> dtype, dtype multiply with `up`
        const float g0 = to_f32<T>(from_f32<T>(acc[mf][0][2 * r])), g1 = to_f32<T>(from_f32<T>(acc[mf][0][2 * r + 1]));
        const float u0 = to_f32<T>(from_f32<T>(acc[mf][1][2 * r])), u1 = to_f32<T>(from_f32<T>(acc[mf][1][2 * r + 1]));
        const uint32_t h = pack2<T>(__fmul_rn(silu_round<T>(g0), u0), __fmul_rn(silu_round<T>(g1), u1));
        *reinterpret_cast<uint32_t*>(og + pos * p.n_out + n0 + warp * 8 + 2 * t4) = h;
      } else {
        // expert output in the model dtype, then a dtype multiply with the routing weight
        const float sc = to_f32<T>(reinterpret_cast<const T*>(p.pair_scale)[p.sorted_pair[pos]]);
#pragma unroll
        for (int nf = 0; nf < 2; ++nf) {
          const float y0 = to_f32<T>(from_f32<T>(acc[mf][nf][2 * r])), y1 = to_f32<T>(from_f32<T>(acc[mf][nf][2 * r + 1]));
          *reinterpret_cast<uint32_t*>(og + pos * p.n_out + n0 + warp * 16 + nf * 8 + 2 * t4) =
              pack2<T>(__fmul_rn(y0, sc), __fmul_rn(y1, sc));
        }
      }
    }
  }
}

// grid (n tiles, num_local experts); gate_up: tile = 64 gate + the 64 matching up rows of W13,
// down: 128 rows of W2.
template <typename T, bool kGateUp>
__global__ void __launch_bounds__(256, 2)
moe_gemm_kernel(const MoeGemmParams p) {
  extern __shared__ __align__(1024) uint8_t moe_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(moe_smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  pdl_wait();
  const int e = blockIdx.y;
  const int row_begin = p.offsets[e];
  const int cnt = p.offsets[e + 1] - row_begin;
  if (cnt <= 0) return;
  const int n0 = blockIdx.x * (kGateUp ? 64 : 128);
  const T* wbase = reinterpret_cast<const T*>(p.w) + (int64_t)e * (kGateUp ? 2 : 1) * p.n_out * p.K;
  for (int m0 = 0; m0 < cnt;) {
    if (cnt - m0 > 32) {
      moe_gemm_pass<T, kGateUp, 2>(p, smem, wbase, row_begin, cnt, m0, n0);
      m0 += 64;
    } else {
      moe_gemm_pass<T, kGateUp, 1>(p, smem, wbase, row_begin, cnt, m0, n0);
      m0 += 32;
    }
  }
}

// ------------------------------------------------------------------------------------ combine
// One CTA per token: the token's local pairs in ascending expert order, out = ((y1 + y2) + y3)...
// with a rounding to the model dtype after every add; zeros if no local expert was selected.
template <typename T>
__global__ void __launch_bounds__(256)
moe_combine_kernel(T* __restrict__ out, const T* __restrict__ y, const int32_t* __restrict__ topk_ids,
                   const int32_t* __restrict__ pos_of_pair, int k, int hidden, int64_t out_stride) {
  __shared__ int s_pos[kMoeMaxTopK];
  __shared__ int s_n;
  pdl_wait();
  const int64_t t = blockIdx.x;
  if (threadIdx.x == 0) {
    int ids[kMoeMaxTopK], pos[kMoeMaxTopK], n = 0;
    for (int j = 0; j < k; ++j) {
      const int ps = pos_of_pair[t * k + j];
      if (ps < 0) continue;
      const int id = topk_ids[t * k + j];
      int at = n++;
      while (at > 0 && ids[at - 1] > id) {
        ids[at] = ids[at - 1];
        pos[at] = pos[at - 1];
        --at;
      }
      ids[at] = id;
      pos[at] = ps;
    }
    for (int j = 0; j < n; ++j) s_pos[j] = pos[j];
    s_n = n;
  }
  __syncthreads();
  const int n = s_n, nvec = hidden / 8;
  uint4* orow = reinterpret_cast<uint4*>(out + t * out_stride);
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int j = 0; j < n; ++j) {
      const uint4 q = ld_nc_v4(reinterpret_cast<const uint4*>(y + (int64_t)s_pos[j] * hidden) + v);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack2<T>(w[i]);
        acc[2 * i] = to_f32<T>(from_f32<T>(__fadd_rn(acc[2 * i], f.x)));
        acc[2 * i + 1] = to_f32<T>(from_f32<T>(__fadd_rn(acc[2 * i + 1], f.y)));
      }
    }
    orow[v] = make_uint4(pack2<T>(acc[0], acc[1]), pack2<T>(acc[2], acc[3]), pack2<T>(acc[4], acc[5]),
                         pack2<T>(acc[6], acc[7]));
  }
}

template <typename T>
static int moe_gemm_launch(bool gate_up, const MoeGemmParams& p, int num_local, cudaStream_t stream) {
  auto kern = gate_up ? moe_gemm_kernel<T, true> : moe_gemm_kernel<T, false>;
  static bool configured[2] = {false, false};
  if (!configured[gate_up]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMoeSmemBytes);
    if (e != cudaSuccess) return (int)e;
    configured[gate_up] = true;
  }
  const dim3 grid(p.n_out / (gate_up ? 64 : 128), num_local);
  launch_pdl(kern, grid, dim3(256), (size_t)kMoeSmemBytes, stream, p);
  return finish_launch();
}

}  // namespace b200

using namespace b200;

extern "C" int b200_moe_route(int32_t* topk_ids, void* topk_weights, const void* router_logits,
                              int32_t tokens, int32_t num_experts, int32_t top_k,
                              int64_t logits_stride, int32_t dtype, b200_stream_t stream) {
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(tokens < 0 || num_experts <= 0 || top_k <= 0 || top_k > num_experts, B200_ERR_BAD_SHAPE);
  B200_RETURN_IF(top_k > kMoeMaxTopK || num_experts > kMoeMaxExperts, B200_ERR_UNSUPPORTED);
  if (tokens == 0) return B200_OK;
  B200_RETURN_IF(!topk_ids || !topk_weights || !router_logits, B200_ERR_NULL_POINTER);
  const dim3 grid((tokens + 7) / 8), block(256);
  const int vpl = (num_experts + 31) / 32;
#define B200_ROUTE(T, V)                                                                       \
  launch_pdl(moe_route_kernel<T, V>, grid, block, 0, (cudaStream_t)stream, topk_ids,           \
             (T*)topk_weights, (const T*)router_logits, tokens, num_experts, top_k, logits_stride)
#define B200_ROUTE_T(T)            \
  if (vpl <= 1) B200_ROUTE(T, 1);  \
  else if (vpl <= 2) B200_ROUTE(T, 2); \
  else if (vpl <= 4) B200_ROUTE(T, 4); \
  else B200_ROUTE(T, 8)
  if (dtype == B200_DTYPE_BF16) {
    B200_ROUTE_T(__nv_bfloat16);
  } else {
    B200_ROUTE_T(__half);
  }
#undef B200_ROUTE_T
#undef B200_ROUTE
  return finish_launch();
}

extern "C" int b200_moe_align(int32_t* offsets, int32_t* sorted_pair, int32_t* pos_of_pair,
                              int32_t* counts, const int32_t* topk_ids, int32_t num_pairs,
                              int32_t expert_begin, int32_t num_local, b200_stream_t stream) {
  B200_RETURN_IF(num_pairs < 0 || expert_begin < 0 || num_local <= 0, B200_ERR_BAD_SHAPE);
  B200_RETURN_IF(num_local > kMoeMaxExperts, B200_ERR_UNSUPPORTED);
  B200_RETURN_IF(!offsets || !sorted_pair || !pos_of_pair || !counts || !topk_ids, B200_ERR_NULL_POINTER);
  launch_pdl(moe_count_kernel, dim3(num_local), dim3(256), 0, (cudaStream_t)stream, counts, pos_of_pair,
             topk_ids, num_pairs, expert_begin);
  launch_pdl(moe_place_kernel, dim3(num_local), dim3(256), 0, (cudaStream_t)stream, offsets, sorted_pair,
             pos_of_pair, (const int32_t*)counts, topk_ids, num_pairs, expert_begin, num_local);
  return finish_launch(2);
}

static int moe_gemm_check(const void* a, const void* w, void* out, const int32_t* offsets,
                          const int32_t* sorted_pair, int32_t K, int32_t n_out, int32_t n_tile,
                          int32_t num_local, int32_t top_k, int64_t a_stride, int32_t dtype) {
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(K <= 0 || n_out <= 0 || num_local <= 0 || top_k <= 0 || a_stride < K, B200_ERR_BAD_SHAPE);
  B200_RETURN_IF(K % kMoeBK != 0 || n_out % n_tile != 0 || num_local > 65535, B200_ERR_UNSUPPORTED);
  B200_RETURN_IF(!a || !w || !out || !offsets || !sorted_pair, B200_ERR_NULL_POINTER);
  B200_RETURN_IF(!aligned16(a) || !aligned16(w) || !aligned16(out) || a_stride % 8 != 0, B200_ERR_MISALIGNED);
  return B200_OK;
}

extern "C" int b200_moe_gemm_gate_up(void* h, const void* x, const void* w13, const int32_t* offsets,
                                     const int32_t* sorted_pair, int32_t hidden, int32_t inter,
                                     int32_t num_local, int32_t top_k, int64_t x_stride, int32_t dtype,
                                     b200_stream_t stream) {
  const int rc = moe_gemm_check(x, w13, h, offsets, sorted_pair, hidden, inter, 64, num_local, top_k,
                                x_stride, dtype);
  if (rc != B200_OK) return rc;
  MoeGemmParams p{x, w13, h, offsets, sorted_pair, nullptr, x_stride, hidden, inter, top_k};
  return dtype == B200_DTYPE_BF16 ? moe_gemm_launch<__nv_bfloat16>(true, p, num_local, (cudaStream_t)stream)
                                  : moe_gemm_launch<__half>(true, p, num_local, (cudaStream_t)stream);
}

extern "C" int b200_moe_gemm_down(void* y, const void* h, const void* w2, const int32_t* offsets,
                                  const int32_t* sorted_pair, const void* topk_weights, int32_t inter,
                                  int32_t hidden, int32_t num_local, int32_t top_k, int32_t dtype,
                                  b200_stream_t stream) {
  const int rc = moe_gemm_check(h, w2, y, offsets, sorted_pair, inter, hidden, 128, num_local, top_k,
                                inter, dtype);
  if (rc != B200_OK) return rc;
  B200_RETURN_IF(!topk_weights, B200_ERR_NULL_POINTER);
  MoeGemmParams p{h, w2, y, offsets, sorted_pair, topk_weights, inter, inter, hidden, top_k};
  return dtype == B200_DTYPE_BF16 ? moe_gemm_launch<__nv_bfloat16>(false, p, num_local, (cudaStream_t)stream)
                                  : moe_gemm_launch<__half>(false, p, num_local, (cudaStream_t)stream);
}

extern "C" int b200_moe_combine(void* out, const void* y, const int32_t* topk_ids,
                                const int32_t* pos_of_pair, int32_t tokens, int32_t top_k,
                                int32_t hidden, int64_t out_stride, int32_t dtype,
                                b200_stream_t stream) {
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(tokens < 0 || top_k <= 0 || hidden <= 0 || out_stride < hidden, B200_ERR_BAD_SHAPE);
  B200_RETURN_IF(top_k > kMoeMaxTopK, B200_ERR_UNSUPPORTED);
  if (tokens == 0) return B200_OK;
  B200_RETURN_IF(!out || !y || !topk_ids || !pos_of_pair, B200_ERR_NULL_POINTER);
  B200_RETURN_IF(hidden % 8 != 0 || out_stride % 8 != 0 || !aligned16(out) || !aligned16(y),
                 B200_ERR_MISALIGNED);
  if (dtype == B200_DTYPE_BF16)
    launch_pdl(moe_combine_kernel<__nv_bfloat16>, dim3(tokens), dim3(256), 0, (cudaStream_t)stream,
               (__nv_bfloat16*)out, (const __nv_bfloat16*)y, topk_ids, pos_of_pair, top_k, hidden,
               out_stride);
  else
    launch_pdl(moe_combine_kernel<__half>, dim3(tokens), dim3(256), 0, (cudaStream_t)stream, (__half*)out,
               (const __half*)y, topk_ids, pos_of_pair, top_k, hidden, out_stride);
  return finish_launch();
}
