// decode_common.cuh -- pieces shared by the paged decode attention kernels (bf16/fp16 cache:
// decode_attn.cu; fp8 cache: decode_attn_fp8.cu): launch constants, the parameter block, the
// stream-K work-list search, the split-KV combine kernel and small host helpers.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kConsumerWarps = 8;
constexpr int kDecodeThreads = (kConsumerWarps + 1) * 32;
constexpr int kMaxItems = 2048;   // batch * head_groups per launch
constexpr int kMaxCtas = 304;     // partial-result slots are sized for this many CTAs

struct DecodeParams {
  const void* q;
  void* out;
  const int32_t* block_tables;
  const int32_t* context_lens;
  float* part_o;    // [kMaxCtas*2][rows][D]
  float* part_ml;   // [kMaxCtas*2][rows][2]
  int32_t* header;  // [items + 3]: prefix[items+1], per, total
  int32_t* owner;   // [kMaxCtas]: item that CTA c must have merged at boundary c|c+1, or -1
  int64_t q_stride, out_stride;
  int batch, H, Hkv, G, block_size, bt_stride;
  int hg_count;  // head groups per sequence = Hkv / HKV_CTA
  int rows_box;  // tokens per TMA op = min(block_size, TILE_TOKENS)
  float scale_log2;
};

// first idx in [1, items] with prefix[idx] > t, minus one
__device__ __forceinline__ int find_item(const int* prefix, int items, int t) {
  int lo = 0, hi = items;  // answer in [lo, hi)
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (prefix[mid] <= t) lo = mid; else hi = mid;
  }
  return lo;
}

// Combine kernel: merges the partial results of items whose tiles were split across CTAs.
// One CTA per range boundary c | c+1: the item that straddles that boundary is merged by the
// FIRST boundary it straddles -- the decode CTA on the left of that boundary records the item in
// owner[c] -- (every other block exits after one load), so the grid is #CTAs-1 blocks
// instead of one block per (item, row). gridDim.y == 1 (large batches, few pieces per item):
// the 8 warps take the rows round-robin and fold the pieces serially. gridDim.y == rows (small
// batches, one item cut into up to #CTAs pieces): block (c, r) folds row r, its 8 warps split
// the pieces and meet in shared memory. Lane = D/32 consecutive columns in both modes.
template <typename T, int D>
__global__ void __launch_bounds__(256)
decode_combine_kernel(const DecodeParams p, int rows /* = HKV_CTA*G */) {
  constexpr int V = D / 32;
  __shared__ float s_acc[8][D];
  __shared__ float s_ml[8][2];
  pdl_wait();
  const int items = p.batch * p.hg_count;
  const int per = p.header[items + 1], total = p.header[items + 2];
  const int c = blockIdx.x;
  const int it = p.owner[c];                    // item cut at this boundary and first cut here
  if (it < 0) return;
  (void)total;
  const int t0 = p.header[it], t1 = p.header[it + 1];
  const int c_first = c, c_last = (t1 - 1) / per;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = it / p.hg_count, hg = it - b * p.hg_count;
  T* obase = reinterpret_cast<T*>(p.out) + (int64_t)b * p.out_stride + (int64_t)hg * rows * D;
  const bool by_rows = gridDim.y == 1;
  const int r_begin = by_rows ? warp : (int)blockIdx.y;
  const int r_step = by_rows ? 8 : rows;        // second mode: exactly one row per block
  for (int r = r_begin; r < rows; r += r_step) {
    float m = -INFINITY, l = 0.f, acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    for (int cc = c_first + (by_rows ? 0 : warp); cc <= c_last; cc += (by_rows ? 1 : 8)) {
      const int slot = 2 * cc + ((cc * per >= t0) ? 0 : 1);
      const float* ml = p.part_ml + ((int64_t)slot * rows + r) * 2;
      const float* po = p.part_o + ((int64_t)slot * rows + r) * D + lane * V;
      const float mc = ml[0], lc = ml[1];
      const float m_new = fmaxf(m, mc);
      const float a = ex2_approx(m - m_new), bsc = ex2_approx(mc - m_new);
      l = l * a + lc * bsc;
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = acc[i] * a + po[i] * bsc;
      m = m_new;
    }
    if (!by_rows) {                             // fold the 8 warps' shares of this row
#pragma unroll
      for (int i = 0; i < V; ++i) s_acc[warp][lane * V + i] = acc[i];
      if (lane == 0) s_ml[warp][0] = m, s_ml[warp][1] = l;
      __syncthreads();
      if (warp != 0) return;
      m = -INFINITY;
#pragma unroll
      for (int w = 0; w < 8; ++w) m = fmaxf(m, s_ml[w][0]);
      l = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const float sc = ex2_approx(s_ml[w][0] - m);   // exp2(-inf) == 0 for idle warps
        l = fmaf(s_ml[w][1], sc, l);
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = fmaf(s_acc[w][lane * V + i], sc, acc[i]);
      }
    }
    const float inv = 1.f / l;
    T* orow = obase + (int64_t)r * D + lane * V;
#pragma unroll
    for (int i = 0; i < V; i += 2)
      *reinterpret_cast<uint32_t*>(orow + i) = pack2<T>(acc[i] * inv, acc[i + 1] * inv);
  }
}

static inline int hkv_cta_of(int Hkv) {
  if (Hkv >= 8) return Hkv % 8 == 0 ? 8 : 0;
  return (Hkv == 1 || Hkv == 2 || Hkv == 4) ? Hkv : 0;
}

static inline int num_sms() {
  static int n = [] {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess)
      cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
  }();
  return n;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace b200
