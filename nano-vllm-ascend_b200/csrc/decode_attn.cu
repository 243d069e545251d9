// decode_attn.cu -- K6: paged decode attention, the HBM-roofline kernel of the path.
//
// Replaces npu_fused_infer_attention_score(_v2) BNSD+block_table (layers/attention.py:60-93),
// flash_attn_with_kvcache (layers/attention_ori.py:72-74) and the oracle
// Attention._decode_attn (layers/attention_torch_native.py:147-196) of the reference.
//
// Design (B200-first):
//  * The cache layout [num_blocks, block_size, Hkv*D] makes the K (or V) rows of one block
//    for ALL kv heads one contiguous run, so a CTA consumes whole blocks: a TMA tensor map
//    over the cache viewed as {64 elems, rows = nb*bs, 128-byte chunks = Hkv*D/64} (strides
//    {row_bytes, 128}) lands each block in shared memory as [chunk][row][128 B] panels with
//    the hardware 128-byte swizzle => bank-conflict-free ldmatrix, no register staging.
//  * Persistent grid (one CTA per SM). The global list of (sequence, head-group, 16/32/../128
//    token tile) work is cut into equal contiguous ranges ("stream-K" over KV tiles), so all
//    SMs stream the same number of bytes regardless of batch size or ragged lengths; only
//    range boundaries produce partial (m, l, O) results, merged by a tiny combine kernel.
//  * Warp specialisation: warp 8 is the TMA producer (block-table lookups for 32 tiles at a
//    time, one lane each), warps 0-7 are consumers; a 3-stage mbarrier ring, separate
//    K-full / V-full barriers so QK^T starts as soon as K lands.
//  * Per consumer warp: one kv head x one 16-token unit per tile. The GQA group (<= 8 query
//    heads) is the M dimension of mma.sync m16n8k16 (rows 8-15 are structurally zero), i.e.
//    the whole group shares a single read of the KV bytes. fp32 online softmax in the exp2
//    domain; running max/sum per thread, reduced with warp shuffles.
//  Algorithmic bytes per launch: sum_b ctx_b * 2 * Hkv * D * 2 B (+ q, out: negligible).
#include "decode_common.cuh"

#include <mutex>
#include <unordered_map>

namespace b200 {

constexpr int kStages = 3;

template <typename T, int D, int HKV_CTA>
__global__ void __launch_bounds__(kDecodeThreads, 1)
paged_decode_kernel(const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const DecodeParams p) {
  constexpr int UNITS = 8 / HKV_CTA;          // 16-token units per tile (per head)
  constexpr int TILE_TOKENS = 16 * UNITS;
  constexpr int CH = D / 64;                  // 128-byte chunks per head row
  constexpr int TILE_BYTES = 16 * 8 * D * 2;  // bytes of K (or V) per tile, all variants
  constexpr int KSTEPS = D / 16;
  constexpr int NT = D / 8;                   // 8-wide output column tiles
  constexpr int MERGE_FLOATS = 8 * D + 16;    // one warp's (O rows, m, l)

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* kbuf = smem;                                  // [kStages][TILE_BYTES]
  uint8_t* vbuf = smem + kStages * TILE_BYTES;           // [kStages][TILE_BYTES]
  float* merge = reinterpret_cast<float*>(smem + 2 * kStages * TILE_BYTES);  // [4][MERGE_FLOATS]
  int* prefix = reinterpret_cast<int*>(merge + (UNITS > 1 ? 4 * MERGE_FLOATS : 0));
  const int items = p.batch * p.hg_count;
  uint64_t* bars = reinterpret_cast<uint64_t*>(
      (reinterpret_cast<uintptr_t>(prefix + items + 1) + 7) & ~uintptr_t(7));
  uint64_t* full_k = bars;
  uint64_t* full_v = bars + kStages;
  uint64_t* empty = bars + 2 * kStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_wait();

  // ---- tiles per item -> exclusive prefix sum in shared memory
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int len = p.context_lens[i / p.hg_count];
    prefix[i + 1] = len > 0 ? (len + TILE_TOKENS - 1) / TILE_TOKENS : 0;
  }
  if (threadIdx.x == 0) {
    prefix[0] = 0;
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_k[s], 1);
      mbar_init(&full_v[s], 1);
      mbar_init(&empty[s], kConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (warp == 0) {
    int carry = 0;
    for (int base = 0; base < items; base += 32) {
      const int i = base + lane;
      int v = i < items ? prefix[i + 1] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      if (i < items) prefix[i + 1] = carry + v;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
  }
  __syncthreads();
  const int total = prefix[items];
  const int per = total > 0 ? (total + gridDim.x - 1) / gridDim.x : 1;
  const int lo = min((int)blockIdx.x * per, total);
  const int hi = min(lo + per, total);
  if (blockIdx.x == 0) {  // publish the partition for the combine kernel
    for (int i = threadIdx.x; i <= items; i += blockDim.x) p.header[i] = prefix[i];
    if (threadIdx.x == 0) p.header[items + 1] = per, p.header[items + 2] = total;
  }

  const int OPS = TILE_TOKENS / p.rows_box;  // TMA ops per full tile (K or V)
  const int op_bytes = p.rows_box * HKV_CTA * D * 2;

  if (warp == kConsumerWarps) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
    }
    const uint64_t policy = l2_policy_evict_first();
    int stage = 0, phase = 0;
    for (int t0 = lo; t0 < hi; t0 += 32) {
      const int tt = t0 + lane;
      int nops = 0, chunk0 = 0;
      int rows[UNITS];
#pragma unroll
      for (int o = 0; o < UNITS; ++o) rows[o] = -1;
      if (tt < hi) {
        const int it = find_item(prefix, items, tt);
        const int j = tt - prefix[it];
        const int b = it / p.hg_count, hg = it - b * p.hg_count;
        const int len = p.context_lens[b];
        const int tok0 = j * TILE_TOKENS;
        nops = min(OPS, (len - tok0 + p.rows_box - 1) / p.rows_box);
        chunk0 = hg * HKV_CTA * CH;
#pragma unroll
        for (int o = 0; o < UNITS; ++o) {
          if (o < nops) {
            const int tok = tok0 + o * p.rows_box;
            const int entry = tok / p.block_size;
            const int blk = p.block_tables[(int64_t)b * p.bt_stride + entry];
            rows[o] = blk < 0 ? -(1 << 30) : blk * p.block_size + (tok - entry * p.block_size);
          }
        }
      }
      const int n = min(32, hi - t0);
      for (int i = 0; i < n; ++i) {
        const int my_nops = __shfl_sync(0xffffffffu, nops, i);
        const int my_chunk0 = __shfl_sync(0xffffffffu, chunk0, i);
        int my_rows[UNITS];
#pragma unroll
        for (int o = 0; o < UNITS; ++o) my_rows[o] = __shfl_sync(0xffffffffu, rows[o], i);
        if (lane == 0) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full_k[stage], my_nops * op_bytes);
#pragma unroll
          for (int o = 0; o < UNITS; ++o)
            if (o < my_nops)
              tma_load_3d(kbuf + stage * TILE_BYTES + o * op_bytes, &tmK, &full_k[stage], 0,
                          my_rows[o], my_chunk0, policy);
          mbar_expect_tx(&full_v[stage], my_nops * op_bytes);
#pragma unroll
          for (int o = 0; o < UNITS; ++o)
            if (o < my_nops)
              tma_load_3d(vbuf + stage * TILE_BYTES + o * op_bytes, &tmV, &full_v[stage], 0,
                          my_rows[o], my_chunk0, policy);
        }
        if (++stage == kStages) stage = 0, phase ^= 1;
      }
    }
    return;
  }

  // =========================== consumers ===========================
  const int head_local = warp % HKV_CTA;
  const int unit = warp / HKV_CTA;
  const int g = lane >> 2, t4 = lane & 3;
  const int G = p.G;
  const float sl2 = p.scale_log2;
  const T* qbase = reinterpret_cast<const T*>(p.q);
  T* obase = reinterpret_cast<T*>(p.out);

  // rows of zero-length sequences (padded graph rows): emit zeros, touch nothing else
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    if (prefix[it + 1] == prefix[it]) {
      const int b = it / p.hg_count, hg = it - b * p.hg_count;
      const int n = HKV_CTA * G * D;  // contiguous elements of this head group
      T* dst = obase + (int64_t)b * p.out_stride + (int64_t)hg * n;
      for (int e = threadIdx.x * 2; e < n; e += kConsumerWarps * 64)
        *reinterpret_cast<uint32_t*>(dst + e) = 0u;
    }
  }

  // fixed per-lane smem offsets inside a unit (swizzle-128B panels of [row][128 B])
  const int tok_off = unit * 16;
  const int op_idx = tok_off / p.rows_box;
  const int r0 = tok_off - op_idx * p.rows_box;
  const int panel_bytes = p.rows_box * 128;
  const int unit_base = op_idx * op_bytes + (head_local * CH) * panel_bytes + r0 * 128;
  // K (non-transposed ldmatrix): matrices (tok0-7,lo16B) (tok0-7,hi16B) (tok8-15,lo) (tok8-15,hi)
  const int k_tok = ((lane >> 4) << 3) + (lane & 7);
  const int k_half = (lane >> 3) & 1;
  // V (transposed ldmatrix): matrices (tok0-7,dt) (tok8-15,dt) (tok0-7,dt+1) (tok8-15,dt+1)
  const int v_tok = (((lane >> 3) & 1) << 3) + (lane & 7);
  const int v_dt = lane >> 4;

  int stage = 0, phase = 0;
  int seg_index = 0;
  int t = lo;
  int own_item = -1;   // split item whose first piece is this CTA's last segment
  int it = total > 0 && lo < hi ? find_item(prefix, items, lo) : items;
  while (t < hi) {
    const int item_t0 = prefix[it], ntiles = prefix[it + 1] - item_t0;
    const int j0 = t - item_t0;
    const int j1 = min(ntiles, j0 + (hi - t));
    const int b = it / p.hg_count, hg = it - b * p.hg_count;
    const int len = p.context_lens[b];
    const int kv_head = hg * HKV_CTA + head_local;

    // ---- Q fragments for this (sequence, kv head): rows = query heads of the GQA group
    uint32_t qa[KSTEPS][2];
    {
      const T* qrow = qbase + (int64_t)b * p.q_stride + (int64_t)(kv_head * G + g) * D;
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        qa[ks][0] = g < G ? *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 2 * t4) : 0u;
        qa[ks][1] = g < G ? *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 8 + 2 * t4) : 0u;
      }
    }
    float o_acc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) o_acc[n][0] = o_acc[n][1] = o_acc[n][2] = o_acc[n][3] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;

    for (int j = j0; j < j1; ++j) {
      const int tok0 = j * TILE_TOKENS + tok_off;  // first token of this warp's unit
      const bool active = tok0 < len;
      mbar_wait(&full_k[stage], phase);
      if (active) {
        const uint32_t kb = smem_u32(kbuf + stage * TILE_BYTES) + unit_base;
        float s[2][4];
#pragma unroll
        for (int n = 0; n < 2; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
          const int c = ks >> 2;
          const int c16 = ((ks & 3) << 1) + k_half;
          uint32_t r0_, r1_, r2_, r3_;
          ldsm_x4(kb + c * panel_bytes + k_tok * 128 + ((c16 ^ (k_tok & 7)) << 4), r0_, r1_, r2_,
                  r3_);
          mma_16816<T>(s[0], qa[ks][0], 0u, qa[ks][1], 0u, r0_, r1_);
          mma_16816<T>(s[1], qa[ks][0], 0u, qa[ks][1], 0u, r2_, r3_);
        }
        // mask tokens beyond the context (only the last tile of a sequence)
        const int valid = len - tok0;  // > 0
        if (valid < 16) {
#pragma unroll
          for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int e = 0; e < 2; ++e)
              if (n * 8 + 2 * t4 + e >= valid) s[n][e] = -INFINITY;
        }
        // online softmax for row g (rows 8..15 are padding)
        float mx = fmaxf(fmaxf(s[0][0], s[0][1]), fmaxf(s[1][0], s[1][1]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float m_new = fmaxf(m_run, mx);
        const float alpha = ex2_approx((m_run - m_new) * sl2);
        const float mb = m_new * sl2;
        const float p00 = ex2_approx(fmaf(s[0][0], sl2, -mb));
        const float p01 = ex2_approx(fmaf(s[0][1], sl2, -mb));
        const float p10 = ex2_approx(fmaf(s[1][0], sl2, -mb));
        const float p11 = ex2_approx(fmaf(s[1][1], sl2, -mb));
        l_run = fmaf(l_run, alpha, (p00 + p01) + (p10 + p11));
        m_run = m_new;
        if (alpha != 1.f) {
#pragma unroll
          for (int n = 0; n < NT; ++n) o_acc[n][0] *= alpha, o_acc[n][1] *= alpha;
        }
        const uint32_t pa0 = pack2<T>(p00, p01), pa2 = pack2<T>(p10, p11);

        mbar_wait(&full_v[stage], phase);
        uint8_t* vstage = vbuf + stage * TILE_BYTES + unit_base;
        if (valid < 16) {
          // rows past the context hold stale cache bytes (possibly NaN patterns): zero them so
          // that 0 * garbage cannot poison the accumulator.
          const int nz = (16 - valid) * CH * 8;
          for (int idx = lane; idx < nz; idx += 32) {
            const int row = valid + idx / (CH * 8);
            const int rem = idx % (CH * 8);
            *reinterpret_cast<uint4*>(vstage + (rem >> 3) * panel_bytes + row * 128 +
                                      ((rem & 7) << 4)) = make_uint4(0, 0, 0, 0);
          }
          __syncwarp();
        }
        const uint32_t vb = smem_u32(vstage);
#pragma unroll
        for (int jp = 0; jp < NT / 2; ++jp) {
          const int dt = 2 * jp + v_dt;
          uint32_t r0_, r1_, r2_, r3_;
          ldsm_x4_trans(vb + (dt >> 3) * panel_bytes + v_tok * 128 + (((dt & 7) ^ (v_tok & 7)) << 4),
                        r0_, r1_, r2_, r3_);
          mma_16816<T>(o_acc[2 * jp], pa0, 0u, pa2, 0u, r0_, r1_);
          mma_16816<T>(o_acc[2 * jp + 1], pa0, 0u, pa2, 0u, r2_, r3_);
        }
        if (valid < 16) fence_proxy_async_smem();
      } else {
        mbar_wait(&full_v[stage], phase);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == kStages) stage = 0, phase ^= 1;
    }

    // ---- finalise this segment
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
    if (UNITS > 1) {
      // tree-merge the warps that share a kv head (different 16-token units of each tile)
#pragma unroll
      for (int stride = UNITS / 2; stride >= 1; stride >>= 1) {
        float* slot = merge + ((unit % stride) * HKV_CTA + head_local) * MERGE_FLOATS;
        if (unit >= stride && unit < 2 * stride) {
#pragma unroll
          for (int n = 0; n < NT; ++n)
            *reinterpret_cast<float2*>(slot + g * D + n * 8 + 2 * t4) =
                make_float2(o_acc[n][0], o_acc[n][1]);
          if (t4 == 0) slot[8 * D + g] = m_run, slot[8 * D + 8 + g] = l_run;
        }
        named_bar_sync(1, kConsumerWarps * 32);
        if (unit < stride) {
          const float m_o = slot[8 * D + g], l_o = slot[8 * D + 8 + g];
          const float m_new = fmaxf(m_run, m_o);
          const float a = m_new == -INFINITY ? 0.f : ex2_approx((m_run - m_new) * sl2);
          const float bsc = m_new == -INFINITY ? 0.f : ex2_approx((m_o - m_new) * sl2);
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            const float2 o2 = *reinterpret_cast<const float2*>(slot + g * D + n * 8 + 2 * t4);
            o_acc[n][0] = o_acc[n][0] * a + o2.x * bsc;
            o_acc[n][1] = o_acc[n][1] * a + o2.y * bsc;
          }
          l_run = l_run * a + l_o * bsc;
          m_run = m_new;
        }
        named_bar_sync(1, kConsumerWarps * 32);
      }
    }
    if (unit == 0 && g < G) {
      const bool whole = (j0 == 0) && (j1 == ntiles);
      if (whole) {
        const float inv = 1.f / l_run;
        T* orow = obase + (int64_t)b * p.out_stride + (int64_t)(kv_head * G + g) * D;
#pragma unroll
        for (int n = 0; n < NT; ++n)
          *reinterpret_cast<uint32_t*>(orow + n * 8 + 2 * t4) =
              pack2<T>(o_acc[n][0] * inv, o_acc[n][1] * inv);
      } else {
        const int slot = blockIdx.x * 2 + (seg_index == 0 ? 0 : 1);
        const int rows = HKV_CTA * G;
        const int r = head_local * G + g;
        float* po = p.part_o + ((int64_t)slot * rows + r) * D;
#pragma unroll
        for (int n = 0; n < NT; ++n)
          *reinterpret_cast<float2*>(po + n * 8 + 2 * t4) = make_float2(o_acc[n][0], o_acc[n][1]);
        if (t4 == 0) {
          float* pml = p.part_ml + ((int64_t)slot * rows + r) * 2;
          pml[0] = m_run * sl2;  // running max in the exp2 domain
          pml[1] = l_run;
        }
      }
    }
    if (j1 < ntiles && item_t0 / per == (int)blockIdx.x) own_item = it;
    t += j1 - j0;
    ++seg_index;
    do { ++it; } while (it < items && prefix[it + 1] == prefix[it]);
  }
  if (threadIdx.x == 0) p.owner[blockIdx.x] = own_item;   // read by the combine kernel
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) !=
            cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

struct MapKey {
  const void* ptr;
  int64_t rows, row_bytes;
  int chunks, rows_box, chunks_box, dtype;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && row_bytes == o.row_bytes && chunks == o.chunks &&
           rows_box == o.rows_box && chunks_box == o.chunks_box && dtype == o.dtype;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    h ^= std::hash<int64_t>()(k.rows * 1315423911ll + k.row_bytes * 40503ll +
                              k.chunks * 2654435761ll + k.rows_box * 97 + k.chunks_box * 13 + k.dtype);
    return h;
  }
};

// Tensor map over a row-major matrix [rows][>= chunks*128 bytes] (row pitch row_bytes) of 16-bit
// (or, dtype == kPanelU8, 8-bit) elements viewed as {128 B, rows, chunks} with strides
// {row_bytes, 128 B}; box {128 B, rows_box, chunks_box}, 128-byte swizzle: every box lands in
// shared memory as [chunk][row][128 B] SW128 panels.
int get_panel_map(const void* cache, int64_t rows, int64_t row_bytes, int chunks, int rows_box,
                  int chunks_box, int dtype, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache_maps;
  MapKey key{cache, rows, row_bytes, chunks, rows_box, chunks_box, dtype};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache_maps.find(key);
  if (it != cache_maps.end()) {
    *out = it->second;
    return B200_OK;
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return B200_ERR_DRIVER;
  const cuuint32_t inner = dtype == kPanelU8 ? 128 : 64;   // elements per 128-byte panel row
  cuuint64_t gdim[3] = {inner, (cuuint64_t)rows, (cuuint64_t)chunks};
  cuuint64_t gstride[2] = {(cuuint64_t)row_bytes, 128};
  cuuint32_t box[3] = {inner, (cuuint32_t)rows_box, (cuuint32_t)chunks_box};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  const CUtensorMapDataType elem = dtype == kPanelU8         ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                   : dtype == B200_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                              : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(&m, elem, 3, const_cast<void*>(cache), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return B200_ERR_DRIVER;
  if (cache_maps.size() > 4096) cache_maps.clear();
  cache_maps.emplace(key, m);
  *out = m;
  return B200_OK;
}

template <typename T, int D, int HKV_CTA>
static int decode_launch_t(const CUtensorMap& mk, const CUtensorMap& mv, const DecodeParams& p,
                           int grid, cudaStream_t stream) {
  constexpr int UNITS = 8 / HKV_CTA;
  constexpr int TILE_BYTES = 16 * 8 * D * 2;
  const int items = p.batch * p.hg_count;
  size_t smem = 1024 + 2 * kStages * TILE_BYTES + (UNITS > 1 ? 4 * (8 * D + 16) * 4 : 0) +
                (items + 1) * 4 + 8 + 3 * kStages * 8;
  auto kern = paged_decode_kernel<T, D, HKV_CTA>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  launch_pdl(kern, dim3(grid), dim3(kDecodeThreads), smem, stream, mk, mv, p);
  if (grid > 1) {
    // large batches: items are cut into 2-3 pieces -> one block per boundary handles all rows;
    // small batches: one item may span every CTA -> one block per (boundary, row)
    const int rows = HKV_CTA * p.G;
    const dim3 cgrid(grid - 1, items >= 64 ? 1 : rows);
    launch_pdl(decode_combine_kernel<T, D>, cgrid, dim3(256), 0, stream, p, rows);
  }
  return finish_launch(grid > 1 ? 2 : 1);
}

template <typename T, int D>
static int decode_launch_d(int hkv_cta, const CUtensorMap& mk, const CUtensorMap& mv,
                           const DecodeParams& p, int grid, cudaStream_t stream) {
  switch (hkv_cta) {
    case 8: return decode_launch_t<T, D, 8>(mk, mv, p, grid, stream);
    case 4: return decode_launch_t<T, D, 4>(mk, mv, p, grid, stream);
    case 2: return decode_launch_t<T, D, 2>(mk, mv, p, grid, stream);
    case 1: return decode_launch_t<T, D, 1>(mk, mv, p, grid, stream);
  }
  return B200_ERR_UNSUPPORTED;
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200_paged_decode_attn_workspace_size(int32_t batch, int32_t num_heads,
                                                        int32_t num_kv_heads, int32_t head_dim) {
  if (batch <= 0 || num_heads <= 0 || num_kv_heads <= 0 || head_dim <= 0) return 0;
  const int hkv_cta = hkv_cta_of(num_kv_heads);
  if (!hkv_cta) return 0;
  const int G = num_heads / num_kv_heads;
  const size_t rows = (size_t)hkv_cta * G;
  const size_t items = (size_t)batch * (num_kv_heads / hkv_cta);
  return align_up((items + 3) * 4, 256) + align_up((size_t)kMaxCtas * 4, 256) +
         align_up((size_t)kMaxCtas * 2 * rows * head_dim * 4, 256) +
         align_up((size_t)kMaxCtas * 2 * rows * 2 * 4, 256);
}

extern "C" int b200_paged_decode_attn(void* out, const void* q, const void* k_cache,
                                      const void* v_cache, const int32_t* block_tables,
                                      const int32_t* context_lens, int32_t batch,
                                      int32_t num_heads, int32_t num_kv_heads, int32_t head_dim,
                                      int64_t q_stride, int64_t out_stride, int32_t num_blocks,
                                      int32_t block_size, int32_t block_table_stride, float scale,
                                      void* workspace, size_t workspace_bytes, int32_t num_ctas,
                                      int32_t dtype, b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(batch < 0 || num_heads <= 0 || num_kv_heads <= 0 || num_blocks <= 0 ||
                     block_size <= 0 || block_table_stride <= 0,
                 B200_ERR_BAD_SHAPE);
  B200_RETURN_IF(num_heads % num_kv_heads != 0, B200_ERR_BAD_SHAPE);
  if (batch == 0) return B200_OK;
  B200_RETURN_IF(!out || !q || !k_cache || !v_cache || !block_tables || !context_lens ||
                     !workspace,
                 B200_ERR_NULL_POINTER);
  B200_RETURN_IF(head_dim != 64 && head_dim != 128, B200_ERR_UNSUPPORTED);
  const int G = num_heads / num_kv_heads;
  const int hkv_cta = hkv_cta_of(num_kv_heads);
  B200_RETURN_IF(G > 8 || hkv_cta == 0, B200_ERR_UNSUPPORTED);
  const int tile_tokens = 16 * (8 / hkv_cta);
  B200_RETURN_IF(block_size % 16 != 0, B200_ERR_UNSUPPORTED);
  B200_RETURN_IF(block_size % tile_tokens != 0 && tile_tokens % block_size != 0,
                 B200_ERR_UNSUPPORTED);
  const int hg_count = num_kv_heads / hkv_cta;
  B200_RETURN_IF((int64_t)batch * hg_count > kMaxItems, B200_ERR_UNSUPPORTED);
  B200_RETURN_IF(q_stride < (int64_t)num_heads * head_dim ||
                     out_stride < (int64_t)num_heads * head_dim,
                 B200_ERR_BAD_SHAPE);
  B200_RETURN_IF(q_stride % 2 != 0 || out_stride % 2 != 0 ||
                     (reinterpret_cast<uintptr_t>(q) & 3) || (reinterpret_cast<uintptr_t>(out) & 3),
                 B200_ERR_MISALIGNED);
  B200_RETURN_IF(!aligned16(k_cache) || !aligned16(v_cache) || !aligned16(workspace),
                 B200_ERR_MISALIGNED);
  B200_RETURN_IF(workspace_bytes <
                     b200_paged_decode_attn_workspace_size(batch, num_heads, num_kv_heads, head_dim),
                 B200_ERR_WORKSPACE);

  DecodeParams p;
  p.q = q;
  p.out = out;
  p.block_tables = block_tables;
  p.context_lens = context_lens;
  const size_t rows = (size_t)hkv_cta * G;
  const size_t items = (size_t)batch * hg_count;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  p.header = reinterpret_cast<int32_t*>(ws);
  ws += align_up((items + 3) * 4, 256);
  p.owner = reinterpret_cast<int32_t*>(ws);
  ws += align_up((size_t)kMaxCtas * 4, 256);
  p.part_o = reinterpret_cast<float*>(ws);
  ws += align_up((size_t)kMaxCtas * 2 * rows * head_dim * 4, 256);
  p.part_ml = reinterpret_cast<float*>(ws);
  p.q_stride = q_stride;
  p.out_stride = out_stride;
  p.batch = batch;
  p.H = num_heads;
  p.Hkv = num_kv_heads;
  p.G = G;
  p.block_size = block_size;
  p.bt_stride = block_table_stride;
  p.hg_count = hg_count;
  p.rows_box = block_size < tile_tokens ? block_size : tile_tokens;
  p.scale_log2 = scale * 1.4426950408889634f;

  const int chunks = num_kv_heads * head_dim / 64;
  const int chunks_box = hkv_cta * head_dim / 64;
  CUtensorMap mk, mv;
  int rc = get_panel_map(k_cache, (int64_t)num_blocks * block_size, (int64_t)chunks * 128, chunks,
                         p.rows_box, chunks_box, dtype, &mk);
  if (rc != B200_OK) return rc;
  rc = get_panel_map(v_cache, (int64_t)num_blocks * block_size, (int64_t)chunks * 128, chunks,
                     p.rows_box, chunks_box, dtype, &mv);
  if (rc != B200_OK) return rc;

  int grid = num_ctas > 0 ? num_ctas : num_sms();
  if (grid > kMaxCtas) grid = kMaxCtas;
  if (dtype == B200_DTYPE_BF16) {
    return head_dim == 128
               ? decode_launch_d<__nv_bfloat16, 128>(hkv_cta, mk, mv, p, grid, stream)
               : decode_launch_d<__nv_bfloat16, 64>(hkv_cta, mk, mv, p, grid, stream);
  }
  return head_dim == 128 ? decode_launch_d<__half, 128>(hkv_cta, mk, mv, p, grid, stream)
                         : decode_launch_d<__half, 64>(hkv_cta, mk, mv, p, grid, stream);
}
