// decode_attn_fp8.cu -- K6 over an FP8 (e4m3) paged KV cache, and the K1/K2 scatter that fills it.
//
// SURVEY.md section 8 row f3 ("FP8 KV-cache variant of K1/K6"): an extension of the path, not
// something the reference has -- its caches are always the model dtype (engine/model_runner.py:
// 218-229). Semantics: cache element = e4m3(x / scale) (round to nearest even, saturating at
// +-448), one fp32 scale per tensor for K and one for V; attention runs over the de-quantised
// values  K = k8 * k_scale, V = v8 * v_scale  with fp32 softmax/accumulation exactly like the
// 16-bit kernel (decode_attn.cu). Halves the bytes the HBM-bound decode step streams.
//
// Kernel structure == paged_decode_kernel (persistent stream-K over KV tiles, TMA producer warp,
// 8 consumer warps, one kv head x one 16-token unit per warp and tile, GQA group on the M
// dimension of mma.sync m16n8k16), with these differences:
//  * the cache is described to TMA as {128 B, rows, chunks = Hkv*D/128} of UINT8, so a panel row
//    is 128 bytes = 128 elements; 6 stages of 16 KiB (K) + 16 KiB (V) keep the same 192 KiB in
//    flight per SM;
//  * K/V fragments are de-quantised in registers to f16 (exact: e4m3 is a subset of f16) with
//    cvt.rn.f16x2.e4m3x2 and fed to the f16 MMA; Q is converted to f16 once per item, P is packed
//    as f16. The contraction index is permuted so that no data has to be shuffled:
//      QK^T: lane (g, t4) reads the D/4 contiguous bytes [t4*D/4, (t4+1)*D/4) of token row g; byte
//            4*ks+{0,1} / 4*ks+{2,3} of that run are k-slots {2t4,2t4+1} / {2t4+8,2t4+9} of MMA
//            step ks, and Q's A fragment is loaded with the same (t4, ks) -> d mapping;
//      PV:   ldmatrix.trans (b16 granularity) hands lane (g, t4) the byte PAIRS (d = 16c+2g,
//            16c+2g+1) of tokens 2t4, 2t4+1, 2t4+8, 2t4+9; two byte-permutes split them into an
//            "even d" and an "odd d" B fragment, i.e. two n-tiles whose column n = g means
//            d = 16c+2g and d = 16c+2g+1. A lane therefore ends up owning the 4 consecutive
//            outputs d = 16c + 4t4 + {0,1,2,3} of every 16-wide column group c.
//  * k_scale is folded into the softmax scale, v_scale into the final normalisation.
#include "decode_common.cuh"

namespace b200 {

constexpr int kStagesQ8 = 6;

__device__ __forceinline__ uint32_t e4m3x2_to_f16x2(uint32_t two_bytes) {
  uint32_t r;
  asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(r) : "h"((uint16_t)two_bytes));
  return r;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(addr));
  return v;
}
// two consecutive T elements (one 32-bit word) -> f16x2
template <typename T>
__device__ __forceinline__ uint32_t to_f16x2(uint32_t w);
template <>
__device__ __forceinline__ uint32_t to_f16x2<__half>(uint32_t w) {
  return w;
}
template <>
__device__ __forceinline__ uint32_t to_f16x2<__nv_bfloat16>(uint32_t w) {
  const float2 f = unpack2<__nv_bfloat16>(w);
  // |q| beyond the f16 range cannot come out of a normalised projection; clamp instead of inf
  return pack2<__half>(fminf(fmaxf(f.x, -65504.f), 65504.f), fminf(fmaxf(f.y, -65504.f), 65504.f));
}

template <typename T, int D, int HKV_CTA>
__global__ void __launch_bounds__(kDecodeThreads, 1)
paged_decode_fp8_kernel(const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const DecodeParams p,
                        const float v_scale) {
  constexpr int UNITS = 8 / HKV_CTA;          // 16-token units per tile (per head)
  constexpr int TILE_TOKENS = 16 * UNITS;
  constexpr int TILE_BYTES = 16 * 8 * D;      // bytes of K (or V) per tile, all variants
  constexpr int KSTEPS = D / 16;              // MMA k-steps of QK^T == 32-bit words per lane and row
  constexpr int NT = D / 8;                   // 8-wide output column tiles
  constexpr int NC = D / 16;                  // 16-wide output column groups (an even + an odd tile)
  constexpr int MERGE_FLOATS = 8 * D + 16;    // one warp's (O rows, m, l)
  static_assert(HKV_CTA * D >= 128, "a CTA's head group must span whole 128-byte chunks");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* kbuf = smem;                                    // [kStagesQ8][TILE_BYTES]
  uint8_t* vbuf = smem + kStagesQ8 * TILE_BYTES;           // [kStagesQ8][TILE_BYTES]
  float* merge = reinterpret_cast<float*>(smem + 2 * kStagesQ8 * TILE_BYTES);  // [4][MERGE_FLOATS]
  int* prefix = reinterpret_cast<int*>(merge + (UNITS > 1 ? 4 * MERGE_FLOATS : 0));
  const int items = p.batch * p.hg_count;
  uint64_t* bars = reinterpret_cast<uint64_t*>(
      (reinterpret_cast<uintptr_t>(prefix + items + 1) + 7) & ~uintptr_t(7));
  uint64_t* full_k = bars;
  uint64_t* full_v = bars + kStagesQ8;
  uint64_t* empty = bars + 2 * kStagesQ8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_wait();

  // ---- tiles per item -> exclusive prefix sum in shared memory
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int len = p.context_lens[i / p.hg_count];
    prefix[i + 1] = len > 0 ? (len + TILE_TOKENS - 1) / TILE_TOKENS : 0;
  }
  if (threadIdx.x == 0) {
    prefix[0] = 0;
    for (int s = 0; s < kStagesQ8; ++s) {
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
  const int op_bytes = p.rows_box * HKV_CTA * D;

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
        chunk0 = hg * (HKV_CTA * D / 128);
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
        if (++stage == kStagesQ8) stage = 0, phase ^= 1;
      }
    }
    return;
  }

  // =========================== consumers ===========================
  const int head_local = warp % HKV_CTA;
  const int unit = warp / HKV_CTA;
  const int g = lane >> 2, t4 = lane & 3;
  const int G = p.G;
  const float sl2 = p.scale_log2;    // softmax scale * log2(e) * k_scale
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
  const int head_byte = head_local * D;              // byte offset of this head in the CTA's run
  const int hb16 = (head_byte & 127) >> 4;           // first 16-byte column of the head in its chunk
  const int unit_base = op_idx * op_bytes + (head_byte >> 7) * panel_bytes + r0 * 128;
  // K: lane (g, t4) reads 16-byte columns hb16 + t4*KSTEPS/4 + i of token rows g and g + 8
  const int k_c16 = hb16 + t4 * (KSTEPS / 4);
  // V (transposed ldmatrix): matrices (tok0-7,c) (tok8-15,c) (tok0-7,c+1) (tok8-15,c+1)
  const int v_tok = (((lane >> 3) & 1) << 3) + (lane & 7);
  const int v_dc = lane >> 4;

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

    // ---- Q fragments (f16): row = query head g of the GQA group, k-slots permuted as above
    uint32_t qa[KSTEPS][2];
    {
      const T* qrow = qbase + (int64_t)b * p.q_stride + (int64_t)(kv_head * G + g) * D + t4 * (D / 4);
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        uint2 w = make_uint2(0u, 0u);
        if (g < G) w = *reinterpret_cast<const uint2*>(qrow + 4 * ks);
        qa[ks][0] = to_f16x2<T>(w.x);
        qa[ks][1] = to_f16x2<T>(w.y);
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
        uint32_t kw[2][KSTEPS];     // token rows g (n = 0) and g + 8 (n = 1): word ks feeds MMA step ks
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
          const uint32_t rowaddr = kb + (n * 8 + g) * 128;
#pragma unroll
          for (int i = 0; i < KSTEPS / 4; ++i) {
            const uint4 x = lds128(rowaddr + (((k_c16 + i) ^ g) << 4));
            kw[n][4 * i] = x.x, kw[n][4 * i + 1] = x.y, kw[n][4 * i + 2] = x.z, kw[n][4 * i + 3] = x.w;
          }
        }
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {   // the two accumulator chains interleave
#pragma unroll
          for (int n = 0; n < 2; ++n)
            mma_16816<__half>(s[n], qa[ks][0], 0u, qa[ks][1], 0u, e4m3x2_to_f16x2(kw[n][ks] & 0xffffu),
                              e4m3x2_to_f16x2(kw[n][ks] >> 16));
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
        const uint32_t pa0 = pack2<__half>(p00, p01), pa2 = pack2<__half>(p10, p11);

        mbar_wait(&full_v[stage], phase);
        uint8_t* vstage = vbuf + stage * TILE_BYTES + unit_base;
        if (valid < 16) {
          // rows past the context hold stale cache bytes (0x7f/0xff decode to NaN): zero this
          // head's bytes of those rows so that 0 * garbage cannot poison the accumulator.
          const int nz = (16 - valid) * NC;
          for (int idx = lane; idx < nz; idx += 32) {
            const int row = valid + idx / NC;
            const int c = hb16 + idx % NC;
            *reinterpret_cast<uint4*>(vstage + row * 128 + ((c ^ (row & 7)) << 4)) =
                make_uint4(0, 0, 0, 0);
          }
          __syncwarp();
        }
        const uint32_t vb = smem_u32(vstage) + v_tok * 128;
#pragma unroll
        for (int jp = 0; jp < D / 32; ++jp) {
          uint32_t r0_, r1_, r2_, r3_;
          ldsm_x4_trans(vb + (((hb16 + 2 * jp + v_dc) ^ (v_tok & 7)) << 4), r0_, r1_, r2_, r3_);
          const uint32_t e0 = __byte_perm(r0_, r1_, 0x6420), o0 = __byte_perm(r0_, r1_, 0x7531);
          const uint32_t e1 = __byte_perm(r2_, r3_, 0x6420), o1 = __byte_perm(r2_, r3_, 0x7531);
          mma_16816<__half>(o_acc[4 * jp + 0], pa0, 0u, pa2, 0u, e4m3x2_to_f16x2(e0 & 0xffffu),
                            e4m3x2_to_f16x2(e0 >> 16));
          mma_16816<__half>(o_acc[4 * jp + 1], pa0, 0u, pa2, 0u, e4m3x2_to_f16x2(o0 & 0xffffu),
                            e4m3x2_to_f16x2(o0 >> 16));
          mma_16816<__half>(o_acc[4 * jp + 2], pa0, 0u, pa2, 0u, e4m3x2_to_f16x2(e1 & 0xffffu),
                            e4m3x2_to_f16x2(e1 >> 16));
          mma_16816<__half>(o_acc[4 * jp + 3], pa0, 0u, pa2, 0u, e4m3x2_to_f16x2(o1 & 0xffffu),
                            e4m3x2_to_f16x2(o1 >> 16));
        }
        if (valid < 16) fence_proxy_async_smem();
      } else {
        mbar_wait(&full_v[stage], phase);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == kStagesQ8) stage = 0, phase ^= 1;
    }

    // ---- finalise this segment. Lane (g, t4) owns, for every 16-wide column group c, the four
    // consecutive outputs d = 16c + 4t4 + {0,1,2,3} = {o_acc[2c][0], o_acc[2c+1][0], o_acc[2c][1],
    // o_acc[2c+1][1]} of query head g.
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
    if (UNITS > 1) {
      // tree-merge the warps that share a kv head (different 16-token units of each tile)
#pragma unroll
      for (int stride = UNITS / 2; stride >= 1; stride >>= 1) {
        float* slot = merge + ((unit % stride) * HKV_CTA + head_local) * MERGE_FLOATS;
        if (unit >= stride && unit < 2 * stride) {
#pragma unroll
          for (int c = 0; c < NC; ++c)
            *reinterpret_cast<float4*>(slot + g * D + 16 * c + 4 * t4) =
                make_float4(o_acc[2 * c][0], o_acc[2 * c + 1][0], o_acc[2 * c][1], o_acc[2 * c + 1][1]);
          if (t4 == 0) slot[8 * D + g] = m_run, slot[8 * D + 8 + g] = l_run;
        }
        named_bar_sync(1, kConsumerWarps * 32);
        if (unit < stride) {
          const float m_o = slot[8 * D + g], l_o = slot[8 * D + 8 + g];
          const float m_new = fmaxf(m_run, m_o);
          const float a = m_new == -INFINITY ? 0.f : ex2_approx((m_run - m_new) * sl2);
          const float bsc = m_new == -INFINITY ? 0.f : ex2_approx((m_o - m_new) * sl2);
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const float4 o4 = *reinterpret_cast<const float4*>(slot + g * D + 16 * c + 4 * t4);
            o_acc[2 * c][0] = o_acc[2 * c][0] * a + o4.x * bsc;
            o_acc[2 * c + 1][0] = o_acc[2 * c + 1][0] * a + o4.y * bsc;
            o_acc[2 * c][1] = o_acc[2 * c][1] * a + o4.z * bsc;
            o_acc[2 * c + 1][1] = o_acc[2 * c + 1][1] * a + o4.w * bsc;
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
        const float inv = v_scale / l_run;
        T* orow = obase + (int64_t)b * p.out_stride + (int64_t)(kv_head * G + g) * D + 4 * t4;
#pragma unroll
        for (int c = 0; c < NC; ++c)
          *reinterpret_cast<uint2*>(orow + 16 * c) =
              make_uint2(pack2<T>(o_acc[2 * c][0] * inv, o_acc[2 * c + 1][0] * inv),
                         pack2<T>(o_acc[2 * c][1] * inv, o_acc[2 * c + 1][1] * inv));
      } else {
        const int slot = blockIdx.x * 2 + (seg_index == 0 ? 0 : 1);
        const int rows = HKV_CTA * G;
        const int r = head_local * G + g;
        float* po = p.part_o + ((int64_t)slot * rows + r) * D + 4 * t4;
#pragma unroll
        for (int c = 0; c < NC; ++c)
          *reinterpret_cast<float4*>(po + 16 * c) =
              make_float4(o_acc[2 * c][0] * v_scale, o_acc[2 * c + 1][0] * v_scale,
                          o_acc[2 * c][1] * v_scale, o_acc[2 * c + 1][1] * v_scale);
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

// ------------------------------------------------------------------ K1/K2 into an fp8 cache
// One 16-byte load (8 elements) per thread and tensor -> 8 e4m3 bytes (one 8-byte store).
template <typename T, bool kSlot2D>
__global__ void __launch_bounds__(256)
kv_scatter_fp8_kernel(const uint4* __restrict__ key, const uint4* __restrict__ value,
                      uint2* __restrict__ k_cache, uint2* __restrict__ v_cache,
                      const int32_t* __restrict__ slots, int num_tokens, int vec_per_row,
                      int64_t key_stride_vec, int64_t value_stride_vec, int num_blocks,
                      int block_size, float k_scale_inv, float v_scale_inv) {
  pdl_wait();
  const int64_t total = (int64_t)num_tokens * vec_per_row * 2;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int vec = (int)(idx % vec_per_row);
    const int64_t rest = idx / vec_per_row;
    const int which = (int)(rest & 1);  // 0 = K, 1 = V
    const int token = (int)(rest >> 1);
    int64_t slot;
    if (kSlot2D) {
      const int blk = slots[2 * token], off = slots[2 * token + 1];
      if (blk < 0 || blk >= num_blocks || off < 0 || off >= block_size) continue;
      slot = (int64_t)blk * block_size + off;
    } else {
      const int s = slots[token];
      if (s < 0 || s >= (int64_t)num_blocks * block_size) continue;  // -1 == "no slot"
      slot = s;
    }
    const uint4 x = which == 0 ? ld_nc_v4(key + token * key_stride_vec + vec)
                               : ld_nc_v4(value + token * value_stride_vec + vec);
    const float sc = which == 0 ? k_scale_inv : v_scale_inv;
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
    uint32_t o[2] = {0u, 0u};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack2<T>(w[j]);
      uint16_t two;
      // d = {e4m3(a), e4m3(b)} with a in the upper byte: pass (hi, lo) = (f.y, f.x)
      asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(two) : "f"(__fmul_rn(f.y, sc)), "f"(__fmul_rn(f.x, sc)));
      o[j >> 1] |= (uint32_t)two << (16 * (j & 1));
    }
    (which == 0 ? k_cache : v_cache)[slot * vec_per_row + vec] = make_uint2(o[0], o[1]);
  }
}

// ------------------------------------------------------------------ paged fp8 cache -> packed 16-bit K/V
// Prefix-hit / chunked prefill over an e4m3 cache: the tcgen05 prefill kernel multiplies 16-bit
// operands, so the cached tokens of every sequence are first gathered through its block table and
// de-quantised (x = e4m3 * scale, exact in fp32, one rounding to T) into packed [total_k, Hkv*D]
// buffers, which the packed-K/V prefill path then reads. One 8-byte load (8 e4m3) -> one 16-byte store.
template <typename T>
__global__ void __launch_bounds__(256)
gather_dequant_fp8_kernel(uint4* __restrict__ k_out, uint4* __restrict__ v_out,
                          const uint2* __restrict__ k_cache, const uint2* __restrict__ v_cache,
                          const int32_t* __restrict__ block_tables, const int32_t* __restrict__ cu_k,
                          int bt_stride, int num_blocks, int block_size, int vec_per_row, float k_scale,
                          float v_scale) {
  pdl_wait();
  const int seq = blockIdx.y;
  const int k0 = cu_k[seq], len = cu_k[seq + 1] - k0;
  const int64_t total = (int64_t)len * vec_per_row * 2;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int vec = (int)(idx % vec_per_row);
    const int64_t rest = idx / vec_per_row;
    const int which = (int)(rest & 1);
    const int tok = (int)(rest >> 1);
    const int blk = block_tables[(int64_t)seq * bt_stride + tok / block_size];
    uint4 o = make_uint4(0, 0, 0, 0);
    if (blk >= 0 && blk < num_blocks) {
      const int64_t slot = (int64_t)blk * block_size + tok % block_size;
      const uint2 x = (which == 0 ? k_cache : v_cache)[slot * vec_per_row + vec];
      const float sc = which == 0 ? k_scale : v_scale;
      const uint32_t w[2] = {x.x, x.y};
      uint32_t r[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t h2 = e4m3x2_to_f16x2((w[j >> 1] >> (16 * (j & 1))) & 0xffffu);
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h2));
        r[j] = pack2<T>(__fmul_rn(f.x, sc), __fmul_rn(f.y, sc));
      }
      o = make_uint4(r[0], r[1], r[2], r[3]);
    }
    (which == 0 ? k_out : v_out)[(int64_t)(k0 + tok) * vec_per_row + vec] = o;
  }
}

template <typename T, int D>
static int decode_fp8_launch_d(int hkv_cta, const CUtensorMap& mk, const CUtensorMap& mv,
                               const DecodeParams& p, float v_scale, int grid, cudaStream_t stream) {
  const int items = p.batch * p.hg_count;
  const int units = 8 / hkv_cta;
  const size_t smem = 1024 + 2 * kStagesQ8 * (16 * 8 * D) + (units > 1 ? 4 * (8 * D + 16) * 4 : 0) +
                      (items + 1) * 4 + 8 + 3 * kStagesQ8 * 8;
  void (*kern)(CUtensorMap, CUtensorMap, DecodeParams, float) = nullptr;
  switch (hkv_cta) {
    case 8: kern = paged_decode_fp8_kernel<T, D, 8>; break;
    case 4: kern = paged_decode_fp8_kernel<T, D, 4>; break;
    case 2: kern = paged_decode_fp8_kernel<T, D, 2>; break;
    case 1:
      if constexpr (D >= 128) kern = paged_decode_fp8_kernel<T, D, 1>;
      break;
  }
  if (kern == nullptr) return B200_ERR_UNSUPPORTED;
  static bool attr_set[9] = {};   // per instantiation (T, D) and head-group size
  if (!attr_set[hkv_cta]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set[hkv_cta] = true;
  }
  launch_pdl(kern, dim3(grid), dim3(kDecodeThreads), smem, stream, mk, mv, p, v_scale);
  if (grid > 1) {
    const int rows = hkv_cta * p.G;
    const dim3 cgrid(grid - 1, items >= 64 ? 1 : rows);
    launch_pdl(decode_combine_kernel<T, D>, cgrid, dim3(256), 0, stream, p, rows);
  }
  return finish_launch(grid > 1 ? 2 : 1);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_paged_decode_attn_fp8(void* out, const void* q, const void* k_cache,
                                          const void* v_cache, const int32_t* block_tables,
                                          const int32_t* context_lens, int32_t batch,
                                          int32_t num_heads, int32_t num_kv_heads, int32_t head_dim,
                                          int64_t q_stride, int64_t out_stride, int32_t num_blocks,
                                          int32_t block_size, int32_t block_table_stride, float scale,
                                          float k_scale, float v_scale, void* workspace,
                                          size_t workspace_bytes, int32_t num_ctas, int32_t dtype,
                                          b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(batch < 0 || num_heads <= 0 || num_kv_heads <= 0 || num_blocks <= 0 ||
                     block_size <= 0 || block_table_stride <= 0,
                 B200_ERR_BAD_SHAPE);
  B200_RETURN_IF(num_heads % num_kv_heads != 0, B200_ERR_BAD_SHAPE);
  B200_RETURN_IF(!(k_scale > 0.f) || !(v_scale > 0.f), B200_ERR_BAD_SHAPE);
  if (batch == 0) return B200_OK;
  B200_RETURN_IF(!out || !q || !k_cache || !v_cache || !block_tables || !context_lens ||
                     !workspace,
                 B200_ERR_NULL_POINTER);
  B200_RETURN_IF(head_dim != 64 && head_dim != 128, B200_ERR_UNSUPPORTED);
  const int G = num_heads / num_kv_heads;
  const int hkv_cta = hkv_cta_of(num_kv_heads);
  B200_RETURN_IF(G > 8 || hkv_cta == 0, B200_ERR_UNSUPPORTED);
  B200_RETURN_IF(hkv_cta * head_dim < 128, B200_ERR_UNSUPPORTED);   // D=64 needs >= 2 kv heads
  const int tile_tokens = 16 * (8 / hkv_cta);
  B200_RETURN_IF(block_size % 16 != 0, B200_ERR_UNSUPPORTED);
  B200_RETURN_IF(block_size % tile_tokens != 0 && tile_tokens % block_size != 0,
                 B200_ERR_UNSUPPORTED);
  const int hg_count = num_kv_heads / hkv_cta;
  B200_RETURN_IF((int64_t)batch * hg_count > kMaxItems, B200_ERR_UNSUPPORTED);
  B200_RETURN_IF(q_stride < (int64_t)num_heads * head_dim ||
                     out_stride < (int64_t)num_heads * head_dim,
                 B200_ERR_BAD_SHAPE);
  B200_RETURN_IF(q_stride % 4 != 0 || out_stride % 4 != 0 ||
                     (reinterpret_cast<uintptr_t>(q) & 7) || (reinterpret_cast<uintptr_t>(out) & 7),
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
  p.scale_log2 = scale * k_scale * 1.4426950408889634f;

  const int chunks = num_kv_heads * head_dim / 128;
  const int chunks_box = hkv_cta * head_dim / 128;
  CUtensorMap mk, mv;
  int rc = get_panel_map(k_cache, (int64_t)num_blocks * block_size, (int64_t)chunks * 128, chunks,
                         p.rows_box, chunks_box, kPanelU8, &mk);
  if (rc != B200_OK) return rc;
  rc = get_panel_map(v_cache, (int64_t)num_blocks * block_size, (int64_t)chunks * 128, chunks,
                     p.rows_box, chunks_box, kPanelU8, &mv);
  if (rc != B200_OK) return rc;

  int grid = num_ctas > 0 ? num_ctas : num_sms();
  if (grid > kMaxCtas) grid = kMaxCtas;
  if (dtype == B200_DTYPE_BF16)
    return head_dim == 128
               ? decode_fp8_launch_d<__nv_bfloat16, 128>(hkv_cta, mk, mv, p, v_scale, grid, stream)
               : decode_fp8_launch_d<__nv_bfloat16, 64>(hkv_cta, mk, mv, p, v_scale, grid, stream);
  return head_dim == 128 ? decode_fp8_launch_d<__half, 128>(hkv_cta, mk, mv, p, v_scale, grid, stream)
                         : decode_fp8_launch_d<__half, 64>(hkv_cta, mk, mv, p, v_scale, grid, stream);
}

extern "C" int b200_reshape_and_cache_fp8(const void* key, const void* value, void* k_cache,
                                          void* v_cache, const int32_t* slots, int32_t slot_is_2d,
                                          int32_t num_tokens, int32_t num_kv_heads, int32_t head_dim,
                                          int64_t key_stride, int64_t value_stride,
                                          int32_t num_blocks, int32_t block_size, float k_scale,
                                          float v_scale, int32_t dtype, b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(num_tokens < 0 || num_kv_heads <= 0 || head_dim <= 0 || num_blocks <= 0 ||
                     block_size <= 0,
                 B200_ERR_BAD_SHAPE);
  B200_RETURN_IF(!(k_scale > 0.f) || !(v_scale > 0.f), B200_ERR_BAD_SHAPE);
  if (num_tokens == 0) return B200_OK;
  B200_RETURN_IF(!key || !value || !k_cache || !v_cache || !slots, B200_ERR_NULL_POINTER);
  const int64_t row = (int64_t)num_kv_heads * head_dim;  // elements
  B200_RETURN_IF(row % 8 != 0 || key_stride % 8 != 0 || value_stride % 8 != 0,
                 B200_ERR_MISALIGNED);
  B200_RETURN_IF(key_stride < row || value_stride < row, B200_ERR_BAD_SHAPE);
  B200_RETURN_IF(!aligned16(key) || !aligned16(value) || !aligned16(k_cache) ||
                     !aligned16(v_cache),
                 B200_ERR_MISALIGNED);
  const int vec_per_row = (int)(row / 8);
  const int64_t total = (int64_t)num_tokens * vec_per_row * 2;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  const float ki = 1.0f / k_scale, vi = 1.0f / v_scale;
#define B200_LAUNCH_SCATTER8(T, S2D)                                                              \
  launch_pdl(kv_scatter_fp8_kernel<T, S2D>, dim3(grid), dim3(256), 0, stream, (const uint4*)key,  \
             (const uint4*)value, (uint2*)k_cache, (uint2*)v_cache, slots, num_tokens, vec_per_row, \
             key_stride / 8, value_stride / 8, num_blocks, block_size, ki, vi)
  if (dtype == B200_DTYPE_BF16) {
    if (slot_is_2d) B200_LAUNCH_SCATTER8(__nv_bfloat16, true);
    else B200_LAUNCH_SCATTER8(__nv_bfloat16, false);
  } else {
    if (slot_is_2d) B200_LAUNCH_SCATTER8(__half, true);
    else B200_LAUNCH_SCATTER8(__half, false);
  }
#undef B200_LAUNCH_SCATTER8
  return finish_launch();
}

extern "C" int b200_gather_dequant_fp8(void* k_out, void* v_out, const void* k_cache, const void* v_cache,
                                       const int32_t* block_tables, const int32_t* cu_seqlens_k,
                                       int32_t num_seqs, int32_t max_seqlen_k, int32_t block_table_stride,
                                       int32_t num_blocks, int32_t block_size, int32_t num_kv_heads,
                                       int32_t head_dim, float k_scale, float v_scale, int32_t dtype,
                                       b200_stream_t stream) {
  using namespace b200;
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(num_seqs < 0 || max_seqlen_k < 0 || num_kv_heads <= 0 || head_dim <= 0 ||
                     block_table_stride <= 0 || num_blocks <= 0 || block_size <= 0,
                 B200_ERR_BAD_SHAPE);
  B200_RETURN_IF((num_kv_heads * head_dim) % 8 != 0, B200_ERR_UNSUPPORTED);
  B200_RETURN_IF(!(k_scale > 0.f) || !(v_scale > 0.f), B200_ERR_BAD_SHAPE);
  if (num_seqs == 0 || max_seqlen_k == 0) return B200_OK;
  B200_RETURN_IF(!k_out || !v_out || !k_cache || !v_cache || !block_tables || !cu_seqlens_k,
                 B200_ERR_NULL_POINTER);
  B200_RETURN_IF(!aligned16(k_out) || !aligned16(v_out) || (reinterpret_cast<uintptr_t>(k_cache) & 7) ||
                     (reinterpret_cast<uintptr_t>(v_cache) & 7),
                 B200_ERR_MISALIGNED);
  B200_RETURN_IF(num_seqs > 65535, B200_ERR_UNSUPPORTED);
  const int vec_per_row = num_kv_heads * head_dim / 8;
  const int64_t per_seq = (int64_t)max_seqlen_k * vec_per_row * 2;
  int gx = (int)((per_seq + 255) / 256);
  if (gx > 148 * 8) gx = 148 * 8;
  if (gx < 1) gx = 1;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == B200_DTYPE_BF16)
    launch_pdl(gather_dequant_fp8_kernel<__nv_bfloat16>, dim3(gx, num_seqs), dim3(256), 0, s, (uint4*)k_out,
               (uint4*)v_out, (const uint2*)k_cache, (const uint2*)v_cache, block_tables, cu_seqlens_k,
               block_table_stride, num_blocks, block_size, vec_per_row, k_scale, v_scale);
  else
    launch_pdl(gather_dequant_fp8_kernel<__half>, dim3(gx, num_seqs), dim3(256), 0, s, (uint4*)k_out,
               (uint4*)v_out, (const uint2*)k_cache, (const uint2*)v_cache, block_tables, cu_seqlens_k,
               block_table_stride, num_blocks, block_size, vec_per_row, k_scale, v_scale);
  return finish_launch();
}
