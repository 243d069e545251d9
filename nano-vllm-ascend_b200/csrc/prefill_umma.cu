// prefill_umma.cu -- K4/K5 on the 5th-generation tensor cores: varlen causal flash-attention
// with tcgen05.mma (UMMA), accumulators in TMEM, operands staged by TMA.
//
// Replaces the same reference call sites as prefill_attn.cu (layers/attention.py:46-59,
// layers/attention_ori.py:63-70). Same semantics: bottom-right causal mask, packed K/V or
// paged K/V through block_tables.
//
// CTA = one (sequence, head, 256 query rows): two 128-row query tiles t = 0, 1.
//   warps 0-3 / 4-7   softmax warpgroup of tile t: one thread per query row (= one TMEM lane);
//                     reads S from TMEM, fp32 online softmax (exp2 domain), rescales O in TMEM
//                     only when the row max grew by > 2^8 (lazy rescale), writes P (bf16) as a
//                     K-major SW128 smem panel (double-buffered)
//   warp 8            TMA producers: lane 0 loads Q then runs a 3-stage K ring, lane 1 a 3-stage V ring
//                     (64-key tiles; K stages recycle after S(j), V stages after PV(j))
//   warp 9            MMA issuer (one lane):  S[t] = Q[t] K^T   (M128 N64 K16 x D/16)
//                                             O[t] += P[t] V    (M128 N=D K16 x 4, V MN-major)
// TMEM (512 columns, one CTA per SM):  O[t] at 128 t,  S[t][buf] at 256 + 128 t + 64 buf.
// S is double-buffered so the tensor pipe computes S(j+1) while the warpgroups are in the
// softmax of tile j; both query tiles share every K/V tile (one L2/TMA read per 256 rows).
// Flops: 4 * Lq * Lk * D per (sequence, head) before causal skipping; tensor-pipe bound.
#include "umma.cuh"

namespace b200 {

struct PrefillUmmaParams {
  const int32_t* cu_q;
  const int32_t* cu_k;
  const int32_t* block_tables;  // nullptr => packed K/V
  void* out;
  int64_t out_stride;
  int H, G, block_size, bt_stride;
  int kv_rows_box;   // rows per K/V TMA op (64 packed; min(block_size, 64) paged)
  int kv_split;      // 1: each TMA op carries a single 64-wide chunk (paged, rows_box < 64)
  float scale_log2;
};

constexpr int kUBM = 128, kUBN = 64, kUStages = 3, kUThreads = 320;

template <typename T, int D>
__global__ void __launch_bounds__(kUThreads, 1)
prefill_umma_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const PrefillUmmaParams p) {
  constexpr int CH = D / 64;                    // 128-byte chunks per head row
  constexpr int QPANEL = kUBM * 128;            // bytes of one Q / P panel (128 rows)
  constexpr int KPANEL = kUBN * 128;            // bytes of one K / V panel (64 rows)
  constexpr int QTILE = CH * QPANEL;
  constexpr int KTILE = CH * KPANEL;
  constexpr int PTILE = QPANEL;                 // P tile: 128 rows x 64 keys
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sQ = smem;                            // [2][QTILE]
  uint8_t* sK = sQ + 2 * QTILE;                  // [kUStages][KTILE]
  uint8_t* sV = sK + kUStages * KTILE;           // [kUStages][KTILE]
  uint8_t* sP = sV + kUStages * KTILE;           // [2 tiles][2 buffers][PTILE]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * PTILE);
  uint64_t* q_full = bars;                       // [1]
  uint64_t* k_full = bars + 1;                   // [kUStages]
  uint64_t* v_full = k_full + kUStages;          // [kUStages]
  uint64_t* k_empty = v_full + kUStages;         // [kUStages] K stage consumed by S(j) MMAs
  uint64_t* v_empty = k_empty + kUStages;        // [kUStages] V stage consumed by PV(j) MMAs
  uint64_t* s_full = v_empty + kUStages;         // [2 tiles][2 buffers]
  uint64_t* p_full = s_full + 4;                 // [2]
  uint64_t* o_full = p_full + 2;                 // [2 tiles][2 P buffers]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 4);

  pdl_wait();
  const int seq = blockIdx.z, head = blockIdx.y;
  const int q0 = p.cu_q[seq], Lq = p.cu_q[seq + 1] - q0;
  const int k0 = p.cu_k[seq], Lk = p.cu_k[seq + 1] - k0;
  const int m_tile = (int)gridDim.x - 1 - (int)blockIdx.x;   // heavy (late) tiles first
  const int m0 = m_tile * 2 * kUBM;
  if (m0 >= Lq) return;                                       // uniform for the whole CTA
  const int kvh = head / p.G;
  const int off = Lk - Lq;
  int n_tiles[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int r0 = m0 + t * kUBM;
    const int n_end = r0 < Lq ? min(Lk, r0 + kUBM + off) : 0;
    n_tiles[t] = n_end > 0 ? (n_end + kUBN - 1) / kUBN : 0;
  }
  const int ntq = (m0 + kUBM < Lq) ? 2 : 1;
  const int NT = max(n_tiles[0], n_tiles[1]);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kUStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&s_full[i], 1);
    for (int t = 0; t < 2; ++t) mbar_init(&p_full[t], kUBM);
    for (int i = 0; i < 4; ++i) mbar_init(&o_full[i], 1);
    mbar_fence_init();
  }
  if (warp == 9) umma::tmem_alloc(tmem_slot, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = *tmem_slot;

  if (warp == 8) {
    // =========================== TMA producer ===========================
    // lane 0: Q once, then the K ring; lane 1: the V ring. K stages are released as soon as the
    // S MMAs that read them retire (one iteration before the matching PV), V stages after PV,
    // so each ring keeps ~3 tiles of lead over its consumer instead of one.
    if (lane < 2) {
      const bool is_k = lane == 0;
      const CUtensorMap* tm = is_k ? &tmK : &tmV;
      tma_prefetch_desc(tm);
      uint64_t pol_keep;
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
      if (is_k) {
        tma_prefetch_desc(&tmQ);
        const uint64_t pol_stream = l2_policy_evict_first();
        mbar_expect_tx(q_full, ntq * QTILE);
        for (int t = 0; t < ntq; ++t)
          tma_load_3d(sQ + t * QTILE, &tmQ, q_full, 0, q0 + m0 + t * kUBM, head * CH, pol_stream);
      }
      uint8_t* ring = is_k ? sK : sV;
      uint64_t* full = is_k ? k_full : v_full;
      uint64_t* empty = is_k ? k_empty : v_empty;
      for (int j = 0; j < NT; ++j) {
        const int s = j % kUStages;
        mbar_wait(&empty[s], ((j / kUStages) & 1) ^ 1);
        const int n0 = j * kUBN;
        if (p.block_tables == nullptr) {
          mbar_expect_tx(&full[s], KTILE);
          tma_load_3d(ring + s * KTILE, tm, &full[s], 0, k0 + n0, kvh * CH, pol_keep);
        } else {
          const int rb = p.kv_rows_box;
          const int nops = min(kUBN / rb, (Lk - n0 + rb - 1) / rb);
          const int per_op = p.kv_split ? CH : 1;
          const int op_bytes = rb * 128 * (p.kv_split ? 1 : CH);
          mbar_expect_tx(&full[s], nops * per_op * op_bytes);
          for (int o = 0; o < nops; ++o) {
            const int tok = n0 + o * rb;
            const int entry = tok / p.block_size;
            const int blk = p.block_tables[(int64_t)seq * p.bt_stride + entry];
            const int row = blk < 0 ? -(1 << 30) : blk * p.block_size + (tok - entry * p.block_size);
            for (int c = 0; c < per_op; ++c) {
              const int dst = (p.kv_split ? c * KPANEL : 0) + o * rb * 128;
              tma_load_3d(ring + s * KTILE + dst, tm, &full[s], 0, row, kvh * CH + c, pol_keep);
            }
          }
        }
      }
    }
  } else if (warp == 9) {
    // =========================== MMA issuer ===========================
    // The whole warp walks the loop (warp-uniform control flow and descriptor arithmetic, so it
    // lives on the uniform datapath); one elected lane issues. Descriptors are loop-invariant
    // bases + compile-time offsets: issuing an MMA costs a couple of integer adds, which matters
    // because a 128x64x16 MMA retires in 32 cycles.
    const uint32_t idesc_s = umma::make_idesc<T>(kUBM, kUBN, 0, 0);   // S = Q K^T, both K-major
    const uint32_t idesc_o = umma::make_idesc<T>(kUBM, D, 0, 1);      // O = P V, V MN-major
    const uint64_t qd0 = umma::desc_kmajor(smem_u32(sQ));
    const uint64_t kd0 = umma::desc_kmajor(smem_u32(sK));
    const uint64_t vd0 = umma::desc_mnmajor(smem_u32(sV), KPANEL);
    const uint64_t pd0 = umma::desc_kmajor(smem_u32(sP));
    const bool leader = umma::elect_one();
    auto issue_s = [&](int j) {
      const int s = j % kUStages;
      mbar_wait(&k_full[s], (j / kUStages) & 1);
      umma::fence_after_sync();
      const uint64_t kd = umma::desc_advance(kd0, s * KTILE);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (j < n_tiles[t]) {
          const uint32_t d = tb + 256 + t * 128 + (j & 1) * 64;
          if (leader) {
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks) {
              const uint64_t ad = umma::desc_advance(qd0, t * QTILE + (ks >> 2) * QPANEL + (ks & 3) * 32);
              const uint64_t bd = umma::desc_advance(kd, (ks >> 2) * KPANEL + (ks & 3) * 32);
              umma::mma_ss(d, ad, bd, idesc_s, ks > 0);
            }
            umma::commit(&s_full[t * 2 + (j & 1)]);
          }
        }
      }
      if (leader) umma::commit(&k_empty[s]);     // K stage free once these S MMAs have retired
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    umma::fence_after_sync();
    if (NT > 0) issue_s(0);
    for (int j = 0; j < NT; ++j) {
      if (j + 1 < NT) issue_s(j + 1);
      const int s = j % kUStages;
      mbar_wait(&v_full[s], (j / kUStages) & 1);
      const uint64_t vd = umma::desc_advance(vd0, s * KTILE);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (j < n_tiles[t]) {
          mbar_wait(&p_full[t], j & 1);
          umma::fence_after_sync();
          if (leader) {
            const uint64_t pd = umma::desc_advance(pd0, (t * 2 + (j & 1)) * PTILE);
#pragma unroll
            for (int kk = 0; kk < kUBN / 16; ++kk)
              umma::mma_ss(tb + t * 128, umma::desc_advance(pd, kk * 32),
                           umma::desc_advance(vd, kk * 16 * 128), idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
            umma::commit(&o_full[t * 2 + (j & 1)]);
          }
        }
      }
      if (leader) umma::commit(&v_empty[s]);
      __syncwarp();
    }
  } else {
    // =========================== softmax warpgroups ===========================
    const int t = warp >> 2;
    const int nt = n_tiles[t];
    const int row = (warp & 3) * 32 + lane;          // row inside the 128-row tile == TMEM lane
    const int qrow = m0 + t * kUBM + row;            // row inside the sequence
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const float sl2 = p.scale_log2;
    // m_ref is the max the accumulators are currently expressed against. It only follows the true
    // running max when that moved by more than 2^kLazy (exp2 domain): P values stay < 2^kLazy,
    // harmless for fp32 accumulation / bf16 P, and the expensive TMEM round trip that rescales O
    // becomes rare after the first tiles.
    constexpr float kLazy = 8.f;
    float m_ref = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nt; ++j) {
      mbar_wait(&s_full[t * 2 + (j & 1)], (j >> 1) & 1);
      umma::fence_after_sync();
      uint32_t sr[2][32];
      const uint32_t s_addr = tb + lane_base + 256 + t * 128 + (j & 1) * 64;
      umma::ld32(s_addr, sr[0]);
      umma::ld32(s_addr + 32, sr[1]);
      umma::wait_ld();
      float sv[64];
#pragma unroll
      for (int c = 0; c < 64; ++c) sv[c] = __uint_as_float(sr[c >> 5][c & 31]);
      const int n0 = j * kUBN;
      // mask (warp-uniform test: does this tile touch the diagonal or the end of the keys?)
      const int warp_row0 = m0 + t * kUBM + (warp & 3) * 32;
      if (n0 + kUBN > Lk || n0 + kUBN - 1 > warp_row0 + off) {
        const int limit = min(Lk - 1, qrow + off) - n0;   // last visible key (tile-relative)
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (c > limit) sv[c] = -INFINITY;
      }
      float mx = sv[0];
#pragma unroll
      for (int c = 1; c < 64; ++c) mx = fmaxf(mx, sv[c]);
      const bool grow = (mx - m_ref) * sl2 > kLazy;       // also true for the first finite max
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? mx : m_ref;
        const float alpha = (m_ref == -INFINITY) ? 0.f : ex2_approx((m_ref - m_new) * sl2);
        if (j > 0) {
          mbar_wait(&o_full[t * 2 + ((j - 1) & 1)], ((j - 1) >> 1) & 1);   // PV(j-1) landed
          umma::fence_after_sync();
#pragma unroll
          for (int c0 = 0; c0 < D; c0 += 32) {
            uint32_t orr[32];
            umma::ld32(tb + lane_base + t * 128 + c0, orr);
            umma::wait_ld();
#pragma unroll
            for (int c = 0; c < 32; ++c) orr[c] = __float_as_uint(__uint_as_float(orr[c]) * alpha);
            umma::st32(tb + lane_base + t * 128 + c0, orr);
          }
          umma::wait_st();
        }
        l_run *= alpha;
        m_ref = m_new;
      }
      const float mb = m_ref == -INFINITY ? 0.f : m_ref * sl2;
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        sv[c] = ex2_approx(fmaf(sv[c], sl2, -mb));
        sum += sv[c];
      }
      l_run += sum;
      // P(j) goes to buffer j&1, last read by PV(j-2)
      if (j >= 2) mbar_wait(&o_full[t * 2 + (j & 1)], ((j >> 1) - 1) & 1);
      uint8_t* prow = sP + (t * 2 + (j & 1)) * PTILE;
#pragma unroll
      for (int c16 = 0; c16 < 8; ++c16) {
        uint4 v;
        v.x = pack2<T>(sv[c16 * 8 + 0], sv[c16 * 8 + 1]);
        v.y = pack2<T>(sv[c16 * 8 + 2], sv[c16 * 8 + 3]);
        v.z = pack2<T>(sv[c16 * 8 + 4], sv[c16 * 8 + 5]);
        v.w = pack2<T>(sv[c16 * 8 + 6], sv[c16 * 8 + 7]);
        *reinterpret_cast<uint4*>(prow + umma::sw128(row, c16)) = v;
      }
      fence_proxy_async_smem();       // generic-proxy smem writes -> visible to the UMMA (async proxy)
      umma::fence_before_sync();
      mbar_arrive(&p_full[t]);
    }
    // ---- epilogue: O / l -> global
    if (m0 + t * kUBM < Lq) {
      T* orow = reinterpret_cast<T*>(p.out) + (int64_t)(q0 + qrow) * p.out_stride + (int64_t)head * D;
      if (nt > 0) {
        mbar_wait(&o_full[t * 2 + ((nt - 1) & 1)], ((nt - 1) >> 1) & 1);
        umma::fence_after_sync();
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
#pragma unroll
        for (int c0 = 0; c0 < D; c0 += 32) {
          uint32_t orr[32];
          umma::ld32(tb + lane_base + t * 128 + c0, orr);
          umma::wait_ld();
          if (qrow < Lq) {
#pragma unroll
            for (int c = 0; c < 32; c += 8) {
              uint4 v;
              v.x = pack2<T>(__uint_as_float(orr[c + 0]) * inv, __uint_as_float(orr[c + 1]) * inv);
              v.y = pack2<T>(__uint_as_float(orr[c + 2]) * inv, __uint_as_float(orr[c + 3]) * inv);
              v.z = pack2<T>(__uint_as_float(orr[c + 4]) * inv, __uint_as_float(orr[c + 5]) * inv);
              v.w = pack2<T>(__uint_as_float(orr[c + 6]) * inv, __uint_as_float(orr[c + 7]) * inv);
              *reinterpret_cast<uint4*>(orow + c0 + c) = v;
            }
          }
        }
      } else if (qrow < Lq) {
        for (int c = 0; c < D; c += 8) *reinterpret_cast<uint4*>(orow + c) = make_uint4(0, 0, 0, 0);
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 9) umma::tmem_dealloc(tb, 512);
}

template <typename T, int D>
static int prefill_umma_launch_t(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv,
                                 const PrefillUmmaParams& p, int num_seqs, int max_seqlen_q,
                                 cudaStream_t stream) {
  constexpr int CH = D / 64;
  constexpr size_t smem = 1024 + 2 * CH * kUBM * 128 + 2 * kUStages * CH * kUBN * 128 +
                          4 * kUBM * 128 + 32 * 8 + 16;
  auto kern = prefill_umma_kernel<T, D>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid((max_seqlen_q + 2 * kUBM - 1) / (2 * kUBM), p.H, num_seqs);
  launch_pdl(kern, grid, dim3(kUThreads), smem, stream, mq, mk, mv, p);
  return finish_launch();
}

// Entry used by b200_prefill_varlen_attn (prefill_attn.cu). Returns B200_ERR_UNSUPPORTED when the
// strides cannot be described by a TMA map (caller falls back to the mma.sync kernel).
int prefill_umma_launch(void* out, const void* q, const void* k, const void* v,
                        const int32_t* cu_q, const int32_t* cu_k, int num_seqs, int total_q,
                        int total_k, int max_seqlen_q, int H, int Hkv, int D, int64_t q_stride,
                        int64_t k_stride, int64_t v_stride, int64_t out_stride,
                        const int32_t* block_tables, int bt_stride, int num_blocks, int block_size,
                        float scale, int dtype, cudaStream_t stream) {
  if (q_stride % 8 != 0 || out_stride % 8 != 0 || !aligned16(out)) return B200_ERR_UNSUPPORTED;
  if (total_q <= 0) return B200_ERR_UNSUPPORTED;
  const int CH = D / 64;
  CUtensorMap mq, mk, mv;
  int rc = get_panel_map(q, total_q, q_stride * 2, H * CH, kUBM, CH, dtype, &mq);
  if (rc != B200_OK) return rc;
  PrefillUmmaParams p;
  p.cu_q = cu_q, p.cu_k = cu_k, p.block_tables = block_tables;
  p.out = out, p.out_stride = out_stride;
  p.H = H, p.G = H / Hkv, p.block_size = block_size, p.bt_stride = bt_stride;
  p.scale_log2 = scale * 1.4426950408889634f;
  if (block_tables == nullptr) {
    if (k_stride % 8 != 0 || v_stride % 8 != 0 || total_k <= 0) return B200_ERR_UNSUPPORTED;
    p.kv_rows_box = kUBN, p.kv_split = 0;
    rc = get_panel_map(k, total_k, k_stride * 2, Hkv * CH, kUBN, CH, dtype, &mk);
    if (rc != B200_OK) return rc;
    rc = get_panel_map(v, total_k, v_stride * 2, Hkv * CH, kUBN, CH, dtype, &mv);
    if (rc != B200_OK) return rc;
  } else {
    if (block_size % 16 != 0 || (block_size < kUBN && kUBN % block_size != 0) ||
        (block_size > kUBN && block_size % kUBN != 0))
      return B200_ERR_UNSUPPORTED;
    p.kv_rows_box = block_size < kUBN ? block_size : kUBN;
    p.kv_split = p.kv_rows_box < kUBN ? 1 : 0;
    const int64_t rows = (int64_t)num_blocks * block_size;
    const int cbox = p.kv_split ? 1 : CH;
    rc = get_panel_map(k, rows, (int64_t)Hkv * D * 2, Hkv * CH, p.kv_rows_box, cbox, dtype, &mk);
    if (rc != B200_OK) return rc;
    rc = get_panel_map(v, rows, (int64_t)Hkv * D * 2, Hkv * CH, p.kv_rows_box, cbox, dtype, &mv);
    if (rc != B200_OK) return rc;
  }
  if (dtype == B200_DTYPE_BF16)
    return D == 128 ? prefill_umma_launch_t<__nv_bfloat16, 128>(mq, mk, mv, p, num_seqs, max_seqlen_q, stream)
                    : prefill_umma_launch_t<__nv_bfloat16, 64>(mq, mk, mv, p, num_seqs, max_seqlen_q, stream);
  return D == 128 ? prefill_umma_launch_t<__half, 128>(mq, mk, mv, p, num_seqs, max_seqlen_q, stream)
                  : prefill_umma_launch_t<__half, 64>(mq, mk, mv, p, num_seqs, max_seqlen_q, stream);
}

}  // namespace b200
