// prefill_umma.cu -- K4/K5 on the 5th-generation tensor cores: varlen causal flash-attention with
// tcgen05.mma (UMMA), accumulators AND the probability tile in TMEM, operands staged by TMA.
//
// Replaces the same reference call sites as prefill_attn.cu (layers/attention.py:46-59,
// layers/attention_ori.py:63-70). Same semantics: bottom-right causal mask, packed K/V or paged K/V
// through block_tables.
//
// Shape of the kernel ("FA4 shape"; the r01 kernel with 64-key S tiles and P through shared memory was
// bound by shared-memory operand traffic -- measured instruction costs on this part,
// profiles/r01_umma_probe.log: SS M128 N64 K16 75 cycles, SS N128 107, TS N128 74 -- and measured
// 537 / 852 / 504 TFLOP/s on the three shapes of tools/kbench.py where this one does 559 / 926 / 602):
// Two CTA layouts, template parameter QT = 128-row query tiles per CTA, chosen per launch (prefill_query_tiles).
// QT = 2: CTA = one (sequence, head, 256 query rows[, key split]): two query tiles t = 0, 1 that ping-pong.
//   warps 0-3 / 4-7   softmax warpgroup of tile t, one thread per query row (= one TMEM lane): reads
//                     S[t] (128 fp32 columns) with tcgen05.ld, fp32 online softmax in the exp2 domain
//                     with lazy rescale of O (only when the row max grew by > 2^8), writes P as packed
//                     16-bit pairs OVER the first 64 columns of S[t] with tcgen05.st, in two halves of 64
//                     keys so that the PV product of the first half runs under the exponentials of the second
//   warp 8            TMA producers: lane 0 loads Q then runs a 2-stage K ring, lane 1 a 2-stage V ring
//                     (128-key tiles)
//   warp 9            MMA issuer (one elected lane):
//                       S[t]  = Q[t] K^T     8 x tcgen05.mma SS  (M128 N128 K16), K-major operands
//                       O[t] += P[t] V       8 x tcgen05.mma TS  (M128 N=D  K16), A = P from TMEM,
//                                            B = V MN-major straight from its [key][d] rows
//                     issue order  S0(0) S1(0) | PV0(0) S0(1) PV1(0) S1(1) | PV0(1) S0(2) ...
//                     so that the tensor pipe works on tile 1-t while warpgroup t is in its softmax.
//   TMEM (512 columns): O[t] at 128 t, S[t] at 256 + 128 t, P[t] aliases S[t] columns [0, 64).
//   Shared memory: Q 2 x 32 KiB + K 2 x 32 KiB + V 2 x 32 KiB = 192 KiB at D = 128; one CTA per SM.
// QT = 1: one query tile per CTA: warps 0-3 softmax, warp 4 TMA, warp 5 MMA; single-stage K / V rings; 256
//   TMEM columns (O at 0, S at 128), 97 KiB of shared memory -> two CTAs per SM (see the kernel's comment).
// Hazards: S[t](j+1) overwrites the columns PV[t](j) reads P from -- both are tcgen05.mma of the same
// issuing thread and execute in issue order; softmax(j+1) of tile t starts only after S[t](j+1) has
// completed, hence after PV[t](j) has.
//
// Split-KV (few, long rows: the prefix-cache-hit form q_len << k_len, BASELINE config 3): when the grid
// would leave most SMs idle, `kv_splits` CTAs share one (sequence, head, 256 rows), each takes a
// contiguous range of key tiles and writes un-normalised (O, m, l) partials to a caller-provided
// workspace; prefill_combine_kernel merges them. Flops: 4 * Lq * Lk * D per (sequence, head) before
// causal skipping; tensor-pipe bound.
#include "umma.cuh"

namespace b200 {

struct PrefillUmmaParams {
  const int32_t* cu_q;
  const int32_t* cu_k;
  const int32_t* block_tables;  // nullptr => packed K/V
  void* out;
  int64_t out_stride;
  int H, G, block_size, bt_stride;
  int kv_rows_box;   // rows per K/V TMA op (128 packed; min(block_size, 128) paged)
  int kv_split;      // 1: each TMA op carries a single 64-wide chunk (paged, rows_box < 128)
  int kv_splits;     // CTAs sharing one (sequence, head, 256 rows) along the keys; 1 = no split
  float* part_o;     // kv_splits > 1: [total_q][H][kv_splits][D] un-normalised O
  float* part_ml;    // kv_splits > 1: [total_q][H][kv_splits][2]  (row max in the exp2 domain, row sum)
  float scale_log2;
};

constexpr int kUBM = 128, kUBN = 128;

// QT = query tiles per CTA. QT = 2: the layout described above (one CTA per SM, two tiles ping-pong, K/V
// tiles shared, 2-stage rings). QT = 1: one 128-row tile per CTA, single-stage K/V rings (K(j+1) is loaded
// under softmax(j) + PV(j), V(j+1) under S(j+1) + softmax(j+1) -- the chain leaves the time), 97 KB of
// shared memory, 256 TMEM columns, 192 threads -> TWO CTAs per SM: the two chains sharing the tensor pipe
// are independent CTAs, so one CTA's prologue (barrier init, TMEM alloc, Q load) and epilogue run under
// the other's main loop and the hardware balances 128-row items.
template <typename T, int D, int QT>
__global__ void __launch_bounds__((4 * QT + 2) * 32, 3 - QT)
prefill_umma_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const PrefillUmmaParams p) {
  constexpr int kUStages = QT;                  // K/V ring depth
  constexpr int kTmaWarp = 4 * QT, kMmaWarp = 4 * QT + 1;
  constexpr uint32_t kTmemCols = QT * 256;      // O[t] at 128 t, S[t] at QT * 128 + 128 t
  constexpr uint32_t kSBase = QT * 128;
  constexpr int CH = D / 64;                    // 128-byte chunks per head row
  constexpr int QPANEL = kUBM * 128;            // bytes of one Q panel (128 rows)
  constexpr int KPANEL = kUBN * 128;            // bytes of one K / V panel (128 rows)
  constexpr int QTILE = CH * QPANEL;
  constexpr int KTILE = CH * KPANEL;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sQ = smem;                            // [QT][QTILE]
  uint8_t* sK = sQ + QT * QTILE;                  // [kUStages][KTILE]
  uint8_t* sV = sK + kUStages * KTILE;           // [kUStages][KTILE]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kUStages * KTILE);
  uint64_t* q_full = bars;                       // [1]
  uint64_t* k_full = bars + 1;                   // [kUStages]
  uint64_t* v_full = k_full + kUStages;          // [kUStages]
  uint64_t* k_empty = v_full + kUStages;         // [kUStages] K stage consumed by both tiles' S(j)
  uint64_t* v_empty = k_empty + kUStages;        // [kUStages] V stage consumed by both tiles' PV(j)
  uint64_t* s_full = v_empty + kUStages;         // [2 tiles]  S[t](j) landed in TMEM
  uint64_t* p_full = s_full + 2;                 // [2 tiles][2 halves]  128 softmax threads stored 64 keys of P[t](j)
  uint64_t* o_full = p_full + 4;                 // [2 tiles]  PV[t](j) retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  pdl_wait();
  const int seq = blockIdx.z, head = blockIdx.y;
  const int q0 = p.cu_q[seq], Lq = p.cu_q[seq + 1] - q0;
  const int k0 = p.cu_k[seq], Lk = p.cu_k[seq + 1] - k0;
  const int split = (int)blockIdx.x % p.kv_splits;
  const int m_tile = (int)gridDim.x / p.kv_splits - 1 - (int)blockIdx.x / p.kv_splits;   // heavy (late) tiles first
  const int m0 = m_tile * QT * kUBM;
  if (m0 >= Lq) return;                                       // uniform for the whole CTA
  const int kvh = head / p.G;
  const int off = Lk - Lq;
  int n_tiles[2] = {0, 0};
#pragma unroll
  for (int t = 0; t < QT; ++t) {
    const int r0 = m0 + t * kUBM;
    const int n_end = r0 < Lq ? min(Lk, r0 + kUBM + off) : 0;
    n_tiles[t] = n_end > 0 ? (n_end + kUBN - 1) / kUBN : 0;
  }
  const int ntq = (QT == 2 && m0 + kUBM < Lq) ? 2 : 1;
  // this CTA's share of the key tiles: [j_lo, j_hi) of the tiles the later query tile needs
  const int nt_all = max(n_tiles[0], n_tiles[1]);
  const int j_lo = (int)((long long)nt_all * split / p.kv_splits);
  const int j_hi = (int)((long long)nt_all * (split + 1) / p.kv_splits);
  const int NT = j_hi - j_lo;                 // loop trip count; tile index = j_lo + j
#pragma unroll
  for (int t = 0; t < QT; ++t) n_tiles[t] = max(0, min(n_tiles[t], j_hi) - j_lo);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kUStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[2 * t], kUBM);
      mbar_init(&p_full[2 * t + 1], kUBM);
      mbar_init(&o_full[t], 1);
    }
    mbar_fence_init();
  }
  if (warp == kMmaWarp) umma::tmem_alloc(tmem_slot, kTmemCols);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = *tmem_slot;

  if (warp == kTmaWarp) {
    // =========================== TMA producer ===========================
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
        const int n0 = (j_lo + j) * kUBN;
        if (p.block_tables == nullptr) {
          mbar_expect_tx(&full[s], KTILE);
          tma_load_3d(ring + s * KTILE, tm, &full[s], 0, k0 + n0, kvh * CH, pol_keep);
        } else {
          // Every op of the tile is issued: boxes past the end of the sequence use an out-of-range
          // row, which TMA zero-fills without touching memory -- P is exactly 0 there, but 0 x (stale
          // shared memory that happens to hold an Inf/NaN pattern) would poison O.
          const int rb = p.kv_rows_box;
          const int nops = kUBN / rb;
          const int per_op = p.kv_split ? CH : 1;
          const int op_bytes = rb * 128 * (p.kv_split ? 1 : CH);
          mbar_expect_tx(&full[s], nops * per_op * op_bytes);
          for (int o = 0; o < nops; ++o) {
            const int tok = n0 + o * rb;
            const int entry = tok / p.block_size;
            const int blk = tok < Lk ? p.block_tables[(int64_t)seq * p.bt_stride + entry] : -1;
            const int row = blk < 0 ? -(1 << 30) : blk * p.block_size + (tok - entry * p.block_size);
            for (int c = 0; c < per_op; ++c) {
              const int dst = (p.kv_split ? c * KPANEL : 0) + o * rb * 128;
              tma_load_3d(ring + s * KTILE + dst, tm, &full[s], 0, row, kvh * CH + c, pol_keep);
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // =========================== MMA issuer ===========================
    // Warp-uniform control flow, loop-invariant descriptor bases + compile-time offsets, one elected
    // lane issues.
    const uint32_t idesc_s = umma::make_idesc<T>(kUBM, kUBN, 0, 0);   // S = Q K^T, both K-major
    const uint32_t idesc_o = umma::make_idesc<T>(kUBM, D, 0, 1);      // O = P V, P from TMEM, V MN-major
    const uint64_t qd0 = umma::desc_kmajor(smem_u32(sQ));
    const uint64_t kd0 = umma::desc_kmajor(smem_u32(sK));
    const uint64_t vd0 = umma::desc_mnmajor(smem_u32(sV), KPANEL);
    const bool leader = umma::elect_one();
    // S[t](j): the caller has waited for k_full of stage j % kUStages
    auto issue_s = [&](int t, int j) {
      if (leader) {
        const uint64_t kd = umma::desc_advance(kd0, (j % kUStages) * KTILE);
        const uint32_t d = tb + kSBase + t * 128;
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
          const uint64_t ad = umma::desc_advance(qd0, t * QTILE + (ks >> 2) * QPANEL + (ks & 3) * 32);
          const uint64_t bd = umma::desc_advance(kd, (ks >> 2) * KPANEL + (ks & 3) * 32);
          umma::mma_ss(d, ad, bd, idesc_s, ks > 0);
        }
        umma::commit(&s_full[t]);
      }
    };
    mbar_wait(q_full, 0);
    umma::fence_after_sync();
    if (NT > 0) {
      mbar_wait(&k_full[0], 0);
      umma::fence_after_sync();
#pragma unroll
      for (int t = 0; t < QT; ++t)
        if (n_tiles[t] > 0) issue_s(t, 0);
      if (leader) umma::commit(&k_empty[0]);
      __syncwarp();
    }
    for (int j = 0; j < NT; ++j) {
      const int s = j % kUStages;
      mbar_wait(&v_full[s], (j / kUStages) & 1);
      const bool more = j + 1 < NT;
      bool k_ready = false;
      umma::fence_after_sync();
      const uint64_t vd = umma::desc_advance(vd0, s * KTILE);
#pragma unroll
      for (int t = 0; t < QT; ++t) {
        if (j < n_tiles[t]) {
          const uint32_t pa = tb + kSBase + t * 128;                  // P[t]: 16 keys = 8 packed columns
#pragma unroll
          for (int h = 0; h < 2; ++h) {                            // 64 keys at a time, as they are stored
            mbar_wait(&p_full[2 * t + h], j & 1);
            umma::fence_after_sync();
            if (leader) {
#pragma unroll
              for (int kk = h * 4; kk < h * 4 + 4; ++kk)
                umma::mma_ts(tb + t * 128, pa + kk * 8, umma::desc_advance(vd, kk * 16 * 128), idesc_o,
                             (j > 0 || kk > 0) ? 1u : 0u);
            }
          }
          if (leader) umma::commit(&o_full[t]);
          if (j + 1 < n_tiles[t]) {                                // overwrites S[t]/P[t]: after PV[t](j)
            if (!k_ready) {
              mbar_wait(&k_full[(j + 1) % kUStages], ((j + 1) / kUStages) & 1);
              umma::fence_after_sync();
              k_ready = true;
            }
            issue_s(t, j + 1);
          }
        }
      }
      if (leader) {
        umma::commit(&v_empty[s]);
        if (more) umma::commit(&k_empty[(j + 1) % kUStages]);
      }
      __syncwarp();
    }
  } else {
    // =========================== softmax warpgroups ===========================
    const int t = warp >> 2;
    const int nt = n_tiles[t];
    const int row = (warp & 3) * 32 + lane;          // row inside the 128-row tile == TMEM lane
    const int qrow = m0 + t * kUBM + row;            // row inside the sequence
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t s_addr = tb + lane_base + kSBase + t * 128;
    const uint32_t o_addr = tb + lane_base + t * 128;
    const float sl2 = p.scale_log2;
    constexpr float kLazy = 8.f;                     // lazy rescale threshold (exp2 domain): O is rescaled only when the max grew by > 2^8
    float m_ref = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nt; ++j) {
      mbar_wait(&s_full[t], j & 1);
      umma::fence_after_sync();
      uint32_t sr[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) umma::ld32(s_addr + c * 32, sr[c]);
      umma::wait_ld();
      const int n0 = (j_lo + j) * kUBN;
      // mask (warp-uniform test: does this tile touch the diagonal or the end of the keys?)
      const int warp_row0 = m0 + t * kUBM + (warp & 3) * 32;
      if (n0 + kUBN > Lk || n0 + kUBN - 1 > warp_row0 + off) {
        const int limit = min(Lk - 1, qrow + off) - n0;   // last visible key (tile-relative)
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (c > limit) sr[c >> 5][c & 31] = 0xff800000u;   // -inf
      }
      float mx = __uint_as_float(sr[0][0]);
#pragma unroll
      for (int c = 1; c < 128; ++c) mx = fmaxf(mx, __uint_as_float(sr[c >> 5][c & 31]));
      const bool grow = (mx - m_ref) * sl2 > kLazy;       // also true for the first finite max
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? mx : m_ref;
        const float alpha = (m_ref == -INFINITY) ? 0.f : ex2_approx((m_ref - m_new) * sl2);
        if (j > 0) {
          mbar_wait(&o_full[t], (j - 1) & 1);             // PV[t](j-1) landed
          umma::fence_after_sync();
#pragma unroll
          for (int c0 = 0; c0 < D; c0 += 16) {     // 16 columns at a time: the S row stays in registers
            uint32_t orr[16];
            umma::ld16(o_addr + c0, orr);
            umma::wait_ld();
#pragma unroll
            for (int c = 0; c < 16; ++c) orr[c] = __float_as_uint(__uint_as_float(orr[c]) * alpha);
            umma::st16(o_addr + c0, orr);
          }
          umma::wait_st();
        }
        l_run *= alpha;
        m_ref = m_new;
      }
      const float mb = m_ref == -INFINITY ? 0.f : m_ref * sl2;
      float sum = 0.f;
      // P(j): 128 keys -> 64 packed columns over S[t] columns [0, 64); this thread's S row is
      // already in registers, other rows live in other TMEM lanes. Packed in place: element i of
      // chunk 2h was consumed at iteration i/2.
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int c = h * 64 + 2 * i;
          const float e0 = ex2_approx(fmaf(__uint_as_float(sr[c >> 5][c & 31]), sl2, -mb));
          const float e1 = ex2_approx(fmaf(__uint_as_float(sr[(c + 1) >> 5][(c + 1) & 31]), sl2, -mb));
          sum += e0 + e1;
          sr[2 * h][i] = pack2<T>(e0, e1);
        }
        umma::st32(s_addr + h * 32, sr[2 * h]);
        umma::wait_st();
        umma::fence_before_sync();
        mbar_arrive(&p_full[2 * t + h]);          // PV of these 64 keys may start
      }
      l_run += sum;
    }
    // ---- epilogue: O / l -> global, or (split-KV) the un-normalised partial -> workspace
    if (m0 + t * kUBM < Lq) {
      T* orow = reinterpret_cast<T*>(p.out) + (int64_t)(q0 + qrow) * p.out_stride + (int64_t)head * D;
      const int64_t pidx = ((int64_t)(q0 + qrow) * p.H + head) * p.kv_splits + split;
      float* prow = p.part_o + pidx * D;
      const bool split_out = p.kv_splits > 1;
      if (split_out && qrow < Lq) {
        p.part_ml[pidx * 2] = (nt > 0 && m_ref != -INFINITY) ? m_ref * sl2 : -INFINITY;
        p.part_ml[pidx * 2 + 1] = nt > 0 ? l_run : 0.f;
      }
      if (nt > 0) {
        mbar_wait(&o_full[t], (nt - 1) & 1);
        umma::fence_after_sync();
        const float inv = split_out ? 1.f : (l_run > 0.f ? 1.f / l_run : 0.f);
#pragma unroll
        for (int c0 = 0; c0 < D; c0 += 32) {
          uint32_t orr[32];
          umma::ld32(o_addr + c0, orr);
          umma::wait_ld();
          if (qrow < Lq) {
            if (split_out) {
#pragma unroll
              for (int c = 0; c < 32; c += 4)
                *reinterpret_cast<uint4*>(prow + c0 + c) = make_uint4(orr[c], orr[c + 1], orr[c + 2], orr[c + 3]);
            } else {
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
        }
      } else if (qrow < Lq) {
        if (split_out) {
          for (int c = 0; c < D; c += 4) *reinterpret_cast<uint4*>(prow + c) = make_uint4(0, 0, 0, 0);
        } else {
          for (int c = 0; c < D; c += 8) *reinterpret_cast<uint4*>(orow + c) = make_uint4(0, 0, 0, 0);
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) umma::tmem_dealloc(tb, kTmemCols);
}

// Split-KV merge: one warp per (query row, head); lane = D/32 consecutive columns. Pieces are
// un-normalised (O, m, l) relative to their own reference max m (exp2 domain).
template <typename T, int D>
__global__ void __launch_bounds__(256)
prefill_combine_kernel(T* __restrict__ out, int64_t out_stride, const float* __restrict__ part_o,
                       const float* __restrict__ part_ml, int64_t rows_x_heads, int H, int splits) {
  constexpr int V = D / 32;
  pdl_wait();
  const int64_t item = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (item >= rows_x_heads) return;
  const int lane = threadIdx.x & 31;
  const int64_t row = item / H;
  const int head = (int)(item - row * H);
  const float* ml = part_ml + item * splits * 2;
  float m = -INFINITY;
  for (int s = 0; s < splits; ++s) m = fmaxf(m, ml[2 * s]);
  float l = 0.f, acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float ms = ml[2 * s];
    if (ms == -INFINITY) continue;
    const float sc = ex2_approx(ms - m);
    l = fmaf(ml[2 * s + 1], sc, l);
    const float* po = part_o + (item * splits + s) * D + lane * V;
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = fmaf(po[i], sc, acc[i]);
  }
  const float inv = l > 0.f ? 1.f / l : 0.f;
  T* orow = out + row * out_stride + (int64_t)head * D + lane * V;
#pragma unroll
  for (int i = 0; i < V; i += 2) *reinterpret_cast<uint32_t*>(orow + i) = pack2<T>(acc[i] * inv, acc[i + 1] * inv);
}

template <typename T, int D, int QT>
static int prefill_umma_launch_t(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv,
                                 const PrefillUmmaParams& p, int num_seqs, int max_seqlen_q, int total_q,
                                 cudaStream_t stream) {
  constexpr int CH = D / 64;
  constexpr size_t smem = 1024 + QT * CH * kUBM * 128 + 2 * QT * CH * kUBN * 128 + 32 * 8 + 16;
  auto kern = prefill_umma_kernel<T, D, QT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int m_tiles = (max_seqlen_q + QT * kUBM - 1) / (QT * kUBM);
  dim3 grid(m_tiles * p.kv_splits, p.H, num_seqs);
  launch_pdl(kern, grid, dim3((4 * QT + 2) * 32), smem, stream, mq, mk, mv, p);
  if (p.kv_splits > 1) {
    const int64_t items = (int64_t)total_q * p.H;
    launch_pdl(prefill_combine_kernel<T, D>, dim3((unsigned)((items + 7) / 8)), dim3(256), 0, stream,
               reinterpret_cast<T*>(p.out), p.out_stride, (const float*)p.part_o, (const float*)p.part_ml,
               items, p.H, p.kv_splits);
  }
  return finish_launch(p.kv_splits > 1 ? 2 : 1);
}

// Key splits for a launch: only when the (sequence, head, QT x 128-row) grid would leave more than half of
// the co-resident CTA slots idle, at least two 128-key tiles per split, at most 8, and only if the
// workspace holds them.
int prefill_kv_splits(int num_seqs, int H, int max_seqlen_q, int max_seqlen_k, int total_q, int D,
                      int64_t workspace_bytes, int qt) {
  const int slots = 148 * (3 - qt);            // co-resident CTAs: one per SM at QT = 2, two at QT = 1
  const int64_t ctas = (int64_t)((max_seqlen_q + qt * kUBM - 1) / (qt * kUBM)) * H * num_seqs;
  if (ctas * 2 > slots) return 1;
  const int key_tiles = (max_seqlen_k + kUBN - 1) / kUBN;
  int splits = (int)(slots / ctas);
  if (splits > 8) splits = 8;
  if (splits > key_tiles / 2) splits = key_tiles / 2;
  while (splits > 1 && (int64_t)total_q * H * splits * (D + 2) * 4 > workspace_bytes) --splits;
  return splits < 1 ? 1 : splits;
}

// Query tiles per CTA for a launch (measured r02, B200, TFLOP/s QT = 1 / QT = 2: 16 x 1024 D128 578 / 550;
// 4 x 4096 D128 894 / 952; 1 x 8192 D64 623 / 566; q 256 against k 8192 D64, key-split 290 / 300):
// two small CTAs per SM win while prologue / epilogue are a visible share of a CTA (short sequences) and
// on the exp2-bound D = 64 shapes; one CTA with two ping-pong tiles sharing each K/V tile wins on long
// D = 128 sequences and when the grid is so small that keys are split across CTAs anyway.
// b200_prefill_set_query_tiles(1|2) forces one layout (A/B runs, parity tests of both instantiations).
static int g_forced_query_tiles = 0;
void prefill_force_query_tiles(int qt) { g_forced_query_tiles = qt; }

int prefill_query_tiles(int num_seqs, int H, int max_seqlen_q, int D) {
  if (g_forced_query_tiles) return g_forced_query_tiles;
  const int64_t ctas2 = (int64_t)((max_seqlen_q + 2 * kUBM - 1) / (2 * kUBM)) * H * num_seqs;
  if (ctas2 * 2 <= 148) return 2;                 // key-split territory (prefill_kv_splits)
  return (D == 128 && max_seqlen_q >= 2048) ? 2 : 1;
}

// Entry used by b200_prefill_varlen_attn (prefill_attn.cu). Returns B200_ERR_UNSUPPORTED when the
// strides cannot be described by a TMA map (caller falls back to the mma.sync kernel).
int prefill_umma_launch(void* out, const void* q, const void* k, const void* v,
                         const int32_t* cu_q, const int32_t* cu_k, int num_seqs, int total_q,
                         int total_k, int max_seqlen_q, int H, int Hkv, int D, int64_t q_stride,
                         int64_t k_stride, int64_t v_stride, int64_t out_stride,
                        const int32_t* block_tables, int bt_stride, int num_blocks, int block_size,
                        float scale, int dtype, void* workspace, int64_t workspace_bytes, int max_seqlen_k,
                        cudaStream_t stream) {
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
  const int qt = prefill_query_tiles(num_seqs, H, max_seqlen_q, D);
  p.kv_splits = (workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0)
                    ? prefill_kv_splits(num_seqs, H, max_seqlen_q, max_seqlen_k, total_q, D, workspace_bytes, qt)
                    : 1;
  p.part_o = reinterpret_cast<float*>(workspace);
  p.part_ml = p.part_o + (int64_t)total_q * H * p.kv_splits * D;
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
#define B200_PREFILL_DISPATCH(T, DD)                                                                   \
  (qt == 2 ? prefill_umma_launch_t<T, DD, 2>(mq, mk, mv, p, num_seqs, max_seqlen_q, total_q, stream) \
           : prefill_umma_launch_t<T, DD, 1>(mq, mk, mv, p, num_seqs, max_seqlen_q, total_q, stream))
  if (dtype == B200_DTYPE_BF16)
    return D == 128 ? B200_PREFILL_DISPATCH(__nv_bfloat16, 128) : B200_PREFILL_DISPATCH(__nv_bfloat16, 64);
  return D == 128 ? B200_PREFILL_DISPATCH(__half, 128) : B200_PREFILL_DISPATCH(__half, 64);
#undef B200_PREFILL_DISPATCH
}

}  // namespace b200
