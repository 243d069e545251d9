// moe_umma.cu -- grouped expert GEMMs for PREFILL-sized MoE batches on tcgen05 / TMEM (staged).
//
// moe.cu's expert GEMMs are weight-streaming kernels for decode (tens of rows per expert). A prefill
// step routes ~1000 rows to every expert (16384 tokens x top-8 / 128 experts): 1.2 TFLOP per layer,
// tensor-bound. Same contract as moe_gemm_kernel<gate_up/down> -- same plan (offsets/sorted_pair),
// same rounding points, same outputs -- but:
//   * tiles are 128 rows of ONE expert x 128 weight rows, accumulated in TMEM (128 fp32 columns)
//     by tcgen05.mma M128 N128 K16, both operands K-major SW128 panels (the QK^T configuration of
//     prefill_umma.cu, descriptors and instruction forms identical);
//   * A rows must be contiguous for TMA, so gate_up reads a gathered copy XS[pos] = X[token(pos)]
//     written by moe_gather_rows_kernel (HBM: 2 x pairs x hidden x 2 B per layer, ~15 % of the GEMM
//     time at 1 PFLOP/s); down reads H, which is already in plan order;
//   * gate_up's weight tile is 64 gate rows + the 64 matching up rows (two TMA boxes into one
//     panel), so SiluAndMul is the TMEM->register epilogue: thread = row, columns [0,64) = gate,
//     [64,128) = up; down's epilogue multiplies by the pair's routing weight;
//   * warp 0 = TMA producer (3- or 6-stage ring of 16 KB A + 16 KB W per 64-wide K block), warp 1 =
//     MMA issuer, warps 2-5 = epilogue (one TMEM lane quarter each); with 3 stages 97 KB smem + 128
//     TMEM columns -> two CTAs per SM, so one CTA's epilogue overlaps the other's main loop;
//   * grid = (N tiles, upper bound on row tiles = pairs/128 + experts): every CTA finds its
//     (expert, row tile) by walking the offsets; CTAs past the last tile exit at once. A row tile
//     that runs past its expert's rows computes them against the wrong weights and discards them
//     in the epilogue (never stored).
#include "umma.cuh"

namespace b200 {

struct MoeUmmaParams {
  const int32_t* offsets;      // [num_local + 1]
  const int32_t* sorted_pair;  // down: pair index per position
  const void* pair_scale;      // down: routing weight per pair
  void* out;                   // gate_up: H [pairs, n_out]; down: Y [pairs, n_out]
  int num_local, K, n_out;
};

// Ring depth: 3 stages (97 KB, two CTAs per SM) for tensor-bound grouped GEMMs, where the second CTA's
// main loop hides the first one's epilogue; 6 stages (193 KB, one CTA per SM) when the grid is small
// and every CTA is a latency-bound weight stream (the dense decode layers of ops.gate_up_silu_umma).
constexpr int kMuPanel = 128 * 128;                 // one 128-row SW128 panel (A or W), bytes
constexpr int kMuStageBytes = 2 * kMuPanel;
constexpr int kMuThreads = 192;
constexpr int mu_smem_bytes(int stages) { return 1024 + stages * kMuStageBytes + 128; }

template <typename T, bool kGateUp, int kMuStages>
__global__ void __launch_bounds__(kMuThreads, kMuStages <= 3 ? 2 : 1)
moe_umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                     const MoeUmmaParams p) {
  extern __shared__ __align__(1024) uint8_t mu_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mu_smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kMuStages * kMuStageBytes);
  uint64_t* full = bars;                       // [kMuStages]
  uint64_t* empty = bars + kMuStages;          // [kMuStages]
  uint64_t* acc_full = bars + 2 * kMuStages;   // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMuStages + 1);

  pdl_wait();
  // ---- which (expert, row tile) is mine? (uniform over the CTA)
  int e = -1, m0 = 0, row_begin = 0, cnt = 0;
  {
    const int tile = blockIdx.y;
    int t = 0;
    for (int i = 0; i < p.num_local; ++i) {
      const int b = p.offsets[i], c = p.offsets[i + 1] - b;
      const int nt = (c + 127) >> 7;
      if (tile < t + nt) {
        e = i, m0 = (tile - t) << 7, row_begin = b, cnt = c;
        break;
      }
      t += nt;
    }
  }
  if (e < 0) return;
  const int n0 = blockIdx.x * (kGateUp ? 64 : 128);
  const int KB = p.K / 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMuStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) umma::tmem_alloc(tmem_slot, 128);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmW);
      uint64_t pol_keep;
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
      // weight rows of this tile inside the [num_local * rows_per_expert, K] weight matrix
      const int w_row0 = kGateUp ? e * 2 * p.n_out + n0 : e * p.n_out + n0;
      const int w_row1 = kGateUp ? e * 2 * p.n_out + p.n_out + n0 : w_row0 + 64;
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % kMuStages;
        mbar_wait(&empty[s], ((kb / kMuStages) & 1) ^ 1);
        uint8_t* sa = smem + s * kMuStageBytes;
        uint8_t* sb = sa + kMuPanel;
        mbar_expect_tx(&full[s], kMuStageBytes);
        tma_load_3d(sa, &tmA, &full[s], 0, row_begin + m0, kb, pol_keep);
        tma_load_3d(sb, &tmW, &full[s], 0, w_row0, kb, pol_keep);
        tma_load_3d(sb + 64 * 128, &tmW, &full[s], 0, w_row1, kb, pol_keep);
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = umma::make_idesc<T>(128, 128, 0, 0);      // both operands K-major
    const uint64_t ad0 = umma::desc_kmajor(smem_u32(smem));
    const uint64_t bd0 = umma::desc_kmajor(smem_u32(smem + kMuPanel));
    const bool leader = umma::elect_one();
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % kMuStages;
      mbar_wait(&full[s], (kb / kMuStages) & 1);
      umma::fence_after_sync();
      if (leader) {
        const uint64_t ad = umma::desc_advance(ad0, s * kMuStageBytes);
        const uint64_t bd = umma::desc_advance(bd0, s * kMuStageBytes);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma::mma_ss(tb, umma::desc_advance(ad, ks * 32), umma::desc_advance(bd, ks * 32), idesc,
                       (kb > 0 || ks > 0) ? 1u : 0u);
        umma::commit(&empty[s]);                 // stage free once these MMAs have retired
        if (kb == KB - 1) umma::commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // =========================== epilogue warps (2..5) ===========================
    const int quarter = warp & 3;                    // the TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;             // row inside the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const bool valid = m0 + row < cnt;
    const int64_t pos = row_begin + m0 + (valid ? row : 0);
    T* orow = reinterpret_cast<T*>(p.out) + pos * p.n_out + n0;
    mbar_wait(acc_full, 0);
    umma::fence_after_sync();
    if (kGateUp) {
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t g[32], u[32];
        umma::ld32(tb + lane_base + c0, g);
        umma::ld32(tb + lane_base + 64 + c0, u);
        umma::wait_ld();
        if (valid) {
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // gate_up in the model dtype, silu in fp32 -> dtype, dtype multiply (activation.py:11-13)
              const float g0 = to_f32<T>(from_f32<T>(__uint_as_float(g[c + 2 * j])));
              const float g1 = to_f32<T>(from_f32<T>(__uint_as_float(g[c + 2 * j + 1])));
              const float u0 = to_f32<T>(from_f32<T>(__uint_as_float(u[c + 2 * j])));
              const float u1 = to_f32<T>(from_f32<T>(__uint_as_float(u[c + 2 * j + 1])));
              const float s0 = to_f32<T>(from_f32<T>(__fdividef(g0, 1.0f + __expf(-g0))));
              const float s1 = to_f32<T>(from_f32<T>(__fdividef(g1, 1.0f + __expf(-g1))));
              w[j] = pack2<T>(__fmul_rn(s0, u0), __fmul_rn(s1, u1));
            }
            *reinterpret_cast<uint4*>(orow + c0 + c) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
    } else {
      // pair_scale == nullptr: a plain (dense) GEMM epilogue -- used when the kernel serves an ordinary
      // linear layer as a one-expert group (ops.linear_umma)
      const float sc = !valid ? 0.f
                       : p.pair_scale == nullptr
                           ? 1.f
                           : to_f32<T>(reinterpret_cast<const T*>(p.pair_scale)[p.sorted_pair[pos]]);
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t a[32];
        umma::ld32(tb + lane_base + c0, a);
        umma::wait_ld();
        if (valid) {
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // expert output in the model dtype, then a dtype multiply with the routing weight
              const float y0 = to_f32<T>(from_f32<T>(__uint_as_float(a[c + 2 * j])));
              const float y1 = to_f32<T>(from_f32<T>(__uint_as_float(a[c + 2 * j + 1])));
              w[j] = pack2<T>(__fmul_rn(y0, sc), __fmul_rn(y1, sc));
            }
            *reinterpret_cast<uint4*>(orow + c0 + c) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tb, 128);
}

// XS[pos] = X[sorted_pair[pos] / top_k] for the positions the plan holds; one warp per row
template <typename T>
__global__ void __launch_bounds__(256)
moe_gather_rows_kernel(T* __restrict__ xs, const T* __restrict__ x, const int32_t* __restrict__ offsets,
                       const int32_t* __restrict__ sorted_pair, int num_local, int top_k, int hidden,
                       int64_t x_stride) {
  pdl_wait();
  const int total = offsets[num_local];
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int nvec = hidden / 8;
  for (int pos = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; pos < total; pos += warps) {
    const uint4* src = reinterpret_cast<const uint4*>(x + (int64_t)(sorted_pair[pos] / top_k) * x_stride);
    uint4* dst = reinterpret_cast<uint4*>(xs + (int64_t)pos * hidden);
    for (int v = lane; v < nvec; v += 32) dst[v] = __ldg(src + v);
  }
}

template <typename T, bool kGateUp, int kStages>
static int moe_umma_launch_s(const CUtensorMap& ma, const CUtensorMap& mw, const MoeUmmaParams& p,
                             dim3 grid, cudaStream_t stream) {
  auto kern = moe_umma_gemm_kernel<T, kGateUp, kStages>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         mu_smem_bytes(kStages));
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  launch_pdl(kern, grid, dim3(kMuThreads), (size_t)mu_smem_bytes(kStages), stream, ma, mw, p);
  return finish_launch();
}

template <typename T>
static int moe_umma_launch(bool gate_up, const CUtensorMap& ma, const CUtensorMap& mw,
                           const MoeUmmaParams& p, int pairs, cudaStream_t stream) {
  const int row_tiles = (pairs + 127) / 128 + p.num_local;      // upper bound; surplus CTAs exit at once
  const dim3 grid(p.n_out / (gate_up ? 64 : 128), row_tiles);
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const bool deep = (long long)grid.x * ((pairs + 127) / 128) <= 2ll * sms;   // few live CTAs: latency-bound
  if (gate_up)
    return deep ? moe_umma_launch_s<T, true, 6>(ma, mw, p, grid, stream)
                : moe_umma_launch_s<T, true, 3>(ma, mw, p, grid, stream);
  return deep ? moe_umma_launch_s<T, false, 6>(ma, mw, p, grid, stream)
              : moe_umma_launch_s<T, false, 3>(ma, mw, p, grid, stream);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_moe_gather_rows(void* xs, const void* x, const int32_t* offsets,
                                    const int32_t* sorted_pair, int32_t pairs, int32_t num_local,
                                    int32_t top_k, int32_t hidden, int64_t x_stride, int32_t dtype,
                                    b200_stream_t stream) {
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(pairs < 0 || num_local <= 0 || top_k <= 0 || hidden <= 0 || x_stride < hidden,
                 B200_ERR_BAD_SHAPE);
  if (pairs == 0) return B200_OK;
  B200_RETURN_IF(!xs || !x || !offsets || !sorted_pair, B200_ERR_NULL_POINTER);
  B200_RETURN_IF(hidden % 8 != 0 || x_stride % 8 != 0 || !aligned16(xs) || !aligned16(x), B200_ERR_MISALIGNED);
  int grid = (pairs + 7) / 8;
  if (grid > 148 * 8) grid = 148 * 8;
  if (dtype == B200_DTYPE_BF16)
    launch_pdl(moe_gather_rows_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream,
               (__nv_bfloat16*)xs, (const __nv_bfloat16*)x, offsets, sorted_pair, num_local, top_k, hidden,
               x_stride);
  else
    launch_pdl(moe_gather_rows_kernel<__half>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (__half*)xs,
               (const __half*)x, offsets, sorted_pair, num_local, top_k, hidden, x_stride);
  return finish_launch();
}

static int moe_umma_common(bool gate_up, void* out, const void* a, const void* w, const int32_t* offsets,
                           const int32_t* sorted_pair, const void* pair_scale, int32_t pairs, int32_t K,
                           int32_t n_out, int32_t num_local, int32_t dtype, b200_stream_t stream) {
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(pairs < 0 || K <= 0 || n_out <= 0 || num_local <= 0, B200_ERR_BAD_SHAPE);
  B200_RETURN_IF(K % 64 != 0 || n_out % (gate_up ? 64 : 128) != 0 || num_local > 256, B200_ERR_UNSUPPORTED);
  if (pairs == 0) return B200_OK;
  B200_RETURN_IF(!out || !a || !w || !offsets, B200_ERR_NULL_POINTER);
  B200_RETURN_IF(!gate_up && ((sorted_pair == nullptr) != (pair_scale == nullptr)), B200_ERR_NULL_POINTER);
  B200_RETURN_IF(!aligned16(out) || !aligned16(a) || !aligned16(w), B200_ERR_MISALIGNED);
  CUtensorMap ma, mw;
  int rc = get_panel_map(a, pairs, (int64_t)K * 2, K / 64, 128, 1, dtype, &ma);
  if (rc != B200_OK) return rc;
  const int64_t w_rows = (int64_t)num_local * (gate_up ? 2 : 1) * n_out;
  rc = get_panel_map(w, w_rows, (int64_t)K * 2, K / 64, 64, 1, dtype, &mw);
  if (rc != B200_OK) return rc;
  MoeUmmaParams p{offsets, sorted_pair, pair_scale, out, num_local, K, n_out};
  return dtype == B200_DTYPE_BF16 ? moe_umma_launch<__nv_bfloat16>(gate_up, ma, mw, p, pairs, (cudaStream_t)stream)
                                  : moe_umma_launch<__half>(gate_up, ma, mw, p, pairs, (cudaStream_t)stream);
}

extern "C" int b200_moe_gemm_gate_up_umma(void* h, const void* xs, const void* w13, const int32_t* offsets,
                                          int32_t pairs, int32_t hidden, int32_t inter, int32_t num_local,
                                          int32_t dtype, b200_stream_t stream) {
  return moe_umma_common(true, h, xs, w13, offsets, nullptr, nullptr, pairs, hidden, inter, num_local, dtype,
                         stream);
}

extern "C" int b200_moe_gemm_down_umma(void* y, const void* h, const void* w2, const int32_t* offsets,
                                       const int32_t* sorted_pair, const void* topk_weights, int32_t pairs,
                                       int32_t inter, int32_t hidden, int32_t num_local, int32_t dtype,
                                       b200_stream_t stream) {
  return moe_umma_common(false, y, h, w2, offsets, sorted_pair, topk_weights, pairs, inter, hidden, num_local,
                         dtype, stream);
}
