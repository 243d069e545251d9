// prefill_attn.cu -- K4/K5: varlen causal flash-attention for the prefill step.
//
// Replaces npu_fused_infer_attention_score_v2 TND sparse_mode=3 (layers/attention.py:46-59),
// flash_attn_varlen_func(causal=True[, block_table]) (layers/attention_ori.py:63-70) and the
// oracle Attention._prefill_attn (layers/attention_torch_native.py:103-145).
//
// This file: the mma.sync variant (kept for strides TMA cannot describe and for A/B checks,
// B200_PREFILL_KERNEL=mma_sync); the default is the tcgen05/TMEM kernel in prefill_umma.cu.
// 64x64 tiles, 4 warps x 16 query rows, K/V tiles double-buffered with
// cp.async into 128-byte-swizzled shared memory (same panel layout the TMA path uses), QK^T
// and PV on mma.sync m16n8k16 with fp32 accumulation, fp32 online softmax in the exp2 domain,
// bottom-right-aligned causal mask, K/V either packed [total_k, Hkv, D] or gathered from the
// paged cache through block_tables (prefix-cache hit). Tensor-bound: 4*Lq*Lk*D flops per
// (sequence, head) before causal halving.
#include "common.cuh"

#include <cstdlib>
#include <cstring>

namespace b200 {

struct PrefillParams {
  const void* q;
  const void* k;
  const void* v;
  void* out;
  const int32_t* cu_q;
  const int32_t* cu_k;
  const int32_t* block_tables;  // nullptr => packed K/V
  int64_t q_stride, k_stride, v_stride, out_stride;
  int H, Hkv, G, block_size, bt_stride, num_blocks;
  float scale_log2;
};

constexpr int kBM = 64, kBN = 64, kPrefillThreads = 128;

// swizzled byte offset of 16-byte chunk `col16` of row `row` in a [D/64][64][128 B] tile
__device__ __forceinline__ int tile_off(int row, int col16) {
  return (col16 >> 3) * (64 * 128) + row * 128 + (((col16 & 7) ^ (row & 7)) << 4);
}

template <typename T, int D>
__global__ void __launch_bounds__(kPrefillThreads)
prefill_attn_kernel(const PrefillParams p) {
  constexpr int KSTEPS = D / 16;
  constexpr int NT = D / 8;
  constexpr int ROW16 = D / 8;            // 16-byte chunks per row
  constexpr int TILE_BYTES = 64 * D * 2;  // one 64-row tile
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sq = smem;
  uint8_t* sk = smem + TILE_BYTES;       // [2][TILE_BYTES]
  uint8_t* sv = smem + 3 * TILE_BYTES;   // [2][TILE_BYTES]
  pdl_wait();

  const int seq = blockIdx.z, head = blockIdx.y;
  const int q0 = p.cu_q[seq], Lq = p.cu_q[seq + 1] - q0;
  const int k0 = p.cu_k[seq], Lk = p.cu_k[seq + 1] - k0;
  // heavier (later) query tiles first: better tail behaviour under causal masking
  const int m_tile = (int)gridDim.x - 1 - (int)blockIdx.x;
  const int m0 = m_tile * kBM;
  if (m0 >= Lq) return;
  const int kvh = head / p.G;
  const int off = Lk - Lq;  // bottom-right alignment: row r sees keys <= r + off
  const int n_end = min(Lk, m0 + kBM + off);
  const int n_tiles = n_end > 0 ? (n_end + kBN - 1) / kBN : 0;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const T* qg = reinterpret_cast<const T*>(p.q);
  const T* kg = reinterpret_cast<const T*>(p.k);
  const T* vg = reinterpret_cast<const T*>(p.v);

  // ---- async loads
  auto load_q = [&]() {
    for (int idx = tid; idx < 64 * ROW16; idx += kPrefillThreads) {
      const int row = idx / ROW16, c16 = idx - row * ROW16;
      const bool ok = m0 + row < Lq;
      const T* src = qg + (int64_t)(q0 + (ok ? m0 + row : 0)) * p.q_stride + (int64_t)head * D + c16 * 8;
      cp_async_16(smem_u32(sq + tile_off(row, c16)), src, ok);
    }
  };
  auto load_kv = [&](int nt, int buf) {
    const int n0 = nt * kBN;
    for (int idx = tid; idx < 64 * ROW16; idx += kPrefillThreads) {
      const int row = idx / ROW16, c16 = idx - row * ROW16;
      const int n = n0 + row;
      const bool ok = n < Lk;
      int64_t krow, vrow;
      if (p.block_tables) {
        const int nn = ok ? n : 0;
        const int entry = nn / p.block_size;
        int blk = p.block_tables[(int64_t)seq * p.bt_stride + entry];
        if (blk < 0 || blk >= p.num_blocks) blk = 0;
        const int64_t slot = (int64_t)blk * p.block_size + (nn - entry * p.block_size);
        krow = vrow = slot * p.Hkv * D;
      } else {
        krow = (int64_t)(k0 + (ok ? n : 0)) * p.k_stride;
        vrow = (int64_t)(k0 + (ok ? n : 0)) * p.v_stride;
      }
      const int64_t col = (int64_t)kvh * D + c16 * 8;
      cp_async_16(smem_u32(sk + buf * TILE_BYTES + tile_off(row, c16)), kg + krow + col, ok);
      cp_async_16(smem_u32(sv + buf * TILE_BYTES + tile_off(row, c16)), vg + vrow + col, ok);
    }
  };

  load_q();
  if (n_tiles > 0) load_kv(0, 0);
  cp_async_commit();

  float o_acc[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) o_acc[n][0] = o_acc[n][1] = o_acc[n][2] = o_acc[n][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const float sl2 = p.scale_log2;
  uint32_t qa[KSTEPS][4];

  // per-lane ldmatrix coordinates
  const int a_row = warp * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);  // Q rows (A operand)
  const int a_half = lane >> 4;
  const int k_tok = ((lane >> 4) << 3) + (lane & 7);                    // K rows (B operand)
  const int k_half = (lane >> 3) & 1;
  const int v_tok = (((lane >> 3) & 1) << 3) + (lane & 7);              // V rows (B, transposed)
  const int v_dt = lane >> 4;

  for (int nt = 0; nt < n_tiles; ++nt) {
    const int buf = nt & 1;
    if (nt + 1 < n_tiles) load_kv(nt + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (nt == 0) {
      const uint32_t qb = smem_u32(sq);
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks)
        ldsm_x4(qb + tile_off(a_row, ks * 2 + a_half), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    }
    const uint32_t kb = smem_u32(sk + buf * TILE_BYTES);
    const uint32_t vb = smem_u32(sv + buf * TILE_BYTES);

    float s[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t r0, r1, r2, r3;
        ldsm_x4(kb + tile_off(np * 16 + k_tok, ks * 2 + k_half), r0, r1, r2, r3);
        mma_16816<T>(s[2 * np], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], r0, r1);
        mma_16816<T>(s[2 * np + 1], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], r2, r3);
      }
    }
    // ---- causal / length mask (only tiles that touch the diagonal or the end)
    const int n0 = nt * kBN;
    const int row_a = m0 + warp * 16 + g;  // rows row_a and row_a + 8
    if (n0 + kBN > Lk || n0 + kBN - 1 > m0 + warp * 16 + off) {
#pragma unroll
      for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = n0 + n * 8 + 2 * t4 + (e & 1);
          const int row = row_a + ((e >> 1) << 3);
          if (key >= Lk || key > row + off) s[n][e] = -INFINITY;
        }
    }
    // ---- online softmax (rows g and g+8)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float mx = -INFINITY;
#pragma unroll
      for (int n = 0; n < 8; ++n) mx = fmaxf(mx, fmaxf(s[n][2 * r], s[n][2 * r + 1]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run[r], mx);
      const float mb = m_new == -INFINITY ? 0.f : m_new * sl2;
      const float alpha = m_new == -INFINITY ? 1.f : ex2_approx((m_run[r] - m_new) * sl2);
      float sum = 0.f;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        s[n][2 * r] = ex2_approx(fmaf(s[n][2 * r], sl2, -mb));
        s[n][2 * r + 1] = ex2_approx(fmaf(s[n][2 * r + 1], sl2, -mb));
        sum += s[n][2 * r] + s[n][2 * r + 1];
      }
      l_run[r] = fmaf(l_run[r], alpha, sum);
      m_run[r] = m_new;
#pragma unroll
      for (int n = 0; n < NT; ++n) o_acc[n][2 * r] *= alpha, o_acc[n][2 * r + 1] *= alpha;
    }
    // ---- O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t pa0 = pack2<T>(s[2 * kk][0], s[2 * kk][1]);
      const uint32_t pa1 = pack2<T>(s[2 * kk][2], s[2 * kk][3]);
      const uint32_t pa2 = pack2<T>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      const uint32_t pa3 = pack2<T>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int jp = 0; jp < NT / 2; ++jp) {
        uint32_t r0, r1, r2, r3;
        ldsm_x4_trans(vb + tile_off(kk * 16 + v_tok, 2 * jp + v_dt), r0, r1, r2, r3);
        mma_16816<T>(o_acc[2 * jp], pa0, pa1, pa2, pa3, r0, r1);
        mma_16816<T>(o_acc[2 * jp + 1], pa0, pa1, pa2, pa3, r2, r3);
      }
    }
    __syncthreads();  // everyone done with `buf` before it is refilled two iterations later
  }
  cp_async_wait<0>();

  // ---- epilogue
  T* og = reinterpret_cast<T*>(p.out);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float l = l_run[r];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    const int row = m0 + warp * 16 + g + r * 8;
    if (row < Lq) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      T* orow = og + (int64_t)(q0 + row) * p.out_stride + (int64_t)head * D;
#pragma unroll
      for (int n = 0; n < NT; ++n)
        *reinterpret_cast<uint32_t*>(orow + n * 8 + 2 * t4) =
            pack2<T>(o_acc[n][2 * r] * inv, o_acc[n][2 * r + 1] * inv);
    }
  }
}

template <typename T, int D>
static int prefill_launch(const PrefillParams& p, int num_seqs, int max_seqlen_q,
                          cudaStream_t stream) {
  constexpr size_t smem = 1024 + 5 * 64 * D * 2;
  auto kern = prefill_attn_kernel<T, D>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid((max_seqlen_q + kBM - 1) / kBM, p.H, num_seqs);
  launch_pdl(kern, grid, dim3(kPrefillThreads), smem, stream, p);
  return finish_launch();
}

int prefill_umma_launch(void* out, const void* q, const void* k, const void* v, const int32_t* cu_q,
                        const int32_t* cu_k, int num_seqs, int total_q, int total_k,
                        int max_seqlen_q, int H, int Hkv, int D, int64_t q_stride, int64_t k_stride,
                        int64_t v_stride, int64_t out_stride, const int32_t* block_tables,
                        int bt_stride, int num_blocks, int block_size, float scale, int dtype,
                        void* workspace, int64_t workspace_bytes, int max_seqlen_k,
                        cudaStream_t stream);  // prefill_umma.cu
void prefill_force_query_tiles(int qt);        // prefill_umma.cu
// B200_PREFILL_KERNEL=mma_sync forces the kernel in this file (A/B checks); default = prefill_umma.cu.
static int prefill_kernel_choice() {
  static int v = [] {
    const char* e = getenv("B200_PREFILL_KERNEL");
    return (e != nullptr && strcmp(e, "mma_sync") == 0) ? 0 : 1;
  }();
  return v;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_prefill_set_query_tiles(int32_t query_tiles_per_cta) {
  B200_RETURN_IF(query_tiles_per_cta < 0 || query_tiles_per_cta > 2, B200_ERR_BAD_SHAPE);
  prefill_force_query_tiles(query_tiles_per_cta);
  return B200_OK;
}

extern "C" int64_t b200_prefill_workspace_size(int32_t head_dim) {
  // 148 CTAs x 256 rows x (head_dim fp32 of O + m + l): the most a split launch can write
  return (int64_t)148 * 256 * (head_dim + 2) * 4;
}

extern "C" int b200_prefill_varlen_attn(void* out, const void* q, const void* k, const void* v,
                                        const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k,
                                        int32_t num_seqs, int32_t total_q, int32_t total_k,
                                        int32_t max_seqlen_q, int32_t max_seqlen_k,
                                        int32_t num_heads,
                                        int32_t num_kv_heads, int32_t head_dim, int64_t q_stride,
                                        int64_t k_stride, int64_t v_stride, int64_t out_stride,
                                        const int32_t* block_tables, int32_t block_table_stride,
                                        int32_t num_blocks, int32_t block_size, float scale,
                                        int32_t dtype, void* workspace, int64_t workspace_bytes,
                                        b200_stream_t stream) {
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(num_seqs < 0 || max_seqlen_q < 0 || max_seqlen_k < 0 || num_heads <= 0 ||
                     num_kv_heads <= 0,
                 B200_ERR_BAD_SHAPE);
  B200_RETURN_IF(num_heads % num_kv_heads != 0, B200_ERR_BAD_SHAPE);
  B200_RETURN_IF(head_dim != 64 && head_dim != 128, B200_ERR_UNSUPPORTED);
  if (num_seqs == 0 || max_seqlen_q == 0) return B200_OK;
  B200_RETURN_IF(!out || !q || !k || !v || !cu_seqlens_q || !cu_seqlens_k, B200_ERR_NULL_POINTER);
  B200_RETURN_IF(num_seqs > 65535 || num_heads > 65535, B200_ERR_UNSUPPORTED);
  if (block_tables) {
    B200_RETURN_IF(block_table_stride <= 0 || num_blocks <= 0 || block_size <= 0, B200_ERR_BAD_SHAPE);
  } else {
    B200_RETURN_IF(k_stride < (int64_t)num_kv_heads * head_dim ||
                       v_stride < (int64_t)num_kv_heads * head_dim,
                   B200_ERR_BAD_SHAPE);
    B200_RETURN_IF(k_stride % 8 != 0 || v_stride % 8 != 0, B200_ERR_MISALIGNED);
  }
  B200_RETURN_IF(q_stride < (int64_t)num_heads * head_dim || out_stride < (int64_t)num_heads * head_dim,
                 B200_ERR_BAD_SHAPE);
  B200_RETURN_IF(q_stride % 8 != 0 || out_stride % 2 != 0 || !aligned16(q) || !aligned16(k) ||
                     !aligned16(v) || (reinterpret_cast<uintptr_t>(out) & 3),
                 B200_ERR_MISALIGNED);
  cudaStream_t s = (cudaStream_t)stream;
  const int choice = prefill_kernel_choice();
  if (choice != 0) {
    const int rc = prefill_umma_launch(out, q, k, v, cu_seqlens_q, cu_seqlens_k, num_seqs, total_q,
        total_k, max_seqlen_q, num_heads, num_kv_heads, head_dim, q_stride, k_stride, v_stride,
        out_stride, block_tables, block_table_stride, num_blocks, block_size, scale, dtype,
        workspace, workspace_bytes, max_seqlen_k, s);
    if (rc != B200_ERR_UNSUPPORTED) return rc;
  }
  PrefillParams p;
  p.q = q, p.k = k, p.v = v, p.out = out;
  p.cu_q = cu_seqlens_q, p.cu_k = cu_seqlens_k;
  p.block_tables = block_tables;
  p.q_stride = q_stride, p.k_stride = k_stride, p.v_stride = v_stride, p.out_stride = out_stride;
  p.H = num_heads, p.Hkv = num_kv_heads, p.G = num_heads / num_kv_heads;
  p.block_size = block_size, p.bt_stride = block_table_stride, p.num_blocks = num_blocks;
  p.scale_log2 = scale * 1.4426950408889634f;
  if (dtype == B200_DTYPE_BF16)
    return head_dim == 128 ? prefill_launch<__nv_bfloat16, 128>(p, num_seqs, max_seqlen_q, s)
                           : prefill_launch<__nv_bfloat16, 64>(p, num_seqs, max_seqlen_q, s);
  return head_dim == 128 ? prefill_launch<__half, 128>(p, num_seqs, max_seqlen_q, s)
                         : prefill_launch<__half, 64>(p, num_seqs, max_seqlen_q, s);
}
