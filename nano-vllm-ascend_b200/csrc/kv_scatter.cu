// kv_scatter.cu -- K1/K2: scatter this step's K/V rows into the paged cache.
//
// Replaces torch_npu._npu_reshape_and_cache (layers/attention.py:25-30),
// torch_npu.scatter_update_ (layers/attention.py:32-35) and the Triton
// store_kvcache_kernel (layers/attention_ori.py:10-40) of the reference.
//
// Pure byte movement: one 16-byte vector per thread, rows of Hkv*D elements are
// contiguous on both sides, so every warp moves 512 contiguous bytes. HBM-bound
// (2 * T * Hkv * D * 2 bytes read + the same written); bit-exact by construction.
#include "common.cuh"

namespace b200 {

template <bool kSlot2D>
__global__ void __launch_bounds__(256)
kv_scatter_kernel(const uint4* __restrict__ key, const uint4* __restrict__ value,
                  uint4* __restrict__ k_cache, uint4* __restrict__ v_cache,
                  const int32_t* __restrict__ slots, int num_tokens, int vec_per_row,
                  int64_t key_stride_vec, int64_t value_stride_vec, int num_blocks,
                  int block_size) {
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
    if (which == 0) {
      k_cache[slot * vec_per_row + vec] = ld_nc_v4(key + token * key_stride_vec + vec);
    } else {
      v_cache[slot * vec_per_row + vec] = ld_nc_v4(value + token * value_stride_vec + vec);
    }
  }
}

static int scatter_launch(const void* key, const void* value, void* k_cache, void* v_cache,
                          const int32_t* slots, bool slot2d, int32_t num_tokens,
                          int32_t num_kv_heads, int32_t head_dim, int64_t key_stride,
                          int64_t value_stride, int32_t num_blocks, int32_t block_size,
                          int32_t dtype, cudaStream_t stream) {
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(num_tokens < 0 || num_kv_heads <= 0 || head_dim <= 0 || num_blocks <= 0 ||
                     block_size <= 0,
                 B200_ERR_BAD_SHAPE);
  if (num_tokens == 0) return B200_OK;
  B200_RETURN_IF(!key || !value || !k_cache || !v_cache || !slots, B200_ERR_NULL_POINTER);
  const int64_t row = (int64_t)num_kv_heads * head_dim;  // elements (2 bytes each)
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
  if (slot2d)
    launch_pdl(kv_scatter_kernel<true>, dim3(grid), dim3(256), 0, stream, (const uint4*)key,
               (const uint4*)value, (uint4*)k_cache, (uint4*)v_cache, slots, num_tokens, vec_per_row,
               key_stride / 8, value_stride / 8, num_blocks, block_size);
  else
    launch_pdl(kv_scatter_kernel<false>, dim3(grid), dim3(256), 0, stream, (const uint4*)key,
               (const uint4*)value, (uint4*)k_cache, (uint4*)v_cache, slots, num_tokens, vec_per_row,
               key_stride / 8, value_stride / 8, num_blocks, block_size);
  return finish_launch();
}

}  // namespace b200

extern "C" int b200_reshape_and_cache(const void* key, const void* value, void* k_cache,
                                      void* v_cache, const int32_t* slot_mapping,
                                      int32_t num_tokens, int32_t num_kv_heads, int32_t head_dim,
                                      int64_t key_stride, int64_t value_stride,
                                      int32_t num_blocks, int32_t block_size, int32_t dtype,
                                      b200_stream_t stream) {
  return b200::scatter_launch(key, value, k_cache, v_cache, slot_mapping, false, num_tokens,
                              num_kv_heads, head_dim, key_stride, value_stride, num_blocks,
                              block_size, dtype, (cudaStream_t)stream);
}

extern "C" int b200_reshape_and_cache_2d(const void* key, const void* value, void* k_cache,
                                         void* v_cache, const int32_t* slot_pairs,
                                         int32_t num_tokens, int32_t num_kv_heads,
                                         int32_t head_dim, int64_t key_stride,
                                         int64_t value_stride, int32_t num_blocks,
                                         int32_t block_size, int32_t dtype,
                                         b200_stream_t stream) {
  return b200::scatter_launch(key, value, k_cache, v_cache, slot_pairs, true, num_tokens,
                              num_kv_heads, head_dim, key_stride, value_stride, num_blocks,
                              block_size, dtype, (cudaStream_t)stream);
}
