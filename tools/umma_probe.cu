// umma_probe.cu -- development probe for the tcgen05 building blocks (NOT part of the product
// library): one CTA computes  S = A * B^T  (A [128,K] K-major, B [N,K] K-major) and
// O = P * V  (P [128,64] K-major, V [64,D] MN-major) through SW128 panels, UMMA and tcgen05.ld,
// so that descriptor / instruction-descriptor / TMEM-layout assumptions can be checked against
// torch.matmul before the full attention kernel depends on them.
#include "../nano-vllm-ascend_b200/csrc/umma.cuh"

using namespace b200;
using bf16 = __nv_bfloat16;

// copy a row-major [rows][cols] bf16 matrix (cols % 64 == 0) into SW128 panels [cols/64][rows][128B]
__device__ void load_panels(uint8_t* dst, const bf16* src, int rows, int cols) {
  const int c16_per_row = cols / 8;
  for (int idx = threadIdx.x; idx < rows * c16_per_row; idx += blockDim.x) {
    const int r = idx / c16_per_row, c = idx % c16_per_row;
    const uint4 v = *reinterpret_cast<const uint4*>(src + (size_t)r * cols + c * 8);
    *reinterpret_cast<uint4*>(dst + (c >> 3) * rows * 128 + umma::sw128(r, c & 7)) = v;
  }
}

// mode 0: out[128][N] = A[128][K] * B[N][K]^T      (both K-major)
// mode 1: out[128][N] = A[128][K] * Bmn[K][N]      (B MN-major: rows = K index, N contiguous)
// mode 2 / 3: as 0 / 1 but A comes from TENSOR MEMORY (the ".ts" form: P of the PV product in an
//   FA4-shaped kernel). Hypothesis under test: a 16-bit A operand [128][K] occupies K/2 32-bit
//   columns, lane = row, column c = elements (2c, 2c+1) of that row -- written here with
//   tcgen05.st.32x32b (thread = lane, register i = column i) from packed bf16 pairs.
extern "C" __global__ void __launch_bounds__(128)
umma_probe_kernel(const bf16* A, const bf16* B, float* out, int N, int K, int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // [K/64][128][128B]
  uint8_t* sB = smem + 128 * K * 2;         // mode0: [K/64][N][128B]; mode1: [N/64][K][128B]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  load_panels(sA, A, 128, K);
  if ((mode & 1) == 0) load_panels(sB, B, N, K); else load_panels(sB, B, K, N);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) umma::tmem_alloc(&tmem_base, 256);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = tmem_base;
  const bool a_tmem = mode >= 2;
  const int bmode = mode & 1;
  const uint32_t ta = tb + 256 - K / 2;          // A operand columns (accumulator uses [0, N))
  if (a_tmem) {
    // thread = row; K/2 packed columns in chunks of 32
    const int row = warp * 32 + (threadIdx.x & 31);
    for (int c0 = 0; c0 < K / 2; c0 += 32) {
      uint32_t r[32];
      for (int i = 0; i < 32; ++i)
        r[i] = *reinterpret_cast<const uint32_t*>(A + (size_t)row * K + 2 * (c0 + i));
      umma::st32(ta + ((uint32_t)(warp * 32) << 16) + c0, r);
    }
    umma::wait_st();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  }
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma::make_idesc<bf16>(128, N, 0, bmode);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t ad = umma::desc_kmajor(smem_u32(sA) + (ks >> 2) * (128 * 128) + (ks & 3) * 32);
      uint64_t bd;
      if (bmode == 0) bd = umma::desc_kmajor(smem_u32(sB) + (ks >> 2) * (N * 128) + (ks & 3) * 32);
      else            bd = umma::desc_mnmajor(smem_u32(sB) + ks * 16 * 128, K * 128);
      if (a_tmem) umma::mma_ts(tb, ta + ks * 8, bd, idesc, ks > 0);   // 16 K-elements = 8 columns
      else        umma::mma_ss(tb, ad, bd, idesc, ks > 0);
    }
    umma::commit(&bar);
  }
  mbar_wait(&bar, 0);
  umma::fence_after_sync();
  const int row = warp * 32 + (threadIdx.x & 31);
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    umma::ld32(tb + ((uint32_t)(warp * 32) << 16) + c0, r);
    umma::wait_ld();
    for (int i = 0; i < 32; ++i) out[(size_t)row * N + c0 + i] = __uint_as_float(r[i]);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tb, 256);
}

extern "C" int umma_probe(const void* A, const void* B, float* out, int N, int K, int mode, void* stream) {
  size_t smem = 1024 + 128 * K * 2 + (size_t)N * K * 2;
  cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const bf16*)A, (const bf16*)B, out, N, K, mode);
  return (int)cudaGetLastError();
}

// ---- throughput probe: one CTA per SM issues `reps` x (ksteps MMAs) back to back and reports the
// cycles per MMA instruction for a given shape / operand majorness.
extern "C" __global__ void __launch_bounds__(128)
umma_rate_kernel(long long* cycles_out, int N, int ksteps, int b_mn_flags, int reps) {
  const int b_mn = b_mn_flags & 1;
  const bool a_tmem = (b_mn_flags & 2) != 0;     // A from tensor memory (columns 448..511; content irrelevant)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (64 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) umma::tmem_alloc(&tmem_base, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = tmem_base;
  if (warp == 1) {
    // warp-uniform loop, precomputed descriptors, one elected lane issues (as the real kernel does)
    const uint32_t idesc = umma::make_idesc<bf16>(128, N, 0, b_mn);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem) + 32 * 1024;
    uint64_t ad[8], bd[8];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      ad[ks] = umma::desc_kmajor(a_addr + (ks >> 2) * (128 * 128) + (ks & 3) * 32);
      bd[ks] = b_mn ? umma::desc_mnmajor(b_addr + (ks & 3) * 16 * 128, 64 * 128)
                    : umma::desc_kmajor(b_addr + (ks >> 2) * (N * 128) + (ks & 3) * 32);
    }
    const bool leader = umma::elect_one();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          if (ks < ksteps) {
            if (a_tmem) umma::mma_ts(tb + (r & 1) * 256 * (N <= 192), tb + 448 + ks * 8, bd[ks], idesc, ks > 0);
            else        umma::mma_ss(tb + (r & 1) * 256 * (N <= 192), ad[ks], bd[ks], idesc, ks > 0);
          }
      }
      __syncwarp();
    }
    if (leader) umma::commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    if (leader) cycles_out[blockIdx.x] = clock64() - t0;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tb, 512);
}

extern "C" int umma_rate(long long* cycles_out, int blocks, int N, int ksteps, int b_mn, int reps, void* stream) {
  size_t smem = 1024 + 96 * 1024;
  cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_rate_kernel<<<blocks, 128, smem, (cudaStream_t)stream>>>(cycles_out, N, ksteps, b_mn, reps);
  return (int)cudaGetLastError();
}
