// allreduce.cu -- C1 fused: row-parallel all-reduce + residual add + RMSNorm in ONE kernel over
// NVLink peer memory (one-shot), for the small messages of a tensor-parallel decode step.
//
// Replaces, for decode-sized messages, the pair  dist.all_reduce(y)  (layers/linear.py:151-152 of the
// reference: HCCL there, NCCL here) + add_rms_forward (layers/layernorm.py:27-38) that follows every
// row-parallel projection (models/qwen3.py:89,118,155-159). Prefill-sized messages keep NCCL.
//
// Each rank's GEMM writes its partial product into a buffer the other ranks can address (CUDA IPC,
// mapped by the Python side). The kernel: (1) block 0 publishes "my partial #epoch is written" into
// every peer's flag array, every block waits until all peers have published; (2) each CTA owns one
// row: 16-byte loads of that row from every rank's buffer (own HBM + P2P over NVLink), fp32 sum in
// rank order (bitwise identical on every rank), + residual, RMSNorm with the reference's rounding
// points, written locally; (3) the block that finishes last advances the epoch. Two data buffers
// alternate between consecutive calls, which makes a trailing barrier unnecessary (a rank can only
// pass the start barrier of call e+1 after every rank has finished reading call e).
// Bytes over NVLink per rank: (world-1) * T * hidden * 2; latency-bound at decode sizes.
#include "common.cuh"

#include <cstring>

namespace b200 {

constexpr int kMaxWorld = 8;

struct ArParams {
  const void* bufs[kMaxWorld];   // rank r's partial [rows, hidden] as seen from this rank
  uint32_t* flags[kMaxWorld];    // rank r's flag array [kMaxWorld] as seen from this rank
  uint32_t* state;               // local: [0] epoch, [1] finished-block counter, [2] start ticket
  int world, rank;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256)
allreduce_add_rmsnorm_kernel(const ArParams p, T* __restrict__ out, T* __restrict__ residual,
                             const T* __restrict__ weight, float eps, int hidden) {
  __shared__ float red[32];
  pdl_wait();
  __shared__ uint32_t s_ticket;
  const uint32_t epoch = p.state[0] + 1;           // read before anyone can advance it (see end)
  if (threadIdx.x == 0) s_ticket = atomicAdd(&p.state[2], 1u);
  __syncthreads();
  if (s_ticket == 0 && threadIdx.x < p.world) {    // whichever block starts first publishes
    __threadfence_system();                        // the GEMM's partial is visible system-wide
    st_release_sys(p.flags[threadIdx.x] + p.rank, epoch);
  }
  if (threadIdx.x < p.world) {
    const uint32_t* mine = p.flags[p.rank] + threadIdx.x;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
    }
  }
  __syncthreads();

  const int64_t row = blockIdx.x;
  const int nvec = hidden / 8;
  uint4* rrow = reinterpret_cast<uint4*>(residual + row * hidden);
  constexpr int kVec = 4;
  float vals[kVec][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int v = threadIdx.x + i * blockDim.x;
    if (v < nvec) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      uint4 part[kMaxWorld];
#pragma unroll
      for (int r = 0; r < kMaxWorld; ++r)
        if (r < p.world) part[r] = __ldcg(reinterpret_cast<const uint4*>(p.bufs[r]) + row * nvec + v);
#pragma unroll
      for (int r = 0; r < kMaxWorld; ++r) {
        if (r < p.world) {
          const uint32_t w[4] = {part[r].x, part[r].y, part[r].z, part[r].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack2<T>(w[j]);
            acc[2 * j] += f.x;
            acc[2 * j + 1] += f.y;
          }
        }
      }
      const uint4 rv = rrow[v];
      const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // the all-reduce result is materialised in the model dtype by the reference (NCCL/HCCL sum
        // of bf16 tensors), then add_rms_forward adds the residual in fp32
        const float2 g = unpack2<T>(rw[j]);
        const float y0 = to_f32<T>(from_f32<T>(acc[2 * j])), y1 = to_f32<T>(from_f32<T>(acc[2 * j + 1]));
        const float s0 = __fadd_rn(y0, g.x), s1 = __fadd_rn(y1, g.y);
        ow[j] = pack2<T>(s0, s1);
        vals[i][2 * j] = s0;
        vals[i][2 * j + 1] = s1;
        ss = fmaf(s0, s0, ss);
        ss = fmaf(s1, s1, ss);
      }
      rrow[v] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
  }
  // block-wide sum of squares
  ss = warp_sum(ss);
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (nw > 1) {
      if (lane == 0) red[warp] = ss;
      __syncthreads();
      float t = lane < nw ? red[lane] : 0.f;
      ss = warp_sum(t);
    }
  }
  const float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fdiv_rn(ss, (float)hidden), eps)));
  uint4* orow = reinterpret_cast<uint4*>(out + row * hidden);
  const uint4* wrow = reinterpret_cast<const uint4*>(weight);
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int v = threadIdx.x + i * blockDim.x;
    if (v < nvec) {
      const uint4 w = wrow[v];
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 wf = unpack2<T>(ww[j]);
        const float a = to_f32<T>(from_f32<T>(__fmul_rn(vals[i][2 * j], inv)));
        const float b = to_f32<T>(from_f32<T>(__fmul_rn(vals[i][2 * j + 1], inv)));
        ow[j] = pack2<T>(__fmul_rn(a, wf.x), __fmul_rn(b, wf.y));
      }
      orow[v] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
  }
  // the block that finishes last advances the epoch (every block has read it by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t done = atomicAdd(&p.state[1], 1u);
    if (done == gridDim.x - 1) {
      p.state[1] = 0;
      p.state[2] = 0;
      __threadfence();
      p.state[0] = epoch;
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_allreduce_add_rmsnorm(void* out, void* residual, const void* weight, float eps,
                                          int32_t rows, int32_t hidden, const void* const* peer_bufs,
                                          void* const* peer_flags, void* state, int32_t world,
                                          int32_t rank, int32_t dtype, b200_stream_t stream) {
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(rows < 0 || hidden <= 0 || world < 2 || world > kMaxWorld || rank < 0 || rank >= world,
                 B200_ERR_BAD_SHAPE);
  if (rows == 0) return B200_OK;
  B200_RETURN_IF(!out || !residual || !weight || !peer_bufs || !peer_flags || !state,
                 B200_ERR_NULL_POINTER);
  B200_RETURN_IF(hidden % 8 != 0 || !aligned16(out) || !aligned16(residual) || !aligned16(weight),
                 B200_ERR_MISALIGNED);
  B200_RETURN_IF(hidden > 8 * 256 * 4, B200_ERR_UNSUPPORTED);
  ArParams p;
  for (int r = 0; r < world; ++r) {
    B200_RETURN_IF(!peer_bufs[r] || !peer_flags[r], B200_ERR_NULL_POINTER);
    B200_RETURN_IF(!aligned16(peer_bufs[r]), B200_ERR_MISALIGNED);
    p.bufs[r] = peer_bufs[r];
    p.flags[r] = reinterpret_cast<uint32_t*>(peer_flags[r]);
  }
  p.state = reinterpret_cast<uint32_t*>(state);
  p.world = world;
  p.rank = rank;
  const int nvec = hidden / 8;
  int threads = (nvec + 3) / 4;
  threads = ((threads + 31) / 32) * 32;
  if (threads < 128 && nvec >= 128) threads = 128;
  if (threads > 256) threads = 256;
  if (threads < 32) threads = 32;
  if (dtype == B200_DTYPE_BF16)
    launch_pdl(allreduce_add_rmsnorm_kernel<__nv_bfloat16>, dim3(rows), dim3(threads), 0,
               (cudaStream_t)stream, p, (__nv_bfloat16*)out, (__nv_bfloat16*)residual,
               (const __nv_bfloat16*)weight, eps, hidden);
  else
    launch_pdl(allreduce_add_rmsnorm_kernel<__half>, dim3(rows), dim3(threads), 0, (cudaStream_t)stream,
               p, (__half*)out, (__half*)residual, (const __half*)weight, eps, hidden);
  return finish_launch();
}

// ---- set-up time helpers (NOT on the hot path: these allocate / map memory once per process)
extern "C" int b200_comm_alloc(size_t bytes, void** ptr) {
  B200_RETURN_IF(!ptr || bytes == 0, B200_ERR_NULL_POINTER);
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(*ptr, 0, bytes);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaDeviceSynchronize();
}
extern "C" int b200_comm_free(void* ptr) { return (int)cudaFree(ptr); }
extern "C" int b200_comm_get_handle(const void* ptr, void* handle64) {
  B200_RETURN_IF(!ptr || !handle64, B200_ERR_NULL_POINTER);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  return (int)cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), const_cast<void*>(ptr));
}
extern "C" int b200_comm_open_handle(const void* handle64, void** ptr) {
  B200_RETURN_IF(!ptr || !handle64, B200_ERR_NULL_POINTER);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  return (int)cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}
extern "C" int b200_comm_close_handle(void* ptr) { return (int)cudaIpcCloseMemHandle(ptr); }
