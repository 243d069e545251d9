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
//
// Two-shot variant (allreduce_add_rmsnorm_twoshot_kernel) for messages where pulling world-1 full
// copies is bandwidth-bound (Qwen3-32B TP=8: 18 MB in per call): rank r reduces only its slice of
// the rows (reads (world-1)/world of one message) and PUSHES the bf16 sums into every rank's gather
// buffer (writes the same amount); after a second flag barrier every rank runs the residual-add +
// RMSNorm over all rows from its local gather buffer. NVLink bytes per rank: 2*(world-1)/world *
// T*hidden*2 instead of (world-1) * T*hidden*2 -- 4.6 MB instead of 18 MB in that configuration.
//
// Multicast variant (allreduce_add_rmsnorm_nvls_kernel): the same two phases, but the buffers are
// bound to an NVSwitch multicast object (NVLS). Phase 1 becomes ONE multimem.ld_reduce per 16 bytes
// (the switch adds the world copies, fp32 accumulation) and ONE multimem.st (the switch replicates
// the sum into every rank's gather buffer): 2/world of a message crosses this rank's links instead
// of 2*(world-1)/world, and no SM spends issue slots on world-way loads and stores.
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

// fp32 sum, in rank order, of the 8 elements each rank contributed (bitwise identical wherever it
// is evaluated)
template <typename T>
__device__ __forceinline__ void sum_parts(const uint4 (&part)[kMaxWorld], int world, float (&acc)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
  for (int r = 0; r < kMaxWorld; ++r) {
    if (r < world) {
      const uint32_t w[4] = {part[r].x, part[r].y, part[r].z, part[r].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack2<T>(w[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
  }
}

// One row of  residual <- cast(y) + residual ; out <- RMSNorm(that sum) * weight  with the rounding
// points of the reference (layers/layernorm.py:27-38): y materialised in T, add in fp32, norm in fp32.
// load_sum(idx, acc) yields the 8 summed elements of 16-byte vector idx = row * hidden/8 + v.
// blockDim.x * 4 vectors cover a row (hidden <= 8192). `red`: 32 floats of shared memory.
struct PlainStore {
  __device__ __forceinline__ void operator()(uint4* dst, const uint4& v) const { *dst = v; }
};

template <typename T, typename LoadSum, typename StoreOut = PlainStore>
__device__ __forceinline__ void add_rmsnorm_row(int64_t row, int hidden, T* __restrict__ out,
                                                T* __restrict__ residual, const T* __restrict__ weight,
                                                float eps, float* red, LoadSum load_sum,
                                                StoreOut store_out = StoreOut()) {
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
      load_sum(row * nvec + v, acc);
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
      store_out(orow + v, make_uint4(ow[0], ow[1], ow[2], ow[3]));
    }
  }
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

  add_rmsnorm_row<T>(blockIdx.x, hidden, out, residual, weight, eps, red, [&](int64_t idx, float (&acc)[8]) {
    uint4 part[kMaxWorld];
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)
      if (r < p.world) part[r] = __ldcg(reinterpret_cast<const uint4*>(p.bufs[r]) + idx);
    sum_parts<T>(part, p.world, acc);
  });
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

// ---------------------------------------------------------------------------------------------
// Two-shot variant. state: [0] epoch, [1] finished-block counter, [2] start ticket, [3] blocks done
// with the reduce-scatter phase. flags: uint32[2*kMaxWorld] per rank -- [0..8) start barrier (shared
// with the one-shot kernel, same epoch sequence), [8..16) "slice pushed" barrier.
// Every CTA spins on flags, so the whole grid must be co-resident: the host caps it at one CTA per SM.
struct Ar2Params {
  ArParams base;
  void* gather[kMaxWorld];       // rank r's gather buffer [rows, hidden] as seen from this rank
  int rows;
};

template <typename T>
__global__ void __launch_bounds__(256)
allreduce_add_rmsnorm_twoshot_kernel(const Ar2Params q, T* __restrict__ out, T* __restrict__ residual,
                                     const T* __restrict__ weight, float eps, int hidden) {
  __shared__ float red[32];
  __shared__ uint32_t s_ticket;
  const ArParams& p = q.base;
  pdl_wait();
  const uint32_t epoch = p.state[0] + 1;
  if (threadIdx.x == 0) s_ticket = atomicAdd(&p.state[2], 1u);
  __syncthreads();
  if (s_ticket == 0 && threadIdx.x < p.world) {
    __threadfence_system();
    st_release_sys(p.flags[threadIdx.x] + p.rank, epoch);
  }
  if (threadIdx.x < p.world) {
    const uint32_t* mine = p.flags[p.rank] + threadIdx.x;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
    }
  }
  __syncthreads();

  // ---- phase 1: reduce my slice of the rows, push the sums (in T) to every rank's gather buffer
  const int nvec = hidden / 8;
  const int slice = (q.rows + p.world - 1) / p.world;
  const int row0 = min(p.rank * slice, q.rows), row1 = min(row0 + slice, q.rows);
  const int64_t v_begin = (int64_t)row0 * nvec, v_end = (int64_t)row1 * nvec;
  for (int64_t idx = v_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < v_end;
       idx += (int64_t)gridDim.x * blockDim.x) {
    uint4 part[kMaxWorld];
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)
      if (r < p.world) part[r] = __ldcg(reinterpret_cast<const uint4*>(p.bufs[r]) + idx);
    float acc[8];
    sum_parts<T>(part, p.world, acc);
    const uint4 y = make_uint4(pack2<T>(acc[0], acc[1]), pack2<T>(acc[2], acc[3]),
                               pack2<T>(acc[4], acc[5]), pack2<T>(acc[6], acc[7]));
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)
      if (r < p.world) reinterpret_cast<uint4*>(q.gather[r])[idx] = y;
  }
  // ---- barrier 2: the last CTA of this rank to finish its pushes tells every peer
  // One fence per CTA, not per thread: bar.sync orders every thread's pushes before thread 0's
  // gpu-scope fence + ticket; the CTA that draws the last ticket issues the system-scope fence and
  // the release stores (fences are cumulative in the PTX memory model). 37 k threads each issuing
  // fence.sys cost several microseconds per call.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_ticket = atomicAdd(&p.state[3], 1u);
  }
  __syncthreads();
  if (s_ticket == gridDim.x - 1 && threadIdx.x < p.world) {
    __threadfence_system();
    st_release_sys(p.flags[threadIdx.x] + kMaxWorld + p.rank, epoch);
  }
  if (threadIdx.x < p.world) {
    const uint32_t* mine = p.flags[p.rank] + kMaxWorld + threadIdx.x;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
    }
  }
  __syncthreads();

  // ---- phase 2: residual add + RMSNorm over all rows from the local gather buffer
  const uint4* gathered = reinterpret_cast<const uint4*>(q.gather[p.rank]);
  for (int row = blockIdx.x; row < q.rows; row += gridDim.x) {
    add_rmsnorm_row<T>(row, hidden, out, residual, weight, eps, red, [&](int64_t idx, float (&acc)[8]) {
      const uint4 y = __ldcg(gathered + idx);
      const uint32_t w[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack2<T>(w[j]);
        acc[2 * j] = f.x;
        acc[2 * j + 1] = f.y;
      }
    });
    __syncthreads();                 // `red` is reused by the next row
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t done = atomicAdd(&p.state[1], 1u);
    if (done == gridDim.x - 1) {
      p.state[1] = 0;
      p.state[2] = 0;
      p.state[3] = 0;
      __threadfence();
      p.state[0] = epoch;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// NVLS (multicast) variant of the two-shot kernel. `mc_partial` / `mc_gather` are MULTICAST addresses
// of the partial and gather buffers (a load through them reduces over every rank's copy inside the
// switch, a store lands in every rank's copy); `local_gather` is this rank's own gather buffer.
// Same flags/state protocol as the two-shot kernel, same epoch sequence.
struct ArMcParams {
  ArParams base;                 // bufs[] unused; flags/state/world/rank as above
  const void* mc_partial;
  void* mc_gather;
  const void* local_gather;
  int rows;
};

template <typename T>
__device__ __forceinline__ uint4 multimem_ld_reduce_add(const uint4* mc_addr);
template <>
__device__ __forceinline__ uint4 multimem_ld_reduce_add<__nv_bfloat16>(const uint4* mc_addr) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc_addr)
               : "memory");
  return v;
}
template <>
__device__ __forceinline__ uint4 multimem_ld_reduce_add<__half>(const uint4* mc_addr) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc_addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(uint4* mc_addr, const uint4& v) {
  // the payload is opaque 16-bit pairs: the type only selects the vector width
  asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(mc_addr), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

template <typename T>
__global__ void __launch_bounds__(256)
allreduce_add_rmsnorm_nvls_kernel(const ArMcParams q, T* __restrict__ out, T* __restrict__ residual,
                                  const T* __restrict__ weight, float eps, int hidden) {
  __shared__ float red[32];
  __shared__ uint32_t s_ticket;
  const ArParams& p = q.base;
  pdl_wait();
  const uint32_t epoch = p.state[0] + 1;
  if (threadIdx.x == 0) s_ticket = atomicAdd(&p.state[2], 1u);
  __syncthreads();
  if (s_ticket == 0 && threadIdx.x < p.world) {
    __threadfence_system();
    st_release_sys(p.flags[threadIdx.x] + p.rank, epoch);
  }
  if (threadIdx.x < p.world) {
    const uint32_t* mine = p.flags[p.rank] + threadIdx.x;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
    }
  }
  __syncthreads();

  // ---- phase 1: my slice of the rows -- in-switch sum, in-switch broadcast
  const int nvec = hidden / 8;
  const int slice = (q.rows + p.world - 1) / p.world;
  const int row0 = min(p.rank * slice, q.rows), row1 = min(row0 + slice, q.rows);
  const int64_t v_begin = (int64_t)row0 * nvec, v_end = (int64_t)row1 * nvec;
  const uint4* mc_in = reinterpret_cast<const uint4*>(q.mc_partial);
  uint4* mc_out = reinterpret_cast<uint4*>(q.mc_gather);
  for (int64_t idx = v_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < v_end;
       idx += (int64_t)gridDim.x * blockDim.x)
    multimem_st(mc_out + idx, multimem_ld_reduce_add<T>(mc_in + idx));
  // ---- barrier 2: the last CTA of this rank to finish its stores tells every peer
  // One fence per CTA, not per thread: bar.sync orders every thread's pushes before thread 0's
  // gpu-scope fence + ticket; the CTA that draws the last ticket issues the system-scope fence and
  // the release stores (fences are cumulative in the PTX memory model). 37 k threads each issuing
  // fence.sys cost several microseconds per call.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_ticket = atomicAdd(&p.state[3], 1u);
  }
  __syncthreads();
  if (s_ticket == gridDim.x - 1 && threadIdx.x < p.world) {
    __threadfence_system();
    st_release_sys(p.flags[threadIdx.x] + kMaxWorld + p.rank, epoch);
  }
  if (threadIdx.x < p.world) {
    const uint32_t* mine = p.flags[p.rank] + kMaxWorld + threadIdx.x;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
    }
  }
  __syncthreads();

  // ---- phase 2: residual add + RMSNorm over all rows from the local gather buffer
  const uint4* gathered = reinterpret_cast<const uint4*>(q.local_gather);
  for (int row = blockIdx.x; row < q.rows; row += gridDim.x) {
    add_rmsnorm_row<T>(row, hidden, out, residual, weight, eps, red, [&](int64_t idx, float (&acc)[8]) {
      const uint4 y = __ldcg(gathered + idx);
      const uint32_t w[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack2<T>(w[j]);
        acc[2 * j] = f.x;
        acc[2 * j + 1] = f.y;
      }
    });
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t done = atomicAdd(&p.state[1], 1u);
    if (done == gridDim.x - 1) {
      p.state[1] = 0;
      p.state[2] = 0;
      p.state[3] = 0;
      __threadfence();
      p.state[0] = epoch;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// NVLS one-shot: the one-shot kernel with the world-way pull replaced by ONE multimem.ld_reduce per 16
// bytes -- every rank asks the switch for the sum of every row (the switch reduces each request; this
// rank's links carry 1 x the message instead of (world-1) x), then adds the residual and normalises
// locally. One flag barrier per call (two alternating partial buffers, as the one-shot form), no slicing,
// so any row count works. Each rank's sums come from its own switch requests.
template <typename T>
__global__ void __launch_bounds__(256)
allreduce_add_rmsnorm_nvls_oneshot_kernel(const ArMcParams q, T* __restrict__ out, T* __restrict__ residual,
                                          const T* __restrict__ weight, float eps, int hidden) {
  __shared__ float red[32];
  __shared__ uint32_t s_ticket;
  const ArParams& p = q.base;
  pdl_wait();
  const uint32_t epoch = p.state[0] + 1;
  if (threadIdx.x == 0) s_ticket = atomicAdd(&p.state[2], 1u);
  __syncthreads();
  if (s_ticket == 0 && threadIdx.x < p.world) {
    __threadfence_system();
    st_release_sys(p.flags[threadIdx.x] + p.rank, epoch);
  }
  if (threadIdx.x < p.world) {
    const uint32_t* mine = p.flags[p.rank] + threadIdx.x;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
    }
  }
  __syncthreads();
  const uint4* mc_in = reinterpret_cast<const uint4*>(q.mc_partial);
  add_rmsnorm_row<T>(blockIdx.x, hidden, out, residual, weight, eps, red, [&](int64_t idx, float (&acc)[8]) {
    const uint4 y = multimem_ld_reduce_add<T>(mc_in + idx);
    const uint32_t w[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack2<T>(w[j]);
      acc[2 * j] = f.x;
      acc[2 * j + 1] = f.y;
    }
  });
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

// ---------------------------------------------------------------------------------------------
// NVLS, fully fused ("sequence-parallel residual"): rank r owns rows [r*slice, (r+1)*slice). For each
// of them it pulls the in-switch sum of the partials, adds ITS copy of the residual row, writes the new
// residual row LOCALLY and multicasts the normalised row into every rank's `out` region -- there is
// no second pass over all rows and the residual never crosses a link: every later all-reduce of the
// same forward pass needs a residual row only on the rank that owns it. Consequence for the caller:
// after this call `residual` is current only in the caller's own row slice.
struct McStore {
  __device__ __forceinline__ void operator()(uint4* dst, const uint4& v) const { multimem_st(dst, v); }
};

template <typename T>
__global__ void __launch_bounds__(256)
allreduce_add_rmsnorm_nvls_fused_kernel(const ArMcParams q, T* __restrict__ residual,
                                        const T* __restrict__ weight, float eps, int hidden) {
  __shared__ float red[32];
  __shared__ uint32_t s_ticket;
  const ArParams& p = q.base;
  pdl_wait();
  const uint32_t epoch = p.state[0] + 1;
  if (threadIdx.x == 0) s_ticket = atomicAdd(&p.state[2], 1u);
  __syncthreads();
  if (s_ticket == 0 && threadIdx.x < p.world) {
    __threadfence_system();
    st_release_sys(p.flags[threadIdx.x] + p.rank, epoch);
  }
  if (threadIdx.x < p.world) {
    const uint32_t* mine = p.flags[p.rank] + threadIdx.x;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
    }
  }
  __syncthreads();

  const int slice = (q.rows + p.world - 1) / p.world;
  const int row0 = min(p.rank * slice, q.rows), row1 = min(row0 + slice, q.rows);
  const uint4* mc_in = reinterpret_cast<const uint4*>(q.mc_partial);
  T* mc_out = reinterpret_cast<T*>(q.mc_gather);          // multicast address of the `out` region
  for (int row = row0 + blockIdx.x; row < row1; row += gridDim.x) {
    add_rmsnorm_row<T>(row, hidden, mc_out, residual, weight, eps, red,
                       [&](int64_t idx, float (&acc)[8]) {
                         const uint4 y = multimem_ld_reduce_add<T>(mc_in + idx);
                         const uint32_t w[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                         for (int j = 0; j < 4; ++j) {
                           const float2 f = unpack2<T>(w[j]);
                           acc[2 * j] = f.x;
                           acc[2 * j + 1] = f.y;
                         }
                       },
                       McStore());
    __syncthreads();
  }
  // ---- every rank's rows have landed in my `out` region once all peers passed this barrier
  // One fence per CTA, not per thread: bar.sync orders every thread's pushes before thread 0's
  // gpu-scope fence + ticket; the CTA that draws the last ticket issues the system-scope fence and
  // the release stores (fences are cumulative in the PTX memory model). 37 k threads each issuing
  // fence.sys cost several microseconds per call.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_ticket = atomicAdd(&p.state[3], 1u);
  }
  __syncthreads();
  if (s_ticket == gridDim.x - 1 && threadIdx.x < p.world) {
    __threadfence_system();
    st_release_sys(p.flags[threadIdx.x] + kMaxWorld + p.rank, epoch);
  }
  if (threadIdx.x < p.world) {
    const uint32_t* mine = p.flags[p.rank] + kMaxWorld + threadIdx.x;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t done = atomicAdd(&p.state[1], 1u);
    if (done == gridDim.x - 1) {
      p.state[1] = 0;
      p.state[2] = 0;
      p.state[3] = 0;
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

extern "C" int b200_allreduce_add_rmsnorm_twoshot(void* out, void* residual, const void* weight,
                                                  float eps, int32_t rows, int32_t hidden,
                                                  const void* const* peer_bufs,
                                                  void* const* peer_gather, void* const* peer_flags,
                                                  void* state, int32_t world, int32_t rank,
                                                  int32_t dtype, b200_stream_t stream) {
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(rows < 0 || hidden <= 0 || world < 2 || world > kMaxWorld || rank < 0 || rank >= world,
                 B200_ERR_BAD_SHAPE);
  if (rows == 0) return B200_OK;
  B200_RETURN_IF(!out || !residual || !weight || !peer_bufs || !peer_gather || !peer_flags || !state,
                 B200_ERR_NULL_POINTER);
  B200_RETURN_IF(hidden % 8 != 0 || !aligned16(out) || !aligned16(residual) || !aligned16(weight),
                 B200_ERR_MISALIGNED);
  B200_RETURN_IF(hidden > 8 * 256 * 4, B200_ERR_UNSUPPORTED);
  Ar2Params q;
  for (int r = 0; r < world; ++r) {
    B200_RETURN_IF(!peer_bufs[r] || !peer_gather[r] || !peer_flags[r], B200_ERR_NULL_POINTER);
    B200_RETURN_IF(!aligned16(peer_bufs[r]) || !aligned16(peer_gather[r]), B200_ERR_MISALIGNED);
    q.base.bufs[r] = peer_bufs[r];
    q.base.flags[r] = reinterpret_cast<uint32_t*>(peer_flags[r]);
    q.gather[r] = peer_gather[r];
  }
  q.base.state = reinterpret_cast<uint32_t*>(state);
  q.base.world = world;
  q.base.rank = rank;
  q.rows = rows;
  // every CTA spins on peer flags: keep the grid co-resident (<= one CTA per SM) and balanced
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int rows_per_cta = (rows + sms - 1) / sms;
  const int grid = (rows + rows_per_cta - 1) / rows_per_cta;
  if (dtype == B200_DTYPE_BF16)
    launch_pdl(allreduce_add_rmsnorm_twoshot_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0,
               (cudaStream_t)stream, q, (__nv_bfloat16*)out, (__nv_bfloat16*)residual,
               (const __nv_bfloat16*)weight, eps, hidden);
  else
    launch_pdl(allreduce_add_rmsnorm_twoshot_kernel<__half>, dim3(grid), dim3(256), 0,
               (cudaStream_t)stream, q, (__half*)out, (__half*)residual, (const __half*)weight, eps,
               hidden);
  return finish_launch();
}

extern "C" int b200_allreduce_add_rmsnorm_nvls(void* out, void* residual, const void* weight, float eps,
                                               int32_t rows, int32_t hidden, const void* mc_partial,
                                               void* mc_gather, const void* local_gather,
                                               void* const* peer_flags, void* state, int32_t world,
                                               int32_t rank, int32_t dtype, b200_stream_t stream) {
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(rows < 0 || hidden <= 0 || world < 2 || world > kMaxWorld || rank < 0 || rank >= world,
                 B200_ERR_BAD_SHAPE);
  if (rows == 0) return B200_OK;
  B200_RETURN_IF(!out || !residual || !weight || !mc_partial || !mc_gather || !local_gather ||
                     !peer_flags || !state,
                 B200_ERR_NULL_POINTER);
  B200_RETURN_IF(hidden % 8 != 0 || !aligned16(out) || !aligned16(residual) || !aligned16(weight) ||
                     !aligned16(mc_partial) || !aligned16(mc_gather) || !aligned16(local_gather),
                 B200_ERR_MISALIGNED);
  B200_RETURN_IF(hidden > 8 * 256 * 4, B200_ERR_UNSUPPORTED);
  ArMcParams q;
  for (int r = 0; r < kMaxWorld; ++r) {
    q.base.bufs[r] = nullptr;
    q.base.flags[r] = nullptr;
  }
  for (int r = 0; r < world; ++r) {
    B200_RETURN_IF(!peer_flags[r], B200_ERR_NULL_POINTER);
    q.base.flags[r] = reinterpret_cast<uint32_t*>(peer_flags[r]);
  }
  q.base.state = reinterpret_cast<uint32_t*>(state);
  q.base.world = world;
  q.base.rank = rank;
  q.mc_partial = mc_partial;
  q.mc_gather = mc_gather;
  q.local_gather = local_gather;
  q.rows = rows;
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int rows_per_cta = (rows + sms - 1) / sms;
  const int grid = (rows + rows_per_cta - 1) / rows_per_cta;
  if (dtype == B200_DTYPE_BF16)
    launch_pdl(allreduce_add_rmsnorm_nvls_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0,
               (cudaStream_t)stream, q, (__nv_bfloat16*)out, (__nv_bfloat16*)residual,
               (const __nv_bfloat16*)weight, eps, hidden);
  else
    launch_pdl(allreduce_add_rmsnorm_nvls_kernel<__half>, dim3(grid), dim3(256), 0, (cudaStream_t)stream,
               q, (__half*)out, (__half*)residual, (const __half*)weight, eps, hidden);
  return finish_launch();
}

extern "C" int b200_allreduce_add_rmsnorm_nvls_fused(void* residual, const void* weight, float eps,
                                                     int32_t rows, int32_t hidden, const void* mc_partial,
                                                     void* mc_out, void* const* peer_flags, void* state,
                                                     int32_t world, int32_t rank, int32_t dtype,
                                                     b200_stream_t stream) {
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(rows < 0 || hidden <= 0 || world < 2 || world > kMaxWorld || rank < 0 || rank >= world,
                 B200_ERR_BAD_SHAPE);
  if (rows == 0) return B200_OK;
  B200_RETURN_IF(!residual || !weight || !mc_partial || !mc_out || !peer_flags || !state,
                 B200_ERR_NULL_POINTER);
  B200_RETURN_IF(hidden % 8 != 0 || !aligned16(residual) || !aligned16(weight) || !aligned16(mc_partial) ||
                     !aligned16(mc_out),
                 B200_ERR_MISALIGNED);
  B200_RETURN_IF(hidden > 8 * 256 * 4, B200_ERR_UNSUPPORTED);
  ArMcParams q;
  for (int r = 0; r < kMaxWorld; ++r) {
    q.base.bufs[r] = nullptr;
    q.base.flags[r] = nullptr;
  }
  for (int r = 0; r < world; ++r) {
    B200_RETURN_IF(!peer_flags[r], B200_ERR_NULL_POINTER);
    q.base.flags[r] = reinterpret_cast<uint32_t*>(peer_flags[r]);
  }
  q.base.state = reinterpret_cast<uint32_t*>(state);
  q.base.world = world;
  q.base.rank = rank;
  q.mc_partial = mc_partial;
  q.mc_gather = mc_out;
  q.local_gather = nullptr;
  q.rows = rows;
  // one CTA per owned row, co-resident (every CTA spins on peer flags). Ranks whose slice is empty
  // still launch one CTA: it takes part in both barriers.
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int slice = (rows + world - 1) / world;
  const int grid = slice < 1 ? 1 : (slice > sms ? sms : slice);
  if (dtype == B200_DTYPE_BF16)
    launch_pdl(allreduce_add_rmsnorm_nvls_fused_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0,
               (cudaStream_t)stream, q, (__nv_bfloat16*)residual, (const __nv_bfloat16*)weight, eps, hidden);
  else
    launch_pdl(allreduce_add_rmsnorm_nvls_fused_kernel<__half>, dim3(grid), dim3(256), 0, (cudaStream_t)stream,
               q, (__half*)residual, (const __half*)weight, eps, hidden);
  return finish_launch();
}

extern "C" int b200_allreduce_add_rmsnorm_nvls_oneshot(void* out, void* residual, const void* weight, float eps,
                                                       int32_t rows, int32_t hidden, const void* mc_partial,
                                                       void* const* peer_flags, void* state, int32_t world,
                                                       int32_t rank, int32_t dtype, b200_stream_t stream) {
  B200_RETURN_IF(dtype != B200_DTYPE_BF16 && dtype != B200_DTYPE_F16, B200_ERR_BAD_DTYPE);
  B200_RETURN_IF(rows < 0 || hidden <= 0 || world < 2 || world > kMaxWorld || rank < 0 || rank >= world,
                 B200_ERR_BAD_SHAPE);
  if (rows == 0) return B200_OK;
  B200_RETURN_IF(!out || !residual || !weight || !mc_partial || !peer_flags || !state, B200_ERR_NULL_POINTER);
  B200_RETURN_IF(hidden % 8 != 0 || !aligned16(out) || !aligned16(residual) || !aligned16(weight) ||
                     !aligned16(mc_partial),
                 B200_ERR_MISALIGNED);
  B200_RETURN_IF(hidden > 8 * 256 * 4, B200_ERR_UNSUPPORTED);
  ArMcParams q;
  for (int r = 0; r < kMaxWorld; ++r) {
    q.base.bufs[r] = nullptr;
    q.base.flags[r] = nullptr;
  }
  for (int r = 0; r < world; ++r) {
    B200_RETURN_IF(!peer_flags[r], B200_ERR_NULL_POINTER);
    q.base.flags[r] = reinterpret_cast<uint32_t*>(peer_flags[r]);
  }
  q.base.state = reinterpret_cast<uint32_t*>(state);
  q.base.world = world;
  q.base.rank = rank;
  q.mc_partial = mc_partial;
  q.mc_gather = nullptr;
  q.local_gather = nullptr;
  q.rows = rows;
  // one CTA per row, all spinning on the start barrier: rows <= the co-resident capacity (8 CTAs/SM)
  B200_RETURN_IF(rows > 148 * 8, B200_ERR_UNSUPPORTED);
  if (dtype == B200_DTYPE_BF16)
    launch_pdl(allreduce_add_rmsnorm_nvls_oneshot_kernel<__nv_bfloat16>, dim3(rows), dim3(256), 0,
               (cudaStream_t)stream, q, (__nv_bfloat16*)out, (__nv_bfloat16*)residual,
               (const __nv_bfloat16*)weight, eps, hidden);
  else
    launch_pdl(allreduce_add_rmsnorm_nvls_oneshot_kernel<__half>, dim3(rows), dim3(256), 0, (cudaStream_t)stream,
               q, (__half*)out, (__half*)residual, (const __half*)weight, eps, hidden);
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
