// api.cu -- library-level entry points of libb200attn.so (version, errors, launch counter).
#include "common.cuh"

namespace b200 {
long long g_launch_count = 0;
}

extern "C" int b200_version(void) { return B200ATTN_VERSION; }

extern "C" int64_t b200_launch_count(void) { return b200::g_launch_count; }
extern "C" void b200_reset_launch_count(void) { b200::g_launch_count = 0; }

extern "C" const char* b200_error_string(int code) {
  switch (code) {
    case B200_OK: return "success";
    case B200_ERR_NULL_POINTER: return "b200attn: required pointer argument is NULL";
    case B200_ERR_BAD_SHAPE: return "b200attn: inconsistent or non-positive shape/stride argument";
    case B200_ERR_UNSUPPORTED:
      return "b200attn: unsupported configuration (head_dim must be 64/128, GQA group <= 8, "
             "kv heads in {1,2,4,8k}, block_size % 16 == 0)";
    case B200_ERR_MISALIGNED: return "b200attn: pointer or stride violates the alignment contract";
    case B200_ERR_WORKSPACE: return "b200attn: workspace too small";
    case B200_ERR_DRIVER: return "b200attn: cuTensorMapEncodeTiled unavailable or failed";
    case B200_ERR_BAD_DTYPE: return "b200attn: dtype must be B200_DTYPE_BF16 or B200_DTYPE_F16";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "b200attn: unknown error code";
}
