// Host-side helpers shared by all translation units of libfvit_sm100.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace fvit {

// per-thread last error message (fvit_last_error)
char* err_buf();
int set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define FVIT_CHECK(cond, ...)                   \
  do {                                          \
    if (!(cond)) return fvit::set_error(__VA_ARGS__); \
  } while (0)

#define FVIT_CUDA(call)                                                                  \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fvit::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                             __LINE__);                                                  \
  } while (0)

inline int post_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int num_sms();

// cuTensorMapEncodeTiled fetched through the runtime (no link-time dependency on libcuda).
// dims/strides innermost-first; strides in bytes for dims 1..rank-1. 16-bit elements, zero OOB fill.
int encode_tmap_16bit(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz);

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace fvit
