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

// Same, memoised on the complete geometry (base, rank, dims, strides, box, swizzle): the launch lists replay the
// same few hundred descriptors every step and cuTensorMapEncodeTiled costs microseconds of host time.
int cached_tmap_16bit(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz);

// fp32 elements (relative-position bias tiles of the key-loop attention kernel)
int cached_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz);

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace fvit

#ifdef __CUDACC__
// erf-GELU (nn.GELU default, fv.py Mlp / ConvBlock) with erf from Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7):
// one ex2 + one rcp + 7 FMAs instead of libdevice erff's branchy polynomial -- the GEMM epilogues that apply it
// are instruction-issue bound. exp(-x^2/2) is shared between erf(x/sqrt2) and the normal pdf of the derivative.
__device__ __forceinline__ float fvit_rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));  // one MUFU (__frcp_rn is a subroutine call)
  return r;
}
__device__ __forceinline__ float fvit_ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fvit_erf_core(float x, float& e) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = fvit_rcp_approx(fmaf(0.3275911f, z, 1.0f));
  e = fvit_ex2_approx(-1.4426950408889634f * z * z);
  const float poly =
      fmaf(fmaf(fmaf(fmaf(1.061405429f, t, -1.453152027f), t, 1.421413741f), t, -0.284496736f), t, 0.254829592f) * t;
  return copysignf(fmaf(-poly, e, 1.0f), x);
}
__device__ __forceinline__ float fvit_gelu(float x) {
  float e;
  return 0.5f * x * (1.0f + fvit_erf_core(x, e));
}
// gelu(x) and gelu'(x) from one erf evaluation (the fc1 epilogue of the training forward saves gelu' so that the
// fc2 data-gradient epilogue only multiplies); the returned value is bit-identical to fvit_gelu(x)
__device__ __forceinline__ float fvit_gelu_both(float x, float& grad) {
  float e;
  const float cdf = 0.5f * (1.0f + fvit_erf_core(x, e));
  grad = fmaf(x * 0.39894228040143267794f, e, cdf);
  return x * cdf;
}
__device__ __forceinline__ float fvit_gelu_grad(float x) {
  float e;
  const float cdf = 0.5f * (1.0f + fvit_erf_core(x, e));
  return fmaf(x * 0.39894228040143267794f, e, cdf);
}
#endif
