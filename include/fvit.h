/* fvit.h — C ABI of libfvit_sm100.so: the B200 (sm_100a) kernels behind FasterViT.forward/backward.
 *
 * The reference (NVlabs/FasterViT) has no FFI: its hot path is torch.nn modules in
 * fastervit/models/faster_vit.py. Each entry point below replaces the ATen/cuDNN/cuBLAS call sequence
 * of the cited reference lines. Conventions:
 *   - plain device pointers + explicit sizes/strides; no C++/torch types cross the boundary;
 *   - every launch goes to the cudaStream_t passed as `stream` (a `void*`); nothing synchronises;
 *   - no allocation or free of caller memory; outputs/workspaces are caller-provided;
 *   - return 0 on success, non-zero on error; fvit_last_error() returns a per-thread message;
 *   - activations are row-major "token-major" (NHWC) matrices [rows, channels]; tensor-core operands
 *     are fp16 (or bf16) with fp32 accumulation, the residual stream and statistics are fp32.
 */
#ifndef FVIT_H_
#define FVIT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FVIT_ABI_VERSION 1

/* ---- library ------------------------------------------------------------------------------- */
int fvit_abi_version(void);
const char* fvit_last_error(void);
/* number of kernels launched by this library since load (or since last reset) — bench.py's
 * "gpu_launches" evidence. */
int64_t fvit_launch_count(void);
void fvit_reset_launch_count(void);

/* ---- epilogue activation codes --------------------------------------------------------------- */
enum {
  FVIT_ACT_NONE = 0,
  FVIT_ACT_RELU = 1,
  FVIT_ACT_GELU = 2,     /* exact erf GELU (nn.GELU default; fv.py:379,491) */
  FVIT_ACT_GELU_BWD = 3, /* v *= gelu'(aux[row,col])  (backward of the above) */
  FVIT_ACT_RELU_BWD = 4  /* v *= (aux[row,col] > 0) */
};

/* ---- tensor-core GEMM with shifted-row taps and fused epilogue --------------------------------
 * D[m, n] = epilogue( sum_{t < ntaps} sum_{k < kc} A[plane_t][m + shift_t][k] * B[n][t*kc_pad + k] )
 *
 *  - linear layers (fv.py:401-404, 545-547, 559, 566, 927): ntaps = 1;
 *  - 3x3 convolutions on zero-bordered NHWC activations (fv.py:434, 458-461, 489-492) are the same
 *    kernel with 9 taps whose row shift is (dy*(W+2)+dx) in the flattened padded pixel index, i.e.
 *    im2col-free implicit GEMM: every tap is a plain 2-D TMA box of the activation matrix;
 *  - backward passes use a_mn_major / b_mn_major = 1 (operand stored [K, MN]) and split_k.
 *
 * Operands are 16-bit (fp16, or bf16 when `bf16` != 0), accumulated in fp32 in tensor memory.
 * Epilogue per element:  v = acc * alpha
 *                        v = v * col_scale[n] + col_shift[n]          (each optional)
 *                        v = act(v [, aux])                            (FVIT_ACT_*)
 *                        v = v * col_scale2[n]                         (optional, after act)
 *                        v += resid[orow, n]                           (optional, fp32)
 *                        out_f32[orow, n] = v ; out_f16[orow, n] = (half)v   (each optional)
 * with orow = row_map ? row_map[m] : m, rows with orow < 0 skipped. If col_sum/col_sumsq are given,
 * sum and sum-of-squares of v (after alpha/scale/shift, before act) over valid rows are atomically
 * added per column (train-mode BatchNorm statistics). With split_k > 1 the partial products are
 * atomically added into out_f32 (which the caller zeroed) and only `alpha` is applied.
 */
typedef struct fvit_gemm_args {
  /* A operand: K-major: [a_planes][a_rows][kc] with row stride lda (elements);
   *            MN-major: [a_rows = K][m] with row stride lda. */
  const void* a;
  int64_t a_rows;
  int64_t lda;
  int64_t a_plane_stride; /* elements; 0 if a_planes == 1 */
  int32_t a_planes;
  int32_t a_mn_major;
  /* B operand: K-major: [n][ntaps*kc_pad] row stride ldb; MN-major: [b_rows = K][n] row stride ldb */
  const void* b;
  int64_t b_rows;
  int64_t ldb;
  int32_t b_mn_major;
  int32_t bf16;
  /* problem */
  int32_t m, n, kc, ntaps;
  int32_t tap_shift[16]; /* row shift of A per tap */
  int32_t tap_plane[16]; /* plane of A per tap */
  int32_t a_row_off;     /* extra row offset on A K-rows (MN-major) */
  int32_t b_row_off;     /* extra row offset on B K-rows (MN-major) */
  int32_t split_k;       /* >= 1 */
  int32_t tile_n;        /* 0 = auto; else multiple of 16 in [16, 256] */
  /* epilogue */
  float alpha;
  int32_t act;
  const float* col_scale;
  const float* col_shift;
  const float* col_scale2;
  const void* aux; /* fp16 [m, n] indexed by m (not orow), row stride ld_aux */
  int64_t ld_aux;
  const float* resid;
  int64_t ld_resid;
  const int32_t* row_map;
  float* out_f32;
  int64_t ld_out_f32;
  void* out_f16;
  int64_t ld_out_f16;
  float* col_sum;
  float* col_sumsq;
} fvit_gemm_args;

int fvit_gemm(const fvit_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FVIT_H_ */
