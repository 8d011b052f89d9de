// fvit_gemm: persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   warp 0      : TMA producer (one elected lane) — cp.async.bulk.tensor boxes into a ring of
//                 128B-swizzled shared-memory stages, mbarrier complete_tx signalling
//   warp 1      : TMEM allocator + MMA issuer (one elected lane) — tcgen05.mma.kind::f16, M=128,
//                 N = tile_n (runtime, multiple of 16), K=16 per instruction, fp32 accumulators in
//                 TMEM, double-buffered (2 x 256 columns) so the epilogue of tile i overlaps the
//                 main loop of tile i+1
//   warps 2..9  : epilogue — tcgen05.ld (32 lanes x 32b x 32 columns), shared-memory transpose, fused
//                 scale/shift/activation/layer-scale/residual, optional per-column statistics,
//                 row-map scatter, coalesced 16-byte global accesses
//
// The A operand of K-block kb is the 2-D box at row (m0 + tap_shift[tap]) of plane tap_plane[tap]:
// with 9 taps this is an im2col-free 3x3 convolution over a zero-bordered NHWC activation matrix
// (out-of-range rows are zero-filled by TMA). See include/fvit.h for the exact contract.
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include <mutex>
#include <unordered_map>

#include "../../include/fvit.h"
#include "common.h"
#include "ptx.cuh"

namespace fvit {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 x 16-bit = 128 B: one SWIZZLE_128B row
constexpr int UMMA_K = 16;
// 12 epilogue warps: the epilogue is latency- and issue-bound (TMEM load, shared-memory transpose, global loads of the
// residual / aux rows), three warps per scheduler hide it better than two -- measured r02f: fv4 fwd+bwd GEMM time
// 47.5 -> 46.5 ms, fv0 14.7 -> 14.1 ms (A/B of two builds: -DFVIT_GEMM_NEPI=8 vs 12)
#ifndef FVIT_GEMM_NEPI
#define FVIT_GEMM_NEPI 12
#endif
constexpr int NEPI = FVIT_GEMM_NEPI;            // epilogue warps: a multiple of 4 (one TMEM lane quadrant each)
constexpr int EPI_THREADS = 32 * NEPI;
constexpr int GEMM_THREADS = 64 + EPI_THREADS;  // TMA warp + MMA warp + the epilogue warps
static_assert(NEPI % 4 == 0 && NEPI >= 8 && NEPI <= 16, "epilogue warps come in groups of four");
constexpr int MAX_STAGES = 8;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int TMEM_COLS = 512;
constexpr int MAX_ACC = 8;  // accumulator stages in TMEM: 512 / (64 | 128 | 256 columns)
constexpr int SMEM_BUDGET = 227 * 1024;
constexpr int SMEM_CTRL_BYTES = 1024;  // barriers + tmem base, placed after the stage ring
constexpr int SMEM_ALIGN_SLACK = 1024;
constexpr int STG_LD = 36;  // floats per staging row: 32 + 4 pad -> conflict-free float4 access both ways
constexpr int SMEM_STG_BYTES = NEPI * 32 * STG_LD * 4;  // one 32x32 fp32 staging tile per epilogue warp
constexpr int STATS_MAX_N = 1024;                    // widest GEMM with per-column statistics (BatchNorm channels)
constexpr int SMEM_STATS_BYTES = 2 * STATS_MAX_N * 4;

struct GemmParams {
  int m, n;
  int num_kb;      // K-blocks (of 64) over all taps
  int kb_per_tap;  // ceil(kc / 64)
  int k_tail16;    // 16-wide MMA steps that hold real K columns in a tap's last K-block (1..4); the rest is zero fill
  int a_mn, b_mn;
  int a_row_off, b_row_off;
  int tile_n;
  int tiles_m, tiles_n, split_k;
  int b_ntaps, tiles_per_tap;  // >1: N tiles enumerate (tap, n tile); tap t reads B rows shifted by tap_shift[t]
  int atomic_out;  // split-K style accumulation: atomicAdd(alpha * acc) into out_f32
  int stages;
  int cg;  // 1: one CTA per 128-row tile; 2: CTA pair (tcgen05 cta_group::2) per 256-row tile, B split over the pair
  int acc_stride, nacc;  // TMEM columns per accumulator stage and number of stages
  uint32_t idesc;
  int tap_shift[16];
  int tap_plane[16];
  // epilogue
  float alpha;
  int act;
  int bf16;
  int vec_ok;
  const float* col_scale;
  const float* col_shift;
  const float* col_scale2;
  const void* aux;
  long long ld_aux;
  const float* resid;
  long long ld_resid;
  const int* row_map;
  float* out_f32;
  long long ld_o32;
  void* out_f16;
  long long ld_o16;
  float* col_sum;
  float* col_sumsq;
  const float* alpha_ptr;   // optional device scalar multiplied into alpha
  const float* row_scale;   // optional fp32 [m] factor applied with col_scale2 (after act)
  void* out_pre16;          // optional fp16 copy of the value before col_scale2 / row_scale / resid
  long long ld_pre16;
  const float* aux_scale;   // optional per-column affine applied to aux before the activation derivative
  const float* aux_shift;
  float* osum;              // optional: osum[c] += *osum_alpha * sum over rows of the rounded out_f16 values
  const float* osum_alpha;
  int pre_is_grad;          // GELU epilogue stores gelu'(v) through out_pre16 instead of v
};

__device__ __forceinline__ float gelu_erf(float x) { return fvit_gelu(x); }
__device__ __forceinline__ float gelu_erf_grad(float x) { return fvit_gelu_grad(x); }

__device__ __forceinline__ float load16_as_float(const void* base, long long idx, int bf16) {
  if (bf16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
  return __half2float(reinterpret_cast<const __half*>(base)[idx]);
}
__device__ __forceinline__ uint32_t pack2_16(float a, float b, int bf16) {
  if (bf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// 256-bit global accesses (sm_100: one full 32-byte sector per lane and instruction)
__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_256(const void* ptr, uint32_t (&v)[8]) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(ptr));
}
__device__ __forceinline__ float2 unpack2_16(uint32_t w, int bf16) {
  if (bf16) return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}

__device__ __forceinline__ float4 ld_vec4_guard(const float* p, int valid, float fill) {
  if (valid >= 4 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) return __ldg(reinterpret_cast<const float4*>(p));
  float4 r = make_float4(fill, fill, fill, fill);
  if (valid > 0) r.x = __ldg(p);
  if (valid > 1) r.y = __ldg(p + 1);
  if (valid > 2) r.z = __ldg(p + 2);
  if (valid > 3) r.w = __ldg(p + 3);
  return r;
}
__device__ __forceinline__ void unpack4_16(uint2 pk, int bf16, float (&a)[4]) {
  if (bf16) {
    const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&pk.x);
    const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&pk.y);
    a[0] = __low2float(lo), a[1] = __high2float(lo), a[2] = __low2float(hi), a[3] = __high2float(hi);
  } else {
    const __half2 lo = *reinterpret_cast<const __half2*>(&pk.x);
    const __half2 hi = *reinterpret_cast<const __half2*>(&pk.y);
    a[0] = __low2float(lo), a[1] = __high2float(lo), a[2] = __low2float(hi), a[3] = __high2float(hi);
  }
}

// Sum 16 per-lane values over the 32 lanes of a warp with 16 shuffles (recursive halving).
// On return lane L (even lanes are the owners) holds in v[0] the total of column
// ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1).
__device__ __forceinline__ void warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int width = 8, mask = 16; width >= 1; width >>= 1, mask >>= 1) {
    const bool upper = (lane & mask) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < width) {
        const float send = upper ? v[i] : v[i + width];
        const float keep = upper ? v[i + width] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
      }
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}
__device__ __forceinline__ int colsum16_owner_col(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

// Epilogue specialisation: GENERIC = every option decided at run time (backward modes, statistics,
// split-K atomics); otherwise ACT / RESID / O32 / O16 are compile-time and the unused paths vanish
// (the epilogue is instruction-issue bound, so this is worth ~2.5x on short-K tiles).
// FEAT: compile-time feature mask (see EF_* below); EF_GENERIC keeps every switch at run time.
enum : uint32_t {
  EF_GENERIC = 1u << 0,
  EF_ACT_SHIFT = 1,          // 3 bits: FVIT_ACT_* code
  EF_RESID = 1u << 4,
  EF_O32 = 1u << 5,
  EF_O16 = 1u << 6,
  EF_STATS = 1u << 7,
  EF_PRE = 1u << 8,
  EF_ATOMIC = 1u << 9,
  EF_ALPHAPTR = 1u << 10,
  EF_CS2 = 1u << 11,
  EF_RS = 1u << 12,
  EF_OSUM = 1u << 13,        // per-column sum of the rounded 16-bit output (bias gradient of the next layer)
  EF_PREGRAD = 1u << 14,     // out_pre16 receives gelu'(v) (consumed by FVIT_ACT_MUL_AUX in the backward pass)
  EF_DIRECT = 1u << 15,      // 16-bit outputs only: row-per-thread epilogue straight from TMEM (no staging tile), 32-byte stores
};
// work unit -> (M tile, N tile, K split). Plain GEMMs keep the K splits of a tile adjacent; the tap-in-N mode
// walks all tiles of one K range first so concurrently running CTAs share the A / B slices in L2.
__device__ __forceinline__ void decode_work(const GemmParams& p, int w, int& tm, int& tn, int& split) {
  int t;
  if (p.b_ntaps > 1) {
    const int nt = p.tiles_m * p.tiles_n;
    split = w / nt;
    t = w - split * nt;
  } else {
    split = w % p.split_k;
    t = w / p.split_k;
  }
  tn = t % p.tiles_n;
  tm = t / p.tiles_n;
}

// CG = 1: one CTA per 128-row tile (no cluster instruction is compiled in); CG = 2: CTA pair (cluster launch).
template <uint32_t FEAT, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                        const __grid_constant__ CUtensorMap tmap_b,
                        const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // SW128 atoms
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // CTA pair: this CTA stages its 128 rows of A and its half of the B tile; the leader issues M = 256 MMAs that
  // read both CTAs' shared memory at the same offsets and write 128 accumulator rows into each CTA's TMEM
  constexpr int cg = CG;
  uint32_t cta_rank = 0u;
  if constexpr (CG == 2) cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int unit0 = blockIdx.x / cg, nunits = gridDim.x / cg;  // work units are dealt to CTAs / CTA pairs
  const int b_cols = p.tile_n / cg;                            // B rows (N columns) staged by this CTA
  // MN-major B is staged as whole 64-column atoms (the MMA reads the first b_cols columns of them)
  const int b_stage_bytes = (p.b_mn ? ((b_cols + 63) & ~63) : b_cols) * BK * 2;
  const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;
  uint8_t* ctrl = smem + p.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* acc_full = empty_bar + MAX_STAGES;
  uint64_t* acc_empty = acc_full + MAX_ACC;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(acc_empty + MAX_ACC);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < p.nacc; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], NEPI * cg);  // the leader's copy collects the epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) tmem_alloc_cg2(tmem_base_slot, TMEM_COLS);
    else tmem_alloc(tmem_base_slot, TMEM_COLS);
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();  // peer barriers are initialised before any remote arrive / TMA signal
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int work_total = p.tiles_m * p.tiles_n * p.split_k;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = unit0; w < work_total; w += nunits) {
        int tm, tn, split;
        decode_work(p, w, tm, tn, split);
        const int m0 = (tm * cg + (int)cta_rank) * BM;
        int n0 = tn * p.tile_n;
        int b_shift = p.b_row_off;
        if (p.b_ntaps > 1) {
          const int tap = tn / p.tiles_per_tap;
          n0 = (tn - tap * p.tiles_per_tap) * p.tile_n;
          b_shift += p.tap_shift[tap];
        }
        n0 += (int)cta_rank * b_cols;
        const int kb0 = (int)((long long)p.num_kb * split / p.split_k);
        const int kb1 = (int)((long long)p.num_kb * (split + 1) / p.split_k);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + A_STAGE_BYTES;
          // both CTAs' boxes complete on the leader's barrier, which expects the bytes of the whole pair
          if (leader) mbar_expect_tx(&full_bar[stage], (uint32_t)(stage_bytes * cg));
          if constexpr (CG == 1) {
            if (!p.a_mn) {
              const int tap = kb / p.kb_per_tap;
              const int kc0 = (kb - tap * p.kb_per_tap) * BK;
              tma_load_3d(sa, &tmap_a, &full_bar[stage], kc0, m0 + p.tap_shift[tap], p.tap_plane[tap]);
            } else {
              // A stored [K rows][M cols]: two 64-wide MN atoms of BK rows each
              tma_load_3d(sa, &tmap_a, &full_bar[stage], m0, kb * BK + p.a_row_off, 0);
              tma_load_3d(sa + BK * 128, &tmap_a, &full_bar[stage], m0 + 64, kb * BK + p.a_row_off, 0);
            }
            if (!p.b_mn) {
              tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n0);
            } else {
              for (int j = 0; j < (b_cols + 63) / 64; ++j)
                tma_load_2d(sb + j * BK * 128, &tmap_b, &full_bar[stage], n0 + j * 64, kb * BK + b_shift);
            }
          } else {
            const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);  // (only instantiated for CG == 2)
            if (!p.a_mn) {
              const int tap = kb / p.kb_per_tap;
              const int kc0 = (kb - tap * p.kb_per_tap) * BK;
              tma_load_3d_cg2(sa, &tmap_a, fb, kc0, m0 + p.tap_shift[tap], p.tap_plane[tap]);
            } else {
              tma_load_3d_cg2(sa, &tmap_a, fb, m0, kb * BK + p.a_row_off, 0);
              tma_load_3d_cg2(sa + BK * 128, &tmap_a, fb, m0 + 64, kb * BK + p.a_row_off, 0);
            }
            if (!p.b_mn) {
              tma_load_2d_cg2(sb, &tmap_b, fb, kb * BK, n0);
            } else {
              for (int j = 0; j < (b_cols + 63) / 64; ++j)
                tma_load_2d_cg2(sb + j * BK * 128, &tmap_b, fb, n0 + j * 64, kb * BK + b_shift);
            }
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0 && leader) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      // descriptor templates: K-major SW128: SBO = 1024 (8 rows x 128 B), LBO unused;
      // MN-major SW128: LBO = BK*128 (next 64-wide MN atom), SBO = 1024 (next 8 K rows).
      const uint32_t a_lbo = p.a_mn ? BK * 128 : 16;
      const uint32_t b_lbo = p.b_mn ? BK * 128 : 16;
      const uint32_t a_kstep = p.a_mn ? UMMA_K * 128 : UMMA_K * 2;
      const uint32_t b_kstep = p.b_mn ? UMMA_K * 128 : UMMA_K * 2;
      for (int w = unit0; w < work_total; w += nunits) {
        int tm_, tn_, split;
        decode_work(p, w, tm_, tn_, split);
        const int kb0 = (int)((long long)p.num_kb * split / p.split_k);
        const int kb1 = (int)((long long)p.num_kb * (split + 1) / p.split_k);
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * p.acc_stride;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint32_t sb = sa + A_STAGE_BYTES;
          // a tap's last K-block is zero beyond kc (TMA out-of-bounds fill / zero-padded weights): its all-zero
          // 16-wide steps are not issued (K = 196 per conv tap: 13 instead of 16 MMAs; K = 784: 49 instead of 52)
          const int nk = (kb + 1) % p.kb_per_tap == 0 ? p.k_tail16 : BK / UMMA_K;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            if (k >= nk) break;
            const uint64_t adesc = make_smem_desc(sa + k * a_kstep, a_lbo, 1024, SWZ_128B);
            const uint64_t bdesc = make_smem_desc(sb + k * b_kstep, b_lbo, 1024, SWZ_128B);
            if constexpr (CG == 2) umma_f16_ss_cg2(d_tmem, adesc, bdesc, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_f16_ss(d_tmem, adesc, bdesc, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // frees the smem stage (in both CTAs of a pair) once these MMAs retire
          if constexpr (CG == 2) umma_commit_cg2(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if constexpr (CG == 2) umma_commit_cg2(&acc_full[acc]);
        else umma_commit(&acc_full[acc]);
        if (++acc == p.nacc) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================================================================== epilogue warps
    // 8 warps: warp w owns TMEM lane quadrant (w & 3) (32 rows of the tile) and every second
    // 32-column chunk ((w - 2) >> 2 selects even / odd chunks). Per chunk a warp
    //   1. issues its residual / aux global loads (addresses do not depend on the accumulator),
    //   2. pulls the accumulators row-per-thread with tcgen05.ld and transposes them through a padded
    //      shared-memory staging tile,
    //   3. works "coalesced": one instruction covers 4 rows x 128 contiguous bytes (8 lanes x float4
    //      per row), so residual loads and fp32 / fp16 stores are full-line accesses.
    // The epilogue is bound by global-memory latency, hence the loads hoisted ahead of the TMEM read
    // and the 256 threads: ~32 KB of residual reads are in flight per SM.
    const int ew = warp - 2;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int chunk_par = ew >> 2;
    float* stg = reinterpret_cast<float*>(ctrl + SMEM_CTRL_BYTES) + ew * (32 * STG_LD);
    // per-CTA partial BatchNorm statistics (sum | sum of squares), flushed with one atomic per column at the end
    float* stat_s = reinterpret_cast<float*>(ctrl + SMEM_CTRL_BYTES + SMEM_STG_BYTES);
    const int sub = lane >> 3;   // row within a 4-row group
    const int c4 = lane & 7;     // which float4 of the 32-column chunk
    constexpr bool GENERIC = (FEAT & EF_GENERIC) != 0;
    constexpr int ACT = (FEAT >> EF_ACT_SHIFT) & 7;
    const bool atomic_out = GENERIC ? (p.atomic_out != 0) : ((FEAT & EF_ATOMIC) != 0);
    const bool use_resid = GENERIC ? (p.resid != nullptr && !p.atomic_out) : ((FEAT & EF_RESID) != 0);
    const int act = GENERIC ? p.act : ACT;
    const bool use_aux = (act == FVIT_ACT_GELU_BWD || act == FVIT_ACT_RELU_BWD || act == FVIT_ACT_MUL_AUX) && !atomic_out;
    const bool out32 = GENERIC ? (p.out_f32 != nullptr) : ((FEAT & EF_O32) != 0);
    const bool out16 = GENERIC ? (p.out_f16 != nullptr) : ((FEAT & EF_O16) != 0);
    const bool stats = GENERIC ? (p.col_sum != nullptr && !p.atomic_out) : ((FEAT & EF_STATS) != 0);
    const bool has_cs2 = GENERIC ? (p.col_scale2 != nullptr) : ((FEAT & EF_CS2) != 0);
    const bool has_rs = GENERIC ? (p.row_scale != nullptr) : ((FEAT & EF_RS) != 0);
    const bool has_pre = GENERIC ? (p.out_pre16 != nullptr) : ((FEAT & EF_PRE) != 0);
    const bool pregrad = GENERIC ? (p.pre_is_grad != 0) : ((FEAT & EF_PREGRAD) != 0);
    const bool has_aptr = GENERIC ? (p.alpha_ptr != nullptr) : ((FEAT & EF_ALPHAPTR) != 0);
    const bool osum = GENERIC ? (p.osum != nullptr) : ((FEAT & EF_OSUM) != 0);
    const float alpha = has_aptr ? p.alpha * __ldg(p.alpha_ptr) : p.alpha;
    if (stats) {
      for (int i = threadIdx.x - 64; i < 2 * p.n; i += GEMM_THREADS - 64) stat_s[i] = 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");  // the epilogue warps only
    }
    if (osum) {  // tile-local column sums: stat_s[0 .. tile_n)
      if (threadIdx.x - 64 < 256) stat_s[threadIdx.x - 64] = 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
    }
    const float osum_alpha = (osum && p.osum_alpha) ? __ldg(p.osum_alpha) : 1.f;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t acc_empty_leader = 0u;
    if constexpr (CG == 2) acc_empty_leader = mapa_shared(smem_u32(&acc_empty[0]), 0);
    for (int w = unit0; w < work_total; w += nunits) {
      int tm, tn, split_;
      decode_work(p, w, tm, tn, split_);
      const int row_base = (tm * cg + (int)cta_rank) * BM + quad * 32;
      int n0 = tn * p.tile_n;
      int n_end = min(p.n, n0 + p.tile_n);
      if (p.b_ntaps > 1) {  // output columns [tap * n, (tap + 1) * n)
        const int tap = tn / p.tiles_per_tap;
        n0 = tap * p.n + (tn - tap * p.tiles_per_tap) * p.tile_n;
        n_end = min((tap + 1) * p.n, n0 + p.tile_n);
      }
      // output rows of the 8 tile rows this lane touches (4k + sub), fetched once per tile
      long long orow[8];
      {
        long long my_orow = -1;
        if (row_base + lane < p.m)
          my_orow = p.row_map ? (long long)p.row_map[row_base + lane] : (long long)(row_base + lane);
#pragma unroll
        for (int k = 0; k < 8; ++k) orow[k] = __shfl_sync(0xffffffffu, my_orow, 4 * k + sub);
      }

      bool waited = false;
      const uint32_t taddr = tmem_base + acc * p.acc_stride + ((uint32_t)(quad * 32) << 16);

      if constexpr ((FEAT & EF_DIRECT) != 0) {
        // Row-per-thread epilogue for launches that only write 16-bit outputs (qkv / fc1 / data gradients): lane l
        // owns tile row quad*32 + l, computes on the 32 accumulator columns tcgen05.ld hands it and stores them as
        // two 32-byte sectors per output -- no staging tile, no per-row address math, one bounds check per 16 columns.
        // (n, the leading dimensions and the base pointers are multiples of 16 elements: checked by the launcher.)
        const int my_row = row_base + lane;
        const long long my_orow =
            my_row < p.m ? (p.row_map ? (long long)p.row_map[my_row] : (long long)my_row) : -1;
        uint16_t* const orow16 = reinterpret_cast<uint16_t*>(p.out_f16) + (my_orow >= 0 ? my_orow : 0) * p.ld_o16;
        uint16_t* const prow16 = reinterpret_cast<uint16_t*>(p.out_pre16) + (long long)my_row * p.ld_pre16;
        const uint16_t* const arow16 = reinterpret_cast<const uint16_t*>(p.aux) + (long long)my_row * p.ld_aux;
        for (int c0 = chunk_par * 32; c0 < p.tile_n; c0 += 32 * (NEPI / 4)) {
          const int nbase = n0 + c0;
          if (nbase >= n_end) break;  // warp-uniform
          const int nvalid = min(min(32, n_end - nbase), p.tile_n - c0);   // 16 or 32
          uint32_t axw[2][8];
          if (use_aux && my_orow >= 0) {
            ld_global_256(arow16 + nbase, axw[0]);
            if (nvalid > 16) ld_global_256(arow16 + nbase + 16, axw[1]);
          }
          if (!waited) {
            mbar_wait(&acc_full[acc], acc_phase);
            tc_fence_after();
            waited = true;
          }
          uint32_t raw[32];
          if (nvalid > 16) {
            tmem_ld32(taddr + c0, raw);
          } else {
            uint32_t lo[16];
            tmem_ld16(taddr + c0, lo);
#pragma unroll
            for (int i = 0; i < 16; ++i) raw[i] = lo[i], raw[16 + i] = 0u;
          }
          tmem_ld_wait();
          float osv[32];  // rounded outputs for the column sums
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            if (hf * 16 >= nvalid) {
              if (osum) {
#pragma unroll
                for (int i = 0; i < 16; ++i) osv[hf * 16 + i] = 0.f;
              }
              continue;
            }
            const int col = nbase + hf * 16;
            uint32_t ow[8], pw[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {  // 4 columns at a time: per-column vectors are warp-uniform 16-byte loads
              float4 cs = make_float4(alpha, alpha, alpha, alpha), sh = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.col_scale) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(p.col_scale + col) + q);
                cs.x *= t.x, cs.y *= t.y, cs.z *= t.z, cs.w *= t.w;
              }
              if (p.col_shift) sh = __ldg(reinterpret_cast<const float4*>(p.col_shift + col) + q);
              float v[4] = {fmaf(__uint_as_float(raw[hf * 16 + 4 * q]), cs.x, sh.x),
                            fmaf(__uint_as_float(raw[hf * 16 + 4 * q + 1]), cs.y, sh.y),
                            fmaf(__uint_as_float(raw[hf * 16 + 4 * q + 2]), cs.z, sh.z),
                            fmaf(__uint_as_float(raw[hf * 16 + 4 * q + 3]), cs.w, sh.w)};
              float pv[4] = {v[0], v[1], v[2], v[3]};
              if (act == FVIT_ACT_RELU) {
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
              } else if (act == FVIT_ACT_GELU) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (has_pre && pregrad) v[e] = fvit_gelu_both(v[e], pv[e]);
                  else v[e] = gelu_erf(v[e]);
                }
              }
              if (use_aux) {
                const float2 a01 = unpack2_16(axw[hf][2 * q], p.bf16), a23 = unpack2_16(axw[hf][2 * q + 1], p.bf16);
                float a[4] = {a01.x, a01.y, a23.x, a23.y};
                if (act == FVIT_ACT_MUL_AUX) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) v[e] *= a[e];
                } else {
                  if (p.aux_scale) {
                    const float4 asc = __ldg(reinterpret_cast<const float4*>(p.aux_scale + col) + q);
                    const float4 ash = __ldg(reinterpret_cast<const float4*>(p.aux_shift + col) + q);
                    a[0] = fmaf(a[0], asc.x, ash.x), a[1] = fmaf(a[1], asc.y, ash.y), a[2] = fmaf(a[2], asc.z, ash.z),
                    a[3] = fmaf(a[3], asc.w, ash.w);
                  }
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    if (act == FVIT_ACT_GELU_BWD) v[e] *= gelu_erf_grad(a[e]);
                    else v[e] = a[e] > 0.f ? v[e] : 0.f;
                  }
                }
              }
              ow[2 * q] = pack2_16(v[0], v[1], p.bf16), ow[2 * q + 1] = pack2_16(v[2], v[3], p.bf16);
              if (has_pre) pw[2 * q] = pack2_16(pv[0], pv[1], p.bf16), pw[2 * q + 1] = pack2_16(pv[2], pv[3], p.bf16);
              if (osum) {
                const float2 r01 = unpack2_16(ow[2 * q], p.bf16), r23 = unpack2_16(ow[2 * q + 1], p.bf16);
                const bool okr = my_orow >= 0;
                osv[hf * 16 + 4 * q] = okr ? r01.x : 0.f, osv[hf * 16 + 4 * q + 1] = okr ? r01.y : 0.f;
                osv[hf * 16 + 4 * q + 2] = okr ? r23.x : 0.f, osv[hf * 16 + 4 * q + 3] = okr ? r23.y : 0.f;
              }
            }
            if (my_orow >= 0) {
              st_global_256(orow16 + col, ow);
              if (has_pre) st_global_256(prow16 + col, pw);
            }
          }
          if (osum) {
            // column sums over the warp's 32 rows: butterfly that halves the live values per step; lane l ends
            // up with column c0 + l
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) {
              const bool up = (lane & sft) != 0;
#pragma unroll
              for (int i = 0; i < sft; ++i) {
                const float send = up ? osv[i] : osv[i + sft];
                const float keep = up ? osv[i + sft] : osv[i];
                osv[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
              }
            }
            if (lane < nvalid && osv[0] != 0.f) atomicAdd(stat_s + c0 + lane, osv[0]);
          }
        }
      } else
      for (int c0 = chunk_par * 32; c0 < p.tile_n; c0 += 32 * (NEPI / 4)) {
        const int nbase = n0 + c0;
        if (nbase >= n_end) break;  // warp-uniform
        const int col = nbase + 4 * c4;            // first of this lane's 4 columns
        const bool cfull = col + 4 <= n_end;        // all 4 columns valid
        const bool cany = col < n_end;
        // ---- 1. global loads first
        float4 rv[8];
        uint2 av[8];
        if (use_resid) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            rv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (cany && orow[k] >= 0) {
              const float* r = p.resid + orow[k] * p.ld_resid + col;
              if (cfull && p.vec_ok) {
                rv[k] = *reinterpret_cast<const float4*>(r);
              } else {
                rv[k].x = r[0];
                if (col + 1 < n_end) rv[k].y = r[1];
                if (col + 2 < n_end) rv[k].z = r[2];
                if (col + 3 < n_end) rv[k].w = r[3];
              }
            }
          }
        }
        if (use_aux) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            av[k] = make_uint2(0u, 0u);
            if (cany && orow[k] >= 0) {
              const long long ab = (long long)(row_base + 4 * k + sub) * p.ld_aux + col;
              const uint16_t* ap = reinterpret_cast<const uint16_t*>(p.aux) + ab;
              if (cfull && p.vec_ok) {
                av[k] = *reinterpret_cast<const uint2*>(ap);
              } else {
                uint32_t e[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (col + i < n_end) e[i] = ap[i];
                av[k] = make_uint2(e[0] | (e[1] << 16), e[2] | (e[3] << 16));
              }
            }
          }
        }
        float4 cs = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f),
               cs2 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (cany) {
          if (p.col_scale) cs = ld_vec4_guard(p.col_scale + col, n_end - col, 1.f);
          if (p.col_shift) sh = ld_vec4_guard(p.col_shift + col, n_end - col, 0.f);
          if (has_cs2) cs2 = ld_vec4_guard(p.col_scale2 + col, n_end - col, 1.f);
        }
        cs.x *= alpha, cs.y *= alpha, cs.z *= alpha, cs.w *= alpha;  // fold alpha
        float4 asc = make_float4(1.f, 1.f, 1.f, 1.f), ash = make_float4(0.f, 0.f, 0.f, 0.f);
        if (use_aux && p.aux_scale && cany) {
          asc = ld_vec4_guard(p.aux_scale + col, n_end - col, 1.f);
          ash = ld_vec4_guard(p.aux_shift + col, n_end - col, 0.f);
        }
        // ---- 2. accumulators: TMEM -> registers -> staging tile
        if (!waited) {
          mbar_wait(&acc_full[acc], acc_phase);
          tc_fence_after();
          waited = true;
        }
        {
          uint32_t raw[32];
          if (c0 + 32 <= p.tile_n) {
            tmem_ld32(taddr + c0, raw);
          } else {  // tile_n is a multiple of 16: last half chunk
            uint32_t lo[16];
            tmem_ld16(taddr + c0, lo);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              raw[i] = lo[i];
              raw[16 + i] = 0u;
            }
          }
          tmem_ld_wait();
          float4* srow = reinterpret_cast<float4*>(stg + lane * STG_LD);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            srow[j] = make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]),
                                  __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3]));
        }
        __syncwarp();
        // ---- 3. coalesced epilogue math + stores
        float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int rl = 4 * k + sub;  // row within the warp's 32
          const bool ok = cany && orow[k] >= 0;
          float4 v = *reinterpret_cast<const float4*>(stg + rl * STG_LD + 4 * c4);
          if (atomic_out) {
            v.x *= alpha, v.y *= alpha, v.z *= alpha, v.w *= alpha;
            if (ok) {
              float* o = p.out_f32 + orow[k] * p.ld_o32 + col;
              if (cfull && p.vec_ok) {  // one 16-byte reduction instead of four (L2 atomic throughput bound)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v.x), "f"(v.y), "f"(v.z),
                             "f"(v.w)
                             : "memory");
              } else {
                atomicAdd(o, v.x);
                if (col + 1 < n_end) atomicAdd(o + 1, v.y);
                if (col + 2 < n_end) atomicAdd(o + 2, v.z);
                if (col + 3 < n_end) atomicAdd(o + 3, v.w);
              }
            }
            continue;
          }
          v.x = fmaf(v.x, cs.x, sh.x), v.y = fmaf(v.y, cs.y, sh.y), v.z = fmaf(v.z, cs.z, sh.z),
          v.w = fmaf(v.w, cs.w, sh.w);
          if (stats && ok) {
            s1.x += v.x, s1.y += v.y, s1.z += v.z, s1.w += v.w;
            s2.x += v.x * v.x, s2.y += v.y * v.y, s2.z += v.z * v.z, s2.w += v.w * v.w;
          }
          if (!ok) continue;
          float4 pv = v;  // what out_pre16 receives: the pre-activation value, or gelu'(v) with pre_is_grad
          if (act == FVIT_ACT_RELU) {
            v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
          } else if (act == FVIT_ACT_GELU) {
            if (has_pre && pregrad) {
              v.x = fvit_gelu_both(v.x, pv.x), v.y = fvit_gelu_both(v.y, pv.y), v.z = fvit_gelu_both(v.z, pv.z),
              v.w = fvit_gelu_both(v.w, pv.w);
            } else {
              v.x = gelu_erf(v.x), v.y = gelu_erf(v.y), v.z = gelu_erf(v.z), v.w = gelu_erf(v.w);
            }
          }
          if (has_pre) {
            uint16_t* o = reinterpret_cast<uint16_t*>(p.out_pre16) + (long long)(row_base + rl) * p.ld_pre16 + col;
            const uint32_t lo = pack2_16(pv.x, pv.y, p.bf16), hi = pack2_16(pv.z, pv.w, p.bf16);
            if (cfull && p.vec_ok) {
              *reinterpret_cast<uint2*>(o) = make_uint2(lo, hi);
            } else {
              o[0] = (uint16_t)(lo & 0xFFFF);
              if (col + 1 < n_end) o[1] = (uint16_t)(lo >> 16);
              if (col + 2 < n_end) o[2] = (uint16_t)(hi & 0xFFFF);
              if (col + 3 < n_end) o[3] = (uint16_t)(hi >> 16);
            }
          }
          if (act == FVIT_ACT_MUL_AUX) {
            float a[4];
            unpack4_16(av[k], p.bf16, a);
            v.x *= a[0], v.y *= a[1], v.z *= a[2], v.w *= a[3];
          } else if (use_aux) {
            float a[4];
            unpack4_16(av[k], p.bf16, a);
            if (p.aux_scale) {  // aux holds the raw convolution output: pre-activation = raw * scale + shift (BatchNorm)
              a[0] = fmaf(a[0], asc.x, ash.x), a[1] = fmaf(a[1], asc.y, ash.y), a[2] = fmaf(a[2], asc.z, ash.z),
              a[3] = fmaf(a[3], asc.w, ash.w);
            }
            if (act == FVIT_ACT_GELU_BWD) {
              v.x *= gelu_erf_grad(a[0]), v.y *= gelu_erf_grad(a[1]), v.z *= gelu_erf_grad(a[2]),
              v.w *= gelu_erf_grad(a[3]);
            } else {
              v.x = a[0] > 0.f ? v.x : 0.f, v.y = a[1] > 0.f ? v.y : 0.f, v.z = a[2] > 0.f ? v.z : 0.f,
              v.w = a[3] > 0.f ? v.w : 0.f;
            }
          }
          if (has_cs2) v.x *= cs2.x, v.y *= cs2.y, v.z *= cs2.z, v.w *= cs2.w;
          if (has_rs) {
            const float rs = __ldg(p.row_scale + row_base + rl);
            v.x *= rs, v.y *= rs, v.z *= rs, v.w *= rs;
          }
          if (use_resid) v.x += rv[k].x, v.y += rv[k].y, v.z += rv[k].z, v.w += rv[k].w;
          if (out32) {
            float* o = p.out_f32 + orow[k] * p.ld_o32 + col;
            if (cfull && p.vec_ok) {
              *reinterpret_cast<float4*>(o) = v;
            } else {
              o[0] = v.x;
              if (col + 1 < n_end) o[1] = v.y;
              if (col + 2 < n_end) o[2] = v.z;
              if (col + 3 < n_end) o[3] = v.w;
            }
          }
          if (out16) {
            uint16_t* o = reinterpret_cast<uint16_t*>(p.out_f16) + orow[k] * p.ld_o16 + col;
            const uint32_t lo = pack2_16(v.x, v.y, p.bf16), hi = pack2_16(v.z, v.w, p.bf16);
            if (cfull && p.vec_ok) {
              *reinterpret_cast<uint2*>(o) = make_uint2(lo, hi);
            } else {
              o[0] = (uint16_t)(lo & 0xFFFF);
              if (col + 1 < n_end) o[1] = (uint16_t)(lo >> 16);
              if (col + 2 < n_end) o[2] = (uint16_t)(hi & 0xFFFF);
              if (col + 3 < n_end) o[3] = (uint16_t)(hi >> 16);
            }
            if (osum) {
              float r4[4];
              unpack4_16(make_uint2(lo, hi), p.bf16, r4);
              s1.x += r4[0], s1.y += r4[1], s1.z += r4[2], s1.w += r4[3];
            }
          }
        }
        __syncwarp();
        if (osum) {
#pragma unroll
          for (int o = 8; o <= 16; o <<= 1) {
            s1.x += __shfl_xor_sync(0xffffffffu, s1.x, o), s1.y += __shfl_xor_sync(0xffffffffu, s1.y, o);
            s1.z += __shfl_xor_sync(0xffffffffu, s1.z, o), s1.w += __shfl_xor_sync(0xffffffffu, s1.w, o);
          }
          if (sub == 0 && cany) {
            const int lc = c0 + 4 * c4;  // column within the tile
            atomicAdd(stat_s + lc, s1.x);
            if (col + 1 < n_end) atomicAdd(stat_s + lc + 1, s1.y);
            if (col + 2 < n_end) atomicAdd(stat_s + lc + 2, s1.z);
            if (col + 3 < n_end) atomicAdd(stat_s + lc + 3, s1.w);
          }
        }
        if (stats) {
          // lanes sharing c4 (lane, lane^8, lane^16, lane^24) hold partial sums of the same 4 columns
#pragma unroll
          for (int o = 8; o <= 16; o <<= 1) {
            s1.x += __shfl_xor_sync(0xffffffffu, s1.x, o), s1.y += __shfl_xor_sync(0xffffffffu, s1.y, o);
            s1.z += __shfl_xor_sync(0xffffffffu, s1.z, o), s1.w += __shfl_xor_sync(0xffffffffu, s1.w, o);
            s2.x += __shfl_xor_sync(0xffffffffu, s2.x, o), s2.y += __shfl_xor_sync(0xffffffffu, s2.y, o);
            s2.z += __shfl_xor_sync(0xffffffffu, s2.z, o), s2.w += __shfl_xor_sync(0xffffffffu, s2.w, o);
          }
          if (sub == 0 && cany) {
            atomicAdd(stat_s + col, s1.x), atomicAdd(stat_s + p.n + col, s2.x);
            if (col + 1 < n_end) atomicAdd(stat_s + col + 1, s1.y), atomicAdd(stat_s + p.n + col + 1, s2.y);
            if (col + 2 < n_end) atomicAdd(stat_s + col + 2, s1.z), atomicAdd(stat_s + p.n + col + 2, s2.z);
            if (col + 3 < n_end) atomicAdd(stat_s + col + 3, s1.w), atomicAdd(stat_s + p.n + col + 3, s2.w);
          }
        }
      }
      if (!waited) {  // this warp had no chunk in the tile: still consume the phase
        mbar_wait(&acc_full[acc], acc_phase);
        tc_fence_after();
      }
      if (osum) {  // flush this tile's column sums: one global atomic per column per tile
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
        const int lc = threadIdx.x - 64;
        if (lc < p.tile_n && n0 + lc < n_end) {
          const float a = stat_s[lc];
          if (a != 0.f) atomicAdd(p.osum + n0 + lc, a * osum_alpha);
        }
        if (lc < 256) stat_s[lc] = 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      }
      // hand the accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(acc_empty_leader + 8u * acc);
        else mbar_arrive(&acc_empty[acc]);
      }
      if (++acc == p.nacc) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (stats) {
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      for (int i = threadIdx.x - 64; i < p.n; i += GEMM_THREADS - 64) {
        const float a = stat_s[i], b = stat_s[p.n + i];
        if (a != 0.f || b != 0.f) {
          atomicAdd(p.col_sum + i, a);
          atomicAdd(p.col_sumsq + i, b);
        }
      }
    }
  }

  tc_fence_before();
  __syncwarp();                     // role loops leave lane 0 behind: reconverge before the aligned barrier
  if constexpr (CG == 2) cluster_sync_all();  // the leader's MMAs read the peer's shared memory / TMEM until the very end
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------ host
// Cost model for (cta_group, tile_n): waves of work units over the CTAs / CTA pairs, each unit costing its MMA time
// (~ tile_n per K block; a pair retires a 256-row tile in the time one CTA needs for 128 rows, and its operand
// traffic per flop is 2/3 of the single-CTA tile's) plus a fixed part for the A stream / epilogue setup.
struct TileChoice {
  int cg, tile_n;
};
static double cg2_gain() {  // relative main-loop time of the pair per 128 rows (FVIT_GEMM_CG2_GAIN to tune)
  static double g = -1.0;
  if (g < 0) {
    const char* e = getenv("FVIT_GEMM_CG2_GAIN");
    g = e ? atof(e) : 0.8;
  }
  return g;
}
static int forced_cg() {  // FVIT_GEMM_CG=1|2 forces the mode where legal (A/B experiments)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FVIT_GEMM_CG");
    v = e ? atoi(e) : 0;
  }
  return v;
}
static int cost_model() {  // FVIT_GEMM_COST=0: the round-1 wave model (A/B); default: the clock model below
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FVIT_GEMM_COST");
    v = e ? atoi(e) : 0;
  }
  return v;
}
// (measured r02k, fv4 / fv0 fwd+bwd and fv0 / ar0 forward: the wave model below stays the default -- the clock model
// wins only on the N <= 64 stem / level-0 GEMMs, by not pairing CTAs there, which the wave model now does too, and
// loses 0.3 ms each on the qkv projection (picks 192 instead of 256) and the fc2 weight gradient (pairs M = 784))
// Clock model of one work unit (SM clocks): per 64-deep K block the tensor pipe needs ~2*tile_n clocks for 128 rows
// (4096 MAC/clk/SM), the operand stream needs bytes_per_CTA * concurrently_streaming_CTAs / 6300 B/clk (chip-wide
// L2 -> SM delivery, B300_MICROARCH.md "LTS throughput cap"; the measured ceiling of this kernel's main loop), and the
// epilogue ~1500 + 12 * tile_n clocks per tile, overlapped with the next unit's main loop except for the last unit.
// Narrow tiles on all 148 SMs are L2-bound (a 128 x 96 tile streams 140 B/clk per SM): fewer, wider tiles win for the
// small-M GEMMs of the carrier branch.
static double unit_clocks(int cg, int bn, int num_kb, long long active_ctas) {
  const double mma = (cg == 2 ? 3.6 : 2.0) * bn;  // (pair main loop measured: 1250 TF/s on 8192^3 = 3.7 * tile_n clocks per K block)
  const double bytes = (128.0 + (double)bn / cg) * 128.0;
  const double l2 = bytes * (double)active_ctas / 6300.0;
  return num_kb * (mma > l2 ? mma : l2);
}
static TileChoice pick_tile_clk(int m, int n, int split_k, int num_kb, int sms, int want_cg, int want_tile_n) {
  TileChoice best{1, 16};
  double best_cost = 1e30;
  for (int cg = 1; cg <= 2; ++cg) {
    if (want_cg && cg != want_cg) continue;
    if (cg == 2 && m <= BM) continue;
    const int tiles_m = ceil_div(m, BM * cg);
    const int step = 16 * cg;
    for (int bn = step; bn <= 256; bn += step) {
      if (want_tile_n > 0 && bn != want_tile_n) continue;
      const int tiles_n = ceil_div(n, bn);
      const long long work = (long long)tiles_m * tiles_n * split_k;
      const int slots = sms / cg;
      const long long waves = (work + slots - 1) / slots;
      const long long active = (work < slots ? work : slots) * cg;
      const int kb = num_kb / (split_k > 0 ? split_k : 1) > 0 ? num_kb / (split_k > 0 ? split_k : 1) : 1;
      const double main = unit_clocks(cg, bn, kb, active);
      const double epi = 1500.0 + 12.0 * bn;
      const double cost = (double)waves * (main > epi ? main : epi) + epi + 2500.0;
      if (cost < best_cost - 1e-9) {
        best_cost = cost;
        best = TileChoice{cg, bn};
      }
    }
  }
  return best;
}
static TileChoice pick_tile(int m, int n, int split_k, int sms, int want_cg, int want_tile_n) {
  TileChoice best{1, 16};
  double best_cost = 1e30;
  for (int cg = 1; cg <= 2; ++cg) {
    if (want_cg && cg != want_cg) continue;
    if (cg == 2 && m <= BM) continue;  // the second CTA of the pair would only see padding rows
    if (cg == 2 && n <= 64 && !want_cg) continue;  // nothing to share: B is a few KB, the pair only adds its barriers
    const int tiles_m = ceil_div(m, BM * cg);
    const int step = 16 * cg;  // each CTA of a pair stages tile_n / 2 columns of B: keep that a multiple of 16
    for (int bn = step; bn <= 256; bn += step) {
      if (want_tile_n > 0 && bn != want_tile_n) continue;
      const int tiles_n = ceil_div(n, bn);
      const long long work = (long long)tiles_m * tiles_n * split_k;
      const int slots = sms / cg;
      const long long waves = (work + slots - 1) / slots;
      const double cost = (double)waves * (bn * (cg == 2 ? cg2_gain() : 1.0) + 48.0);
      if (cost < best_cost - 1e-9) {
        best_cost = cost;
        best = TileChoice{cg, bn};
      }
    }
  }
  return best;
}

// 3-D (or degenerate) 16-bit tensor map with a {b0, b1, 1} box, cached by geometry.
static int get_tmap(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                    uint64_t s1_bytes, uint64_t s2_bytes, uint32_t b0, uint32_t b1, int rank) {
  uint64_t dims[3] = {d0, d1, d2};
  uint64_t strides[2] = {s1_bytes, s2_bytes};
  uint32_t box[3] = {b0, b1, 1};
  return cached_tmap_16bit(out, base, rank, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace fvit

extern "C" int fvit_gemm(const fvit_gemm_args* a, void* stream) {
  using namespace fvit;
  FVIT_CHECK(a != nullptr, "fvit_gemm: null args");
  FVIT_CHECK(a->m > 0 && a->n > 0 && a->kc > 0, "fvit_gemm: bad dims m=%d n=%d kc=%d", a->m, a->n,
             a->kc);
  FVIT_CHECK(a->ntaps >= 1 && a->ntaps <= 16, "fvit_gemm: ntaps=%d out of range", a->ntaps);
  FVIT_CHECK(a->a && a->b, "fvit_gemm: null operand");
  FVIT_CHECK(a->lda % 8 == 0 && a->ldb % 8 == 0, "fvit_gemm: lda/ldb must be multiples of 8");
  FVIT_CHECK(!(a->a_mn_major && a->ntaps != 1), "fvit_gemm: taps need a K-major A operand");
  FVIT_CHECK(a->out_f32 || a->out_f16 || a->out_pre16, "fvit_gemm: no output");
  const int b_ntaps = a->b_ntaps > 1 ? a->b_ntaps : 1;
  if (b_ntaps > 1)
    FVIT_CHECK(b_ntaps <= 16 && a->a_mn_major && a->b_mn_major && a->ntaps == 1 && a->n % 4 == 0 && !a->col_sum,
               "fvit_gemm: b_ntaps needs MN-major A and B, ntaps == 1, n %% 4 == 0");
  const int split_k = a->split_k > 1 ? a->split_k : 1;
  if (split_k > 1)
    FVIT_CHECK(a->out_f32 && !a->out_f16 && !a->col_sum, "fvit_gemm: split_k needs out_f32 only");
  if (a->act == FVIT_ACT_GELU_BWD || a->act == FVIT_ACT_RELU_BWD || a->act == FVIT_ACT_MUL_AUX)
    FVIT_CHECK(a->aux != nullptr, "fvit_gemm: backward activation needs aux");
  FVIT_CHECK(a->act >= FVIT_ACT_NONE && a->act <= FVIT_ACT_MUL_AUX, "fvit_gemm: unknown activation code %d", a->act);
  FVIT_CHECK((a->col_sum == nullptr) == (a->col_sumsq == nullptr),
             "fvit_gemm: col_sum and col_sumsq go together");
  FVIT_CHECK(!a->col_sum || a->n <= STATS_MAX_N, "fvit_gemm: statistics support n <= %d", STATS_MAX_N);
  FVIT_CHECK((a->aux_scale == nullptr) == (a->aux_shift == nullptr), "fvit_gemm: aux_scale and aux_shift go together");
  FVIT_CHECK(!a->out_colsum || (a->out_f16 && !a->col_sum && split_k == 1 && b_ntaps == 1),
             "fvit_gemm: out_colsum needs out_f16, no BN statistics, no split-K");

  const int sms = num_sms();
  int want_cg = a->cta_group == 1 || a->cta_group == 2 ? a->cta_group : forced_cg();
  if (want_cg == 2 && (a->m <= BM || (a->tile_n > 0 && a->tile_n % 32 != 0))) want_cg = 1;
  const int num_kb_all = ceil_div(a->kc, BK) * a->ntaps;
  const TileChoice tc = cost_model()
                            ? pick_tile_clk(a->m, a->n, split_k * b_ntaps, num_kb_all * b_ntaps, sms, want_cg,
                                            a->tile_n > 0 ? a->tile_n : 0)
                            : pick_tile(a->m, a->n, split_k * b_ntaps, sms, want_cg, a->tile_n > 0 ? a->tile_n : 0);
  const int tile_n = tc.tile_n, cg = tc.cg;
  FVIT_CHECK(tile_n >= 16 && tile_n <= 256 && tile_n % (16 * cg) == 0, "fvit_gemm: tile_n=%d invalid (cta_group %d)",
             tile_n, cg);

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.m = a->m;
  p.n = a->n;
  p.kb_per_tap = ceil_div(a->kc, BK);
  p.num_kb = p.kb_per_tap * a->ntaps;
  {
    static int kskip = -1;  // FVIT_GEMM_KSKIP=0: issue the zero-filled K steps too (A/B)
    if (kskip < 0) {
      const char* e = getenv("FVIT_GEMM_KSKIP");
      kskip = e ? atoi(e) : 1;
    }
    const int tail = a->kc - (p.kb_per_tap - 1) * BK;  // 1..64 real columns in the last block
    p.k_tail16 = kskip ? ceil_div(tail, UMMA_K) : BK / UMMA_K;
  }
  p.a_mn = a->a_mn_major ? 1 : 0;
  p.b_mn = a->b_mn_major ? 1 : 0;
  p.a_row_off = a->a_row_off;
  p.b_row_off = a->b_row_off;
  p.tile_n = tile_n;
  p.cg = cg;
  p.tiles_m = ceil_div(a->m, BM * cg);
  p.tiles_n = ceil_div(a->n, tile_n);
  p.b_ntaps = a->b_ntaps > 1 ? a->b_ntaps : 1;
  p.tiles_per_tap = p.tiles_n;
  p.tiles_n *= p.b_ntaps;
  p.split_k = split_k < p.num_kb ? split_k : p.num_kb;
  if (p.split_k < 1) p.split_k = 1;
  p.atomic_out = split_k > 1 ? 1 : 0;
  const int b_cols = tile_n / cg;  // B columns staged per CTA
  const int stage_bytes = A_STAGE_BYTES + (a->b_mn_major ? ((b_cols + 63) & ~63) : b_cols) * BK * 2;
  const int stats_bytes = a->col_sum ? SMEM_STATS_BYTES : (a->out_colsum ? 256 * 4 : 0);
  int stages = (SMEM_BUDGET - SMEM_CTRL_BYTES - SMEM_ALIGN_SLACK - SMEM_STG_BYTES - stats_bytes) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  FVIT_CHECK(stages >= 2, "fvit_gemm: not enough shared memory for 2 stages");
  p.stages = stages;
  p.acc_stride = tile_n <= 64 ? 64 : (tile_n <= 128 ? 128 : 256);
  p.nacc = TMEM_COLS / p.acc_stride;
  p.idesc = make_idesc_f16(BM * cg, tile_n, p.a_mn, p.b_mn, a->bf16 ? 1u : 0u);
  for (int i = 0; i < 16; ++i) {
    p.tap_shift[i] = i < (b_ntaps > 1 ? b_ntaps : a->ntaps) ? a->tap_shift[i] : 0;
    p.tap_plane[i] = i < a->ntaps ? a->tap_plane[i] : 0;
  }
  p.alpha = a->alpha;
  p.act = a->act;
  p.bf16 = a->bf16 ? 1 : 0;
  p.col_scale = a->col_scale;
  p.col_shift = a->col_shift;
  p.col_scale2 = a->col_scale2;
  p.aux = a->aux;
  p.ld_aux = a->ld_aux;
  p.resid = a->resid;
  p.ld_resid = a->ld_resid;
  p.row_map = a->row_map;
  p.out_f32 = a->out_f32;
  p.ld_o32 = a->ld_out_f32;
  p.out_f16 = a->out_f16;
  p.ld_o16 = a->ld_out_f16;
  p.col_sum = a->col_sum;
  p.col_sumsq = a->col_sumsq;
  p.alpha_ptr = a->alpha_ptr;
  p.row_scale = a->row_scale;
  p.out_pre16 = a->out_pre16;
  p.ld_pre16 = a->ld_out_pre16;
  p.aux_scale = a->aux_scale;
  p.aux_shift = a->aux_shift;
  p.osum = a->out_colsum;
  p.osum_alpha = a->out_colsum_alpha;
  p.pre_is_grad = (a->pre_is_grad && a->out_pre16 && a->act == FVIT_ACT_GELU) ? 1 : 0;
  // vector path: every touched row segment must be 16-byte aligned
  bool vec = true;
  if (a->out_f32)
    vec = vec && (a->ld_out_f32 % 4 == 0) && ((reinterpret_cast<uintptr_t>(a->out_f32) & 15) == 0);
  if (a->out_f16)
    vec = vec && (a->ld_out_f16 % 8 == 0) && ((reinterpret_cast<uintptr_t>(a->out_f16) & 15) == 0);
  if (a->resid)
    vec = vec && (a->ld_resid % 4 == 0) && ((reinterpret_cast<uintptr_t>(a->resid) & 15) == 0);
  if (a->out_pre16)
    vec = vec && (a->ld_out_pre16 % 4 == 0) && ((reinterpret_cast<uintptr_t>(a->out_pre16) & 7) == 0);
  if (a->aux)
    vec = vec && (a->ld_aux % 4 == 0) && ((reinterpret_cast<uintptr_t>(a->aux) & 7) == 0);
  p.vec_ok = vec ? 1 : 0;
  CUtensorMap tma, tmb;
  int rc;
  if (!p.a_mn) {
    const int planes = a->a_planes > 0 ? a->a_planes : 1;
    rc = get_tmap(&tma, a->a, (uint64_t)a->kc, (uint64_t)a->a_rows, (uint64_t)planes,
                  (uint64_t)a->lda * 2,
                  (uint64_t)(planes > 1 ? a->a_plane_stride : a->a_rows * a->lda) * 2, BK, BM, 3);
  } else {
    rc = get_tmap(&tma, a->a, (uint64_t)a->m, (uint64_t)a->a_rows, 1, (uint64_t)a->lda * 2,
                  (uint64_t)a->a_rows * a->lda * 2, 64, BK, 3);
  }
  if (rc) return rc;
  if (!p.b_mn) {
    const uint64_t kdim = a->ntaps == 1 ? (uint64_t)a->kc : (uint64_t)p.num_kb * BK;
    rc = get_tmap(&tmb, a->b, kdim, (uint64_t)a->n, 1, (uint64_t)a->ldb * 2, 0, BK,
                  (uint32_t)b_cols, 2);
  } else {
    rc = get_tmap(&tmb, a->b, (uint64_t)a->n, (uint64_t)a->b_rows, 1, (uint64_t)a->ldb * 2, 0, 64,
                  BK, 2);
  }
  if (rc) return rc;

  const int smem_bytes = stages * stage_bytes + SMEM_CTRL_BYTES + SMEM_ALIGN_SLACK + SMEM_STG_BYTES + stats_bytes;
  const long long work = (long long)p.tiles_m * p.tiles_n * p.split_k;
  const int slots = sms / cg;
  const int grid = (int)(work < slots ? work : slots) * cg;
  // feature mask of this call; launch the matching specialisation if one was instantiated
  uint32_t feat = ((uint32_t)a->act << EF_ACT_SHIFT);
  if (a->resid && !p.atomic_out) feat |= EF_RESID;
  if (a->out_f32) feat |= EF_O32;
  if (a->out_f16) feat |= EF_O16;
  if (a->col_sum) feat |= EF_STATS;
  if (a->out_pre16) feat |= EF_PRE;
  if (p.atomic_out) feat |= EF_ATOMIC;
  if (a->alpha_ptr) feat |= EF_ALPHAPTR;
  if (a->col_scale2) feat |= EF_CS2;
  if (a->row_scale) feat |= EF_RS;
  if (a->out_colsum) feat |= EF_OSUM;
  if (p.pre_is_grad) feat |= EF_PREGRAD;
  {
    // row-per-thread epilogue (EF_DIRECT): 16-bit outputs only, everything a lane touches in 32-byte pieces
    static int direct = -1;  // FVIT_GEMM_DIRECT=0: staging-tile epilogue everywhere (A/B)
    if (direct < 0) {
      const char* e = getenv("FVIT_GEMM_DIRECT");
      direct = e ? atoi(e) : 1;
    }
    auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; };
    const uint32_t allowed = EF_O16 | EF_PRE | EF_PREGRAD | EF_ALPHAPTR | EF_OSUM | (7u << EF_ACT_SHIFT);
    const bool ok = direct && (feat & ~allowed) == 0 && (feat & EF_O16) && b_ntaps == 1 && a->n % 16 == 0 &&
                    a->ld_out_f16 % 16 == 0 && al32(a->out_f16) &&
                    (!a->out_pre16 || (a->ld_out_pre16 % 16 == 0 && al32(a->out_pre16))) &&
                    (!a->aux || (a->ld_aux % 16 == 0 && al32(a->aux))) &&
                    (!a->col_scale || (reinterpret_cast<uintptr_t>(a->col_scale) & 15) == 0) &&
                    (!a->col_shift || (reinterpret_cast<uintptr_t>(a->col_shift) & 15) == 0) &&
                    (!a->aux_scale || ((reinterpret_cast<uintptr_t>(a->aux_scale) & 15) == 0 &&
                                       (reinterpret_cast<uintptr_t>(a->aux_shift) & 15) == 0));
    if (ok) {
      switch (feat) {   // the combinations with a direct specialisation (launch table below)
        case (0u << EF_ACT_SHIFT) | EF_O16:
        case (1u << EF_ACT_SHIFT) | EF_O16:
        case (2u << EF_ACT_SHIFT) | EF_O16:
        case (2u << EF_ACT_SHIFT) | EF_O16 | EF_PRE | EF_PREGRAD:
        case (0u << EF_ACT_SHIFT) | EF_O16 | EF_ALPHAPTR:
        case (3u << EF_ACT_SHIFT) | EF_O16:
        case (5u << EF_ACT_SHIFT) | EF_O16 | EF_ALPHAPTR:
        case (5u << EF_ACT_SHIFT) | EF_O16 | EF_ALPHAPTR | EF_OSUM:
          feat |= EF_DIRECT;
          break;
        default:
          break;
      }
    }
  }
  const bool needs_generic = false;
#define FVIT_GEMM_LAUNCH(F)                                                                            \
  do {                                                                                                 \
    if (cg == 1) {                                                                                     \
      auto kfn = gemm_tcgen05_kernel<(F), 1>;                                                          \
      static bool attr_set = false;                                                                    \
      if (!attr_set) {                                                                                 \
        FVIT_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET)); \
        attr_set = true;                                                                               \
      }                                                                                                \
      kfn<<<grid, GEMM_THREADS, smem_bytes, (cudaStream_t)stream>>>(tma, tmb, p);                      \
      return post_launch("gemm_tcgen05_kernel");                                                       \
    }                                                                                                  \
    auto kfn2 = gemm_tcgen05_kernel<(F), 2>;                                                           \
    static bool attr_set2 = false;                                                                     \
    if (!attr_set2) {                                                                                  \
      FVIT_CUDA(cudaFuncSetAttribute(kfn2, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET)); \
      attr_set2 = true;                                                                                \
    }                                                                                                  \
    cudaLaunchConfig_t cfg;                                                                            \
    memset(&cfg, 0, sizeof(cfg));                                                                      \
    cfg.gridDim = dim3((unsigned)grid), cfg.blockDim = dim3(GEMM_THREADS);                             \
    cfg.dynamicSmemBytes = (size_t)smem_bytes, cfg.stream = (cudaStream_t)stream;                      \
    cudaLaunchAttribute attr[1];                                                                       \
    attr[0].id = cudaLaunchAttributeClusterDimension;                                                  \
    attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;          \
    cfg.attrs = attr, cfg.numAttrs = 1;                                                                \
    cudaError_t le_ = cudaLaunchKernelEx(&cfg, kfn2, tma, tmb, p);                                     \
    if (le_ != cudaSuccess) {                                                                          \
      (void)cudaGetLastError(); /* do not leave the error behind for the next launch */                \
      return set_error("launch of the CTA-pair gemm_tcgen05_kernel (grid %d, %d B smem) failed: %s", grid, smem_bytes, \
                       cudaGetErrorString(le_));                                                       \
    }                                                                                                  \
    return post_launch("gemm_tcgen05_kernel");                                                         \
  } while (0)
#define FVIT_ACTF(x) ((uint32_t)(x) << EF_ACT_SHIFT)
  if (!needs_generic) {
    switch (feat) {
      // ---- inference / shared forward shapes
      case FVIT_ACTF(0) | EF_O16: FVIT_GEMM_LAUNCH(FVIT_ACTF(0) | EF_O16);                               // qkv
      case FVIT_ACTF(2) | EF_O16: FVIT_GEMM_LAUNCH(FVIT_ACTF(2) | EF_O16);                               // fc1, conv1 (eval)
      case FVIT_ACTF(1) | EF_O16: FVIT_GEMM_LAUNCH(FVIT_ACTF(1) | EF_O16);                               // stem conv1 (eval)
      case FVIT_ACTF(1) | EF_O32 | EF_O16: FVIT_GEMM_LAUNCH(FVIT_ACTF(1) | EF_O32 | EF_O16);             // stem conv2 (eval)
      case FVIT_ACTF(0) | EF_RESID | EF_O32: FVIT_GEMM_LAUNCH(FVIT_ACTF(0) | EF_RESID | EF_O32);         // proj, fc2, conv dgrad
      case FVIT_ACTF(0) | EF_RESID | EF_O32 | EF_O16: FVIT_GEMM_LAUNCH(FVIT_ACTF(0) | EF_RESID | EF_O32 | EF_O16);  // conv2 (eval)
      case FVIT_ACTF(0) | EF_O32: FVIT_GEMM_LAUNCH(FVIT_ACTF(0) | EF_O32);                               // downsample, head
      case FVIT_ACTF(0) | EF_O32 | EF_O16: FVIT_GEMM_LAUNCH(FVIT_ACTF(0) | EF_O32 | EF_O16);             // downsample -> next level
      // ---- training forward
      case FVIT_ACTF(0) | EF_O16 | EF_STATS: FVIT_GEMM_LAUNCH(FVIT_ACTF(0) | EF_O16 | EF_STATS);         // raw conv + BN statistics
      case FVIT_ACTF(2) | EF_O16 | EF_PRE: FVIT_GEMM_LAUNCH(FVIT_ACTF(2) | EF_O16 | EF_PRE);             // fc1 saving the pre-GELU value
      case FVIT_ACTF(2) | EF_O16 | EF_PRE | EF_PREGRAD:
        FVIT_GEMM_LAUNCH(FVIT_ACTF(2) | EF_O16 | EF_PRE | EF_PREGRAD);                                    // fc1 saving gelu'
      case FVIT_ACTF(0) | EF_RESID | EF_O32 | EF_RS: FVIT_GEMM_LAUNCH(FVIT_ACTF(0) | EF_RESID | EF_O32 | EF_RS);  // branch + stochastic depth
      case FVIT_ACTF(0) | EF_RESID | EF_O32 | EF_CS2 | EF_PRE:
        FVIT_GEMM_LAUNCH(FVIT_ACTF(0) | EF_RESID | EF_O32 | EF_CS2 | EF_PRE);                             // branch with layer scale
      case FVIT_ACTF(0) | EF_RESID | EF_O32 | EF_CS2 | EF_PRE | EF_RS:
        FVIT_GEMM_LAUNCH(FVIT_ACTF(0) | EF_RESID | EF_O32 | EF_CS2 | EF_PRE | EF_RS);                     // layer scale + stochastic depth
      // ---- backward
      case FVIT_ACTF(0) | EF_O32 | EF_ATOMIC | EF_ALPHAPTR: FVIT_GEMM_LAUNCH(FVIT_ACTF(0) | EF_O32 | EF_ATOMIC | EF_ALPHAPTR);  // wgrad split-K
      case FVIT_ACTF(0) | EF_O32 | EF_ALPHAPTR: FVIT_GEMM_LAUNCH(FVIT_ACTF(0) | EF_O32 | EF_ALPHAPTR);   // wgrad single pass
      case FVIT_ACTF(0) | EF_O16 | EF_ALPHAPTR: FVIT_GEMM_LAUNCH(FVIT_ACTF(0) | EF_O16 | EF_ALPHAPTR);   // dgrad
      case FVIT_ACTF(3) | EF_O16 | EF_ALPHAPTR: FVIT_GEMM_LAUNCH(FVIT_ACTF(3) | EF_O16 | EF_ALPHAPTR);   // dgrad * gelu'
      case FVIT_ACTF(3) | EF_O16: FVIT_GEMM_LAUNCH(FVIT_ACTF(3) | EF_O16);                               // conv dgrad * gelu'
      case FVIT_ACTF(3) | EF_O16 | EF_ALPHAPTR | EF_OSUM:
        FVIT_GEMM_LAUNCH(FVIT_ACTF(3) | EF_O16 | EF_ALPHAPTR | EF_OSUM);                                  // fc2 dgrad * gelu' + fc1 bias gradient
      case FVIT_ACTF(5) | EF_O16 | EF_ALPHAPTR: FVIT_GEMM_LAUNCH(FVIT_ACTF(5) | EF_O16 | EF_ALPHAPTR);   // fc2 dgrad * saved gelu'
      case FVIT_ACTF(5) | EF_O16 | EF_ALPHAPTR | EF_OSUM:
        FVIT_GEMM_LAUNCH(FVIT_ACTF(5) | EF_O16 | EF_ALPHAPTR | EF_OSUM);                                  // ... + fc1 bias gradient
      // ---- row-per-thread epilogue variants of the 16-bit-output launches
      case EF_DIRECT | FVIT_ACTF(0) | EF_O16: FVIT_GEMM_LAUNCH(EF_DIRECT | FVIT_ACTF(0) | EF_O16);
      case EF_DIRECT | FVIT_ACTF(1) | EF_O16: FVIT_GEMM_LAUNCH(EF_DIRECT | FVIT_ACTF(1) | EF_O16);
      case EF_DIRECT | FVIT_ACTF(2) | EF_O16: FVIT_GEMM_LAUNCH(EF_DIRECT | FVIT_ACTF(2) | EF_O16);
      case EF_DIRECT | FVIT_ACTF(2) | EF_O16 | EF_PRE | EF_PREGRAD:
        FVIT_GEMM_LAUNCH(EF_DIRECT | FVIT_ACTF(2) | EF_O16 | EF_PRE | EF_PREGRAD);
      case EF_DIRECT | FVIT_ACTF(0) | EF_O16 | EF_ALPHAPTR: FVIT_GEMM_LAUNCH(EF_DIRECT | FVIT_ACTF(0) | EF_O16 | EF_ALPHAPTR);
      case EF_DIRECT | FVIT_ACTF(3) | EF_O16: FVIT_GEMM_LAUNCH(EF_DIRECT | FVIT_ACTF(3) | EF_O16);
      case EF_DIRECT | FVIT_ACTF(5) | EF_O16 | EF_ALPHAPTR: FVIT_GEMM_LAUNCH(EF_DIRECT | FVIT_ACTF(5) | EF_O16 | EF_ALPHAPTR);
      case EF_DIRECT | FVIT_ACTF(5) | EF_O16 | EF_ALPHAPTR | EF_OSUM:
        FVIT_GEMM_LAUNCH(EF_DIRECT | FVIT_ACTF(5) | EF_O16 | EF_ALPHAPTR | EF_OSUM);
      default: break;
    }
  }
  FVIT_GEMM_LAUNCH(EF_GENERIC);
#undef FVIT_ACTF
#undef FVIT_GEMM_LAUNCH
}
