// fvit_hat_attn_fwd: the fused hierarchical-attention kernel of WindowAttention.forward (fv.py:557-565) for sm_100a —
// QKV projection, softmax(Q K^T * scale + relative-position bias) and the V contraction in ONE tcgen05 kernel; the
// [tokens, 3C] qkv matrix and the score / probability matrices never exist in HBM (inference), or qkv is written once
// and never read back (training, where the backward pass needs it).
//
// Work item = (tile of 128 token rows = gpt whole windows in slots of 16/32/64/128 rows, head); head-major order so the
// head's weight slice [3*hdp, C] and bias [S, S] stay hot in L2 / shared memory.
//
//   warp 0     : TMA producer — per 64-wide K block the (window_size^2 + ct_size^2) x 64 activation boxes of the tile's
//                windows and the head's W_q / W_k / W_v boxes into a ring of 128B-swizzled stages
//   warp 1     : projection MMA issuer — [Q | K | V](128 x 3*hdp, fp32 in TMEM, double buffered) = X W_h^T over C
//   warp 2     : attention MMA issuer — S = Q K^T (128 x 128) and O = P V from the fp16 operand tiles the epilogue
//                warps wrote to shared memory
//   warps 3..6 : epilogue — one thread per token row: (1) Q, K, V accumulators + qkv bias -> fp16 -> swizzled UMMA
//                operand tiles (and optionally the qkv matrix for the backward pass); (2) softmax over the window's keys
//                (block-diagonal mask, bias from shared memory, exp2), P as fp16 operand; (3) O / rowsum -> fp16 out.
// The projection of item i+1 runs on the tensor core while the epilogue warps are busy with item i.
#include <cuda_fp16.h>

#include "../../include/fvit.h"
#include "common.h"
#include "ptx.cuh"

namespace fvit {

constexpr int HT_THREADS = 224;
constexpr int HT_ROWS = 128;
constexpr int HT_BK = 64;
constexpr int HT_MAX_STAGES = 4;

struct HatParams {
  int groups, S, heads, gpt, tiles, slot;
  int num_kb;   // ceil(C / 64)
  int k_tail16; // 16-wide MMA steps with real columns in the last K-block (the rest is TMA zero fill: not issued)
  int stages;
  int Cp;       // heads * hdp
  float scale_log2e;
  const float* qkv_bias;  // [3 * Cp] head-padded, or null
  const float* bias;      // [heads, S, S] or null
  __half* out;
  long long ldo;
  __half* qkv_out;        // optional [rows, 3 * Cp]
  long long ldq;
  int rows_total;
};

__device__ __forceinline__ void ht_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int HDP>
__global__ void __launch_bounds__(HT_THREADS, 1)
    hat_attn_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ HatParams p) {
  constexpr uint32_t SWZ = HDP == 64 ? SWZ_128B : SWZ_64B;
  constexpr uint32_t ROW_BYTES = HDP * 2;
  constexpr uint32_t SBO_QKV = 8 * ROW_BYTES;
  constexpr int A_BYTES = HT_ROWS * HT_BK * 2;        // 16 KB
  constexpr int B_BYTES = 3 * HDP * HT_BK * 2;        // 24 KB / 12 KB
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int OPER_BYTES = HT_ROWS * HDP * 2;       // one of Q / K / V as fp16 operand tile
  constexpr int NQKV = 3 * HDP;                       // accumulator columns per stage
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S;
  uint8_t* s_oper = smem + p.stages * STAGE_BYTES;         // Q, K, V operand tiles
  uint8_t* s_p = s_oper + 3 * OPER_BYTES;                  // P: 128 x 128 fp16 (two 64-wide K atoms)
  float* bias_s = reinterpret_cast<float*>(s_p + HT_ROWS * 128 * 2);
  uint8_t* ctrl = reinterpret_cast<uint8_t*>(bias_s) + ((S * S * 4 + 15) & ~15);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);   // [HT_MAX_STAGES]
  uint64_t* empty_bar = full_bar + HT_MAX_STAGES;           // [HT_MAX_STAGES]
  uint64_t* acc_full = empty_bar + HT_MAX_STAGES;           // [2]
  uint64_t* acc_empty = acc_full + 2;                       // [2]
  uint64_t* qkv_ready = acc_empty + 2;
  uint64_t* s_full = qkv_ready + 1;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_full = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    for (int i = 0; i < p.stages; ++i) mbar_init(&full_bar[i], 1), mbar_init(&empty_bar[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i], 4);
    mbar_init(qkv_ready, 4), mbar_init(s_full, 1), mbar_init(p_full, 4), mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc[2] = {tmem_base, tmem_base + NQKV};
  const uint32_t tmem_S = tmem_base + 384;  // scores; O overwrites its first HDP columns

  const int items = p.tiles * p.heads;
  const int nslots = HT_ROWS / p.slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x) {
        const int head = w / p.tiles, tile = w % p.tiles;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          for (int sl = 0; sl < nslots; ++sl)  // window (tile*gpt + sl) -> rows [sl*slot, (sl+1)*slot); OOB rows are zero
            tma_load_2d(sa + sl * p.slot * 128, &tmap_x, &full_bar[stage], kb * HT_BK, (tile * p.gpt + sl) * S);
#pragma unroll
          for (int which = 0; which < 3; ++which)
            tma_load_2d(sb + which * (HDP * 128), &tmap_w, &full_bar[stage], kb * HT_BK, which * p.Cp + head * HDP);
          if (++stage == p.stages) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== projection MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, NQKV, 0, 0);
      int stage = 0;
      uint32_t phase = 0, it = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
        const int a = it & 1;
        mbar_wait(&acc_empty[a], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + A_BYTES;
          const int nk = kb + 1 == p.num_kb ? p.k_tail16 : HT_BK / 16;
#pragma unroll
          for (int k = 0; k < HT_BK / 16; ++k) {
            if (k >= nk) break;
            umma_f16_ss(tmem_acc[a], make_smem_desc(sa + k * 32, 16, 1024, SWZ_128B),
                        make_smem_desc(sb + k * 32, 16, 1024, SWZ_128B), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) stage = 0, phase ^= 1;
        }
        umma_commit(&acc_full[a]);
      }
    }
  } else if (warp == 2) {
    // ===================================================================== attention MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0);
      const uint32_t idesc_o = make_idesc_f16(128, HDP, 0, 1);  // B = V, MN-major
      const uint32_t sQ = smem_u32(s_oper), sK = sQ + OPER_BYTES, sV = sK + OPER_BYTES, sP = smem_u32(s_p);
      uint32_t it = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
        mbar_wait(qkv_ready, it & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HDP / 16; ++k)
          umma_f16_ss(tmem_S, make_smem_desc(sQ + k * 32, 16, SBO_QKV, SWZ), make_smem_desc(sK + k * 32, 16, SBO_QKV, SWZ),
                      idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
        mbar_wait(p_full, it & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)  // 128 keys in steps of 16
          umma_f16_ss(tmem_S, make_smem_desc(sP + (ks >> 2) * (HT_ROWS * 128) + (ks & 3) * 32, 16, 1024, SWZ_128B),
                      make_smem_desc(sV + ks * 16 * ROW_BYTES, 0, SBO_QKV, SWZ), idesc_o, ks > 0 ? 1u : 0u);
        umma_commit(o_full);
      }
    }
  } else {
    // ===================================================================== epilogue warps
    const int quad = warp & 3;
    const int r = quad * 32 + lane;    // row in tile
    const int tid = threadIdx.x - 96;  // 0..127 among the epilogue threads
    const int slot = p.slot;
    const int sl = r / slot;
    const int j = r - sl * slot;       // token index inside the window
    const int lo = sl * slot;
    const int wlo = slot >= 32 ? lo : quad * 32;
    const int wn = slot >= 32 ? S : 32;
    const int nch = (wn + 31) / 32;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    // columns of P outside the warp's windows are never written: zero them once
    for (int c = 0; c < 128; c += 8) {
      if (c + 8 <= wlo || c >= wlo + 32 * nch)
        *reinterpret_cast<uint4*>(s_p + (c >> 6) * (HT_ROWS * 128) + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4)) =
            make_uint4(0, 0, 0, 0);
    }
    // byte offset of 16-byte chunk c8 of this thread's row inside a Q / K / V operand tile
    auto oper_off = [&](int c8) -> uint32_t {
      if (HDP == 64) return (uint32_t)(r * 128 + ((c8 ^ (r & 7)) << 4));
      return (uint32_t)(r * 64 + ((c8 ^ ((r >> 1) & 3)) << 4));
    };
    uint32_t it = 0;
    int staged_head = -1;
    for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
      const int head = w / p.tiles, tile = w % p.tiles;
      const int grp = tile * p.gpt + sl;
      const bool row_ok = j < S && sl < p.gpt && grp < p.groups;
      const long long grow = (long long)grp * S + j;
      if (p.bias && head != staged_head) {
        ht_bar_sync(1, 128);
        const float* bsrc = p.bias + (long long)head * S * S;
        const int n = S * S;
        for (int i0 = tid; i0 < n; i0 += 128 * 8) {
          float tmp[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * 128;
            tmp[u] = i < n ? __ldg(bsrc + i) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * 128;
            if (i < n) bias_s[i] = tmp[u] * 1.4426950408889634f;
          }
        }
        ht_bar_sync(1, 128);
        staged_head = head;
      }
      // ---- (1) projection accumulators -> fp16 operand tiles (+ qkv bias, + optional global copy)
      const int a = it & 1;
      mbar_wait(&acc_full[a], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int which = 0; which < 3; ++which) {
        uint8_t* dst = s_oper + which * OPER_BYTES;
        const float* qb = p.qkv_bias ? p.qkv_bias + which * p.Cp + head * HDP : nullptr;
        __half* grow_ptr = (p.qkv_out && row_ok) ? p.qkv_out + grow * p.ldq + which * p.Cp + head * HDP : nullptr;
#pragma unroll
        for (int c0 = 0; c0 < HDP; c0 += 32) {
          uint32_t raw[32];
          tmem_ld32(tmem_acc[a] + lane_off + which * HDP + c0, raw);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t o4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int c = q * 8 + 2 * u;
              const float b0 = qb ? __ldg(qb + c0 + c) : 0.f, b1 = qb ? __ldg(qb + c0 + c + 1) : 0.f;
              const __half2 hh = __floats2half2_rn(__uint_as_float(raw[c]) + b0, __uint_as_float(raw[c + 1]) + b1);
              o4[u] = *reinterpret_cast<const uint32_t*>(&hh);
            }
            const uint4 v4 = make_uint4(o4[0], o4[1], o4[2], o4[3]);
            *reinterpret_cast<uint4*>(dst + oper_off((c0 >> 3) + q)) = v4;
            if (grow_ptr) *reinterpret_cast<uint4*>(grow_ptr + c0 + q * 8) = v4;
          }
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&acc_empty[a]);
        mbar_arrive(qkv_ready);
      }
      // ---- (2) softmax over the window's keys
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      const uint32_t ts = tmem_S + lane_off;
      const float* brow = bias_s + (row_ok ? j : 0) * S;
      const bool use_bias = p.bias != nullptr;
      float mx = -INFINITY;
      for (int ch = 0; ch < nch; ++ch) {
        const int c0 = wlo + 32 * ch;
        uint32_t raw[32];
        tmem_ld32(ts + c0, raw);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          const int kk = c0 + u - lo;
          const bool in = kk >= 0 && kk < S;
          const float b = use_bias ? brow[in ? kk : 0] : 0.f;
          const float sc = fmaf(__uint_as_float(raw[u]), p.scale_log2e, b);
          mx = fmaxf(mx, in ? sc : -INFINITY);
        }
      }
      if (!row_ok) mx = 0.f;
      float sum = 0.f;
      for (int ch = 0; ch < nch; ++ch) {
        const int c0 = wlo + 32 * ch;
        uint32_t raw[32];
        tmem_ld32(ts + c0, raw);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int u = 0; u < 32; u += 2) {
          float e[2];
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int kk = c0 + u + t - lo;
            const bool in = row_ok && kk >= 0 && kk < S;
            const float b = use_bias ? brow[in ? kk : 0] : 0.f;
            const float sc = fmaf(__uint_as_float(raw[u + t]), p.scale_log2e, b - mx);
            e[t] = in ? exp2f(sc) : 0.f;
          }
          const __half2 h = __floats2half2_rn(e[0], e[1]);
          sum += __low2float(h) + __high2float(h);
          pk[u >> 1] = *reinterpret_cast<const uint32_t*>(&h);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = c0 + q * 8;
          if (c < 128)
            *reinterpret_cast<uint4*>(s_p + (c >> 6) * (HT_ROWS * 128) + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4)) =
                make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // ---- (3) O epilogue
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      const float inv = row_ok ? 1.f / sum : 0.f;
      __half* orow = p.out + grow * p.ldo + head * HDP;
#pragma unroll
      for (int c0 = 0; c0 < HDP; c0 += 32) {
        uint32_t raw[32];
        tmem_ld32(ts + c0, raw);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t o4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const __half2 h = __floats2half2_rn(__uint_as_float(raw[q * 8 + 2 * u]) * inv,
                                                  __uint_as_float(raw[q * 8 + 2 * u + 1]) * inv);
              o4[u] = *reinterpret_cast<const uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(orow + c0 + q * 8) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
          }
        }
      }
      tc_fence_before();  // the next item's score MMA may overwrite the stage once qkv_ready is signalled again
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int HDP>
static int launch_hat(const CUtensorMap& tx, const CUtensorMap& tw, HatParams& p, cudaStream_t st) {
  constexpr int STAGE_BYTES = HT_ROWS * HT_BK * 2 + 3 * HDP * HT_BK * 2;
  const size_t fixed = 1024 + (size_t)3 * HT_ROWS * HDP * 2 + HT_ROWS * 128 * 2 + (((size_t)p.S * p.S * 4 + 15) & ~(size_t)15) + 256;
  int stages = (int)((227 * 1024 - fixed) / STAGE_BYTES);
  if (stages > HT_MAX_STAGES) stages = HT_MAX_STAGES;
  FVIT_CHECK(stages >= 2, "fvit_hat_attn_fwd: S=%d leaves no room for a 2-stage operand ring", p.S);
  p.stages = stages;
  const size_t smem = fixed + (size_t)stages * STAGE_BYTES;
  static bool configured = false;
  if (!configured) {
    FVIT_CUDA(cudaFuncSetAttribute(hat_attn_kernel<HDP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int items = p.tiles * p.heads;
  const int sms = num_sms();
  hat_attn_kernel<HDP><<<items < sms ? items : sms, HT_THREADS, smem, st>>>(tx, tw, p);
  return post_launch("hat_attn_kernel");
}

}  // namespace fvit

using namespace fvit;

extern "C" int fvit_hat_attn_fwd(const void* xn16, int64_t ldx, int32_t C, const void* wqkv16, int64_t ldw,
                                 const float* qkv_bias, int32_t groups, int32_t S, int32_t heads, int32_t hdp,
                                 const float* bias, float scale, void* out, int64_t ldo, void* qkv_out, int64_t ldq,
                                 void* stream) {
  FVIT_CHECK(xn16 && wqkv16 && out && groups > 0 && heads > 0 && C > 0, "fvit_hat_attn_fwd: bad arguments");
  FVIT_CHECK(S >= 1 && S <= 128, "fvit_hat_attn_fwd: S=%d unsupported (1..128)", S);
  FVIT_CHECK(hdp == 32 || hdp == 64, "fvit_hat_attn_fwd: padded head dim %d unsupported (32 or 64)", hdp);
  FVIT_CHECK(ldx % 8 == 0 && ldw % 8 == 0 && ldo % 8 == 0 && ldo >= heads * hdp, "fvit_hat_attn_fwd: bad leading dimensions");
  FVIT_CHECK(!qkv_out || (ldq % 8 == 0 && ldq >= 3 * heads * hdp), "fvit_hat_attn_fwd: bad qkv_out leading dimension");
  FVIT_CHECK((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(qkv_out) & 15) == 0,
             "fvit_hat_attn_fwd: outputs must be 16-byte aligned");
  HatParams p;
  memset(&p, 0, sizeof(p));
  p.groups = groups, p.S = S, p.heads = heads;
  p.slot = S <= 16 ? 16 : (S <= 32 ? 32 : (S <= 64 ? 64 : 128));
  p.gpt = HT_ROWS / p.slot;
  p.tiles = (groups + p.gpt - 1) / p.gpt;
  p.num_kb = (C + HT_BK - 1) / HT_BK;
  p.k_tail16 = (C - (p.num_kb - 1) * HT_BK + 15) / 16;
  p.Cp = heads * hdp;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.qkv_bias = qkv_bias, p.bias = bias;
  p.out = (__half*)out, p.ldo = ldo, p.qkv_out = (__half*)qkv_out, p.ldq = ldq;
  p.rows_total = groups * S;
  CUtensorMap tx, tw;
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)p.rows_total};
    uint64_t strides[1] = {(uint64_t)ldx * 2};
    uint32_t box[2] = {(uint32_t)HT_BK, (uint32_t)p.slot};
    int rc = cached_tmap_16bit(&tx, xn16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    uint64_t dimsw[2] = {(uint64_t)C, (uint64_t)(3 * p.Cp)};
    uint64_t stridesw[1] = {(uint64_t)ldw * 2};
    uint32_t boxw[2] = {(uint32_t)HT_BK, (uint32_t)hdp};
    rc = cached_tmap_16bit(&tw, wqkv16, 2, dimsw, stridesw, boxw, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  if (hdp == 64) return launch_hat<64>(tx, tw, p, (cudaStream_t)stream);
  return launch_hat<32>(tx, tw, p, (cudaStream_t)stream);
}
