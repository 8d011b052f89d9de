// fvit_attn_loop_fwd: tensor-core attention core of WindowAttention.forward (fv.py:559-565; fvar.py:805-817 for the
// any-res level-2 geometry S = 12*12 + 4 = 148; fv.py:1253-1418 for the 21k windows S = 196 .. 2304) for window
// sequences longer than one 128-row tile.
//
// Work item = (window, head, tile of 128 query rows). Keys are visited in tiles of 128, twice:
//   pass 1: S_j = Q K_j^T (tcgen05, 128 x 128 fp32 in TMEM)  -> running row maximum of S_j*scale + bias
//   pass 2: S_j again, P_j = exp2(S_j*scale*log2e + bias*log2e - max) as fp16 into the swizzled operand tile,
//           O += P_j V_j (TMEM accumulator over the key tiles), row sums of the rounded probabilities
// so no accumulator is ever rescaled; the score recomputation is 2*S*S*hd FLOPs of a contraction that is < 3 % of
// the model. Scores / probabilities never touch HBM.
//
//   warp 0     : TMA producer  — Q tile, K / V tiles through a 4-stage ring, relative-position-bias tiles
//                (fp32 [128 rows x 64 keys] units, two 128B-swizzled boxes each) through a 3-stage ring
//   warp 1     : MMA issuer    — score MMAs double-buffered in TMEM so S_{j+1} is in flight while the softmax
//                warps work on S_j; O accumulates in a third TMEM region
//   warps 2..5 : softmax       — one thread per query row (tcgen05.ld, bias from shared memory, exp2, P store),
//                then the O epilogue (1 / rowsum, fp16, 16-byte stores) and the optional log-sum-exp row vector
#include <cuda_fp16.h>

#include "../../include/fvit.h"
#include "common.h"
#include "ptx.cuh"

namespace fvit {

constexpr int AL_THREADS = 192;
constexpr int AL_ROWS = 128;
constexpr int AL_NKV = 4;    // K / V tile ring depth
constexpr int AL_NB = 3;     // bias unit ring depth
constexpr int AL_BUNIT = 64; // key columns per bias unit
constexpr int AL_BUNIT_BYTES = AL_ROWS * AL_BUNIT * 4;

struct AttnLoopParams {
  int groups, S, heads, nqt, nkt;
  float scale_log2e;
  int has_bias;
  __half* out;
  long long ldo;
  float* lse;  // optional [groups * S, heads]: log2-domain log-sum-exp of the scaled, biased scores
};

// Order in which pass 2 consumes K / V tiles: K_0, then for every j: K_{j+1} (if any), V_j — the score MMA of the
// next key tile is issued before the P V MMA of the current one. Producer and MMA issuer walk the same sequence.
__device__ __forceinline__ void pass2_tile(int t, int nkt, bool& is_v, int& j) {
  if (t == 0) {
    is_v = false, j = 0;
    return;
  }
  // t >= 1: pairs (K_{j+1}, V_j) for j < nkt - 1, then the lone V_{nkt-1}
  const int pair = (t - 1) >> 1;
  if (pair < nkt - 1) {
    is_v = ((t - 1) & 1) != 0;
    j = is_v ? pair : pair + 1;
  } else {
    is_v = true, j = nkt - 1;
  }
}

template <int HDP>
__global__ void __launch_bounds__(AL_THREADS, 1)
    attn_loop_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_bias,
                     const __grid_constant__ AttnLoopParams p) {
  constexpr uint32_t SWZ = HDP == 64 ? SWZ_128B : SWZ_64B;
  constexpr uint32_t ROW_BYTES = HDP * 2;
  constexpr uint32_t SBO_QKV = 8 * ROW_BYTES;
  constexpr int TILE_BYTES = AL_ROWS * HDP * 2;
  constexpr int Q_OFF = 0;
  constexpr int KV_OFF = TILE_BYTES;
  constexpr int P_OFF = KV_OFF + AL_NKV * TILE_BYTES;
  constexpr int BIAS_OFF = P_OFF + AL_ROWS * 128 * 2;
  constexpr int CTRL_OFF = BIAS_OFF + AL_NB * AL_BUNIT_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S, nkt = p.nkt;
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + CTRL_OFF);
  uint64_t* q_empty = q_full + 1;
  uint64_t* kv_full = q_empty + 1;         // [AL_NKV]
  uint64_t* kv_empty = kv_full + AL_NKV;   // [AL_NKV]
  uint64_t* b_full = kv_empty + AL_NKV;    // [AL_NB]
  uint64_t* b_empty = b_full + AL_NB;      // [AL_NB]
  uint64_t* s_full = b_empty + AL_NB;      // [2]
  uint64_t* s_empty = s_full + 2;          // [2]
  uint64_t* p_full = s_empty + 2;
  uint64_t* p_empty = p_full + 1;
  uint64_t* o_full = p_empty + 1;
  uint64_t* o_empty = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    if (p.has_bias) tma_prefetch_desc(&tmap_bias);
    mbar_init(q_full, 1), mbar_init(q_empty, 1);
    for (int i = 0; i < AL_NKV; ++i) mbar_init(&kv_full[i], 1), mbar_init(&kv_empty[i], 1);
    for (int i = 0; i < AL_NB; ++i) mbar_init(&b_full[i], 1), mbar_init(&b_empty[i], 4);
    for (int i = 0; i < 2; ++i) mbar_init(&s_full[i], 1), mbar_init(&s_empty[i], 4);
    mbar_init(p_full, 4), mbar_init(p_empty, 1), mbar_init(o_full, 1), mbar_init(o_empty, 4);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S[2] = {tmem_base, tmem_base + 128};
  const uint32_t tmem_O = tmem_base + 256;

  const int items = p.groups * p.heads * p.nqt;
  // item -> (head, q tile) outer, window inner: CTAs running side by side share the bias tiles in L2
  auto decode = [&](int w, int& grp, int& head, int& qi) {
    grp = w % p.groups;
    const int hq = w / p.groups;
    qi = hq % p.nqt;
    head = hq / p.nqt;
  };
  auto units_of = [&](int j) { return (S - j * 128) > AL_BUNIT ? 2 : 1; };  // bias units of key tile j

  if (warp == 0) {
    if (lane == 0) {
      uint32_t kv_cnt = 0, b_cnt = 0, it = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
        int grp, head, qi;
        decode(w, grp, head, qi);
        const int row0 = grp * S;
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_expect_tx(q_full, TILE_BYTES);
        tma_load_2d(smem + Q_OFF, &tmap_qkv, q_full, head * HDP, row0 + qi * 128);
        // K / V tiles and bias units are produced in lock step with their consumers' order; the two streams are
        // merged so that neither ring starves the other: per key tile first its K (and V) boxes, then its bias
        for (int pass = 0; pass < 2; ++pass) {
          const int ntiles = pass == 0 ? nkt : 2 * nkt;
          int bias_j = 0;  // next key tile whose bias units have not been issued in this pass
          for (int t = 0; t < ntiles; ++t) {
            bool is_v = false;
            int j = t;
            if (pass == 1) pass2_tile(t, nkt, is_v, j);
            const int st = kv_cnt % AL_NKV;
            mbar_wait(&kv_empty[st], ((kv_cnt / AL_NKV) & 1) ^ 1);
            mbar_expect_tx(&kv_full[st], TILE_BYTES);
            tma_load_2d(smem + KV_OFF + st * TILE_BYTES, &tmap_qkv, &kv_full[st],
                        ((is_v ? 2 : 1) * p.heads + head) * HDP, row0 + j * 128);
            ++kv_cnt;
            if (p.has_bias && !is_v && bias_j <= j) {
              for (; bias_j <= j; ++bias_j) {
                for (int u = 0; u < units_of(bias_j); ++u) {
                  const int bs = b_cnt % AL_NB;
                  mbar_wait(&b_empty[bs], ((b_cnt / AL_NB) & 1) ^ 1);
                  mbar_expect_tx(&b_full[bs], AL_BUNIT_BYTES);
                  uint8_t* dst = smem + BIAS_OFF + bs * AL_BUNIT_BYTES;
                  const int c0 = bias_j * 128 + u * AL_BUNIT;
                  tma_load_2d(dst, &tmap_bias, &b_full[bs], c0, head * S + qi * 128);
                  tma_load_2d(dst + AL_ROWS * 128, &tmap_bias, &b_full[bs], c0 + 32, head * S + qi * 128);
                  ++b_cnt;
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0);
      const uint32_t idesc_o = make_idesc_f16(128, HDP, 0, 1);  // B = V, MN-major
      const uint32_t sQ = smem_u32(smem + Q_OFF);
      const uint32_t sP = smem_u32(smem + P_OFF);
      uint32_t kv_cnt = 0, s_cnt = 0, p_cnt = 0, it = 0;
      auto issue_s = [&]() {  // S = Q K^T from the next ring stage into the next score stage
        const int st = kv_cnt % AL_NKV;
        mbar_wait(&kv_full[st], (kv_cnt / AL_NKV) & 1);
        const int ss = s_cnt & 1;
        mbar_wait(&s_empty[ss], ((s_cnt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t sK = smem_u32(smem + KV_OFF + st * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < HDP / 16; ++k)
          umma_f16_ss(tmem_S[ss], make_smem_desc(sQ + k * 32, 16, SBO_QKV, SWZ), make_smem_desc(sK + k * 32, 16, SBO_QKV, SWZ),
                      idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&s_full[ss]);
        umma_commit(&kv_empty[st]);
        ++kv_cnt, ++s_cnt;
      };
      auto issue_pv = [&](bool first) {  // O (+)= P V from the next ring stage
        const int st = kv_cnt % AL_NKV;
        mbar_wait(&kv_full[st], (kv_cnt / AL_NKV) & 1);
        mbar_wait(p_full, p_cnt & 1);
        tc_fence_after();
        const uint32_t sV = smem_u32(smem + KV_OFF + st * TILE_BYTES);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_f16_ss(tmem_O, make_smem_desc(sP + (ks >> 2) * (AL_ROWS * 128) + (ks & 3) * 32, 16, 1024, SWZ_128B),
                      make_smem_desc(sV + ks * 16 * ROW_BYTES, 0, SBO_QKV, SWZ), idesc_o, (!first || ks > 0) ? 1u : 0u);
        umma_commit(p_empty);
        umma_commit(&kv_empty[st]);
        ++kv_cnt, ++p_cnt;
      };
      for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
        mbar_wait(q_full, it & 1);
        for (int j = 0; j < nkt; ++j) issue_s();  // pass 1: maxima only
        mbar_wait(o_empty, (it & 1) ^ 1);         // previous item's O has been read out
        issue_s();                                // pass 2: K_0
        for (int j = 0; j < nkt; ++j) {
          if (j + 1 < nkt) issue_s();
          else umma_commit(q_empty);              // last score MMA of the item issued: Q may be overwritten once it retires
          issue_pv(j == 0);
        }
        umma_commit(o_full);
      }
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // query row inside the tile
    uint8_t* sP = smem + P_OFF;
    uint32_t s_cnt = 0, b_cnt = 0, p_cnt = 0, it = 0;
    const bool use_bias = p.has_bias != 0;
    for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
      int grp, head, qi;
      decode(w, grp, head, qi);
      const int qrow = qi * 128 + r;
      const bool row_ok = qrow < S;
      float mx = -INFINITY, sum = 0.f;
      for (int pass = 0; pass < 2; ++pass) {
        for (int j = 0; j < nkt; ++j) {
          const int ss = s_cnt & 1;
          mbar_wait(&s_full[ss], (s_cnt >> 1) & 1);
          tc_fence_after();
          const uint32_t ts = tmem_S[ss] + ((uint32_t)(quad * 32) << 16);
          const int ncols = min(128, S - j * 128);  // valid keys of this tile
          if (pass == 1) mbar_wait(p_empty, (p_cnt & 1) ^ 1);  // the previous P V MMA has consumed the P tile
          for (int u = 0; u < 2; ++u) {
            const bool unit_live = u * AL_BUNIT < ncols;
            const uint8_t* bunit = nullptr;
            if (use_bias && unit_live) {
              const int bs = b_cnt % AL_NB;
              mbar_wait(&b_full[bs], (b_cnt / AL_NB) & 1);
              bunit = smem + BIAS_OFF + bs * AL_BUNIT_BYTES;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int c0 = u * AL_BUNIT + h * 32;  // first key column of this 32-wide chunk
              if (pass == 0 && c0 >= ncols) continue;
              uint32_t raw[32];
              float sc[32];
              if (c0 < ncols) {
                tmem_ld32(ts + c0, raw);
                tmem_ld_wait();
                const uint8_t* brow = bunit ? bunit + h * (AL_ROWS * 128) + r * 128 : nullptr;
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (brow) b4 = *reinterpret_cast<const float4*>(brow + ((q4 ^ (r & 7)) << 4));
                  const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const int c = q4 * 4 + e;
                    const float v = fmaf(__uint_as_float(raw[c]), p.scale_log2e, bb[e] * 1.4426950408889634f);
                    sc[c] = (c0 + c < ncols) ? v : -INFINITY;
                  }
                }
              } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) sc[c] = -INFINITY;
              }
              if (pass == 0) {
#pragma unroll
                for (int c = 0; c < 32; ++c) mx = fmaxf(mx, sc[c]);
              } else {
                uint32_t pk[16];
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                  const float e0 = row_ok ? exp2f(sc[c] - mx) : 0.f;  // exp2(-inf) = 0 for masked keys
                  const float e1 = row_ok ? exp2f(sc[c + 1] - mx) : 0.f;
                  const __half2 hh = __floats2half2_rn(e0, e1);
                  sum += __low2float(hh) + __high2float(hh);  // what the tensor core will multiply
                  pk[c >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const int c = c0 + q * 8;
                  *reinterpret_cast<uint4*>(sP + (c >> 6) * (AL_ROWS * 128) + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4)) =
                      make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                }
              }
            }
            if (use_bias && unit_live) {
              __syncwarp();
              if (lane == 0) mbar_arrive(&b_empty[b_cnt % AL_NB]);
              ++b_cnt;
            }
          }
          if (pass == 0 && !row_ok) mx = 0.f;
          if (pass == 1) {
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
            ++p_cnt;
          } else {
            tc_fence_before();
            __syncwarp();
          }
          if (lane == 0) mbar_arrive(&s_empty[ss]);
          ++s_cnt;
        }
        if (pass == 0 && !row_ok) mx = 0.f;
      }
      // O epilogue
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      const float inv = row_ok ? 1.f / sum : 0.f;
      const long long grow = (long long)grp * S + qrow;
      const uint32_t to = tmem_O + ((uint32_t)(quad * 32) << 16);
#pragma unroll
      for (int c0 = 0; c0 < HDP; c0 += 32) {
        uint32_t raw[32];
        tmem_ld32(to + c0, raw);
        tmem_ld_wait();
        if (row_ok) {
          __half* orow = p.out + grow * p.ldo + head * HDP;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t o4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const __half2 hh = __floats2half2_rn(__uint_as_float(raw[q * 8 + 2 * u]) * inv,
                                                   __uint_as_float(raw[q * 8 + 2 * u + 1]) * inv);
              o4[u] = *reinterpret_cast<const uint32_t*>(&hh);
            }
            *reinterpret_cast<uint4*>(orow + c0 + q * 8) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
          }
        }
      }
      if (p.lse && row_ok) p.lse[grow * p.heads + head] = mx + log2f(sum);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
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
static int launch_attn_loop(const CUtensorMap& tq, const CUtensorMap& tb, const AttnLoopParams& p, cudaStream_t st) {
  constexpr size_t smem = 1024 + (size_t)(1 + AL_NKV) * AL_ROWS * HDP * 2 + AL_ROWS * 128 * 2 + AL_NB * AL_BUNIT_BYTES + 256;
  static bool configured = false;
  if (!configured) {
    FVIT_CUDA(cudaFuncSetAttribute(attn_loop_kernel<HDP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const long long items = (long long)p.groups * p.heads * p.nqt;
  const int sms = num_sms();
  attn_loop_kernel<HDP><<<(unsigned)(items < sms ? items : sms), AL_THREADS, smem, st>>>(tq, tb, p);
  return post_launch("attn_loop_kernel");
}

}  // namespace fvit

using namespace fvit;

extern "C" int fvit_attn_loop_fwd(const void* qkv, int64_t ldq, int32_t groups, int32_t S, int32_t heads, int32_t hdp,
                                  const float* bias, float scale, void* out, int64_t ldo, float* lse, void* stream) {
  FVIT_CHECK(qkv && out && groups > 0 && heads > 0, "fvit_attn_loop_fwd: bad arguments");
  FVIT_CHECK(S > 128, "fvit_attn_loop_fwd: S=%d (windows of up to 128 tokens use fvit_attn_tc_fwd)", S);
  FVIT_CHECK(hdp == 32 || hdp == 64, "fvit_attn_loop_fwd: padded head dim %d unsupported (32 or 64)", hdp);
  FVIT_CHECK(ldq % 8 == 0 && ldo % 8 == 0 && ldq >= 3 * heads * hdp && ldo >= heads * hdp,
             "fvit_attn_loop_fwd: bad leading dimensions");
  FVIT_CHECK(!bias || S % 4 == 0, "fvit_attn_loop_fwd: the bias rows of S=%d tokens are not 16-byte aligned (TMA)", S);
  FVIT_CHECK((reinterpret_cast<uintptr_t>(out) & 15) == 0, "fvit_attn_loop_fwd: out must be 16-byte aligned");
  AttnLoopParams p;
  p.groups = groups, p.S = S, p.heads = heads;
  p.nqt = p.nkt = (S + 127) / 128;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.has_bias = bias ? 1 : 0;
  p.out = (__half*)out, p.ldo = ldo, p.lse = lse;
  CUtensorMap tq, tb;
  {
    uint64_t dims[2] = {(uint64_t)(3 * heads * hdp), (uint64_t)groups * S};
    uint64_t strides[1] = {(uint64_t)ldq * 2};
    uint32_t box[2] = {(uint32_t)hdp, (uint32_t)AL_ROWS};
    int rc = cached_tmap_16bit(&tq, qkv, 2, dims, strides, box,
                               hdp == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  if (bias) {
    uint64_t dims[2] = {(uint64_t)S, (uint64_t)heads * S};
    uint64_t strides[1] = {(uint64_t)S * 4};
    uint32_t box[2] = {32, (uint32_t)AL_ROWS};
    int rc = cached_tmap_f32(&tb, bias, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  } else {
    tb = tq;
  }
  if (hdp == 64) return launch_attn_loop<64>(tq, tb, p, (cudaStream_t)stream);
  return launch_attn_loop<32>(tq, tb, p, (cudaStream_t)stream);
}
