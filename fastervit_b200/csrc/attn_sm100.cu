// fvit_attn_tc_fwd: tensor-core attention core of WindowAttention.forward (fv.py:559-565) for sm_100a.
//
// Work item = (tile of 128 consecutive token rows, head). A tile holds gpt = floor(128 / S) whole
// windows (groups of S tokens: ct_size^2 carrier tokens + ws^2 window tokens, or the carrier grid of
// one image); rows of the box beyond gpt*S belong to the next tile and are masked.
//
//   warp 0     : TMA producer — Q, K, V head slices (128 x hdp, 16-bit) of the packed qkv matrix into
//                a 2-stage ring of swizzled shared-memory tiles
//   warp 1     : MMA issuer — S = Q K^T  (tcgen05.mma, M=128, N=128, K=hdp, fp32 in TMEM, 2 stages)
//                and O = P V (A = P from shared memory, B = V as an MN-major operand, N = hdp)
//   warps 2..5 : softmax — one thread per query row: tcgen05.ld of its 128 scores, block-diagonal
//                window mask, relative-position bias from shared memory, fp32 max / exp2 / sum,
//                P written as fp16 straight into the UMMA (128B-swizzled) operand layout; then the
//                O epilogue (1/rowsum, fp16, 16-byte stores).
// Scores and probabilities never touch HBM; per row the kernel reads 3*hdp*2 B and writes hdp*2 B.
#include <cuda_fp16.h>

#include <mutex>
#include <unordered_map>

#include "../../include/fvit.h"
#include "common.h"
#include "ptx.cuh"

namespace fvit {

constexpr int AT_THREADS = 192;
constexpr int AT_ROWS = 128;

struct AttnParams {
  int groups, S, heads, gpt, tiles;
  int rows_total;
  float scale_log2e;
  const float* bias;  // [heads, S, S] or null
  __half* out;
  long long ldo;
};

template <int HDP>
struct AttnSmem {
  static constexpr int TILE_BYTES = AT_ROWS * HDP * 2;       // one of Q / K / V
  static constexpr int STAGE_BYTES = 3 * TILE_BYTES;
  static constexpr int P_BYTES = AT_ROWS * 128 * 2;          // 128 x 128 fp16, two 64-wide K atoms
  static constexpr int P_OFF = 2 * STAGE_BYTES;
  static constexpr int BIAS_OFF = P_OFF + P_BYTES;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int HDP>
__global__ void __launch_bounds__(AT_THREADS, 1)
    attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ AttnParams p) {
  using SM = AttnSmem<HDP>;
  constexpr uint32_t SWZ = HDP == 64 ? SWZ_128B : SWZ_64B;
  constexpr uint32_t ROW_BYTES = HDP * 2;            // bytes per smem row of Q / K / V
  constexpr uint32_t SBO_QKV = 8 * ROW_BYTES;        // 8-row group stride
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S;
  float* bias_s = reinterpret_cast<float*>(smem + SM::BIAS_OFF);
  uint8_t* ctrl = smem + SM::BIAS_OFF + ((S * S * 4 + 15) & ~15);
  uint64_t* qkv_full = reinterpret_cast<uint64_t*>(ctrl);  // [2]
  uint64_t* qkv_empty = qkv_full + 2;                      // [2]
  uint64_t* s_full = qkv_empty + 2;                        // [2]
  uint64_t* s_empty = s_full + 2;                          // [2]
  uint64_t* p_full = s_empty + 2;                          // [1]
  uint64_t* o_full = p_full + 1;                           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qkv_full[i], 1);
      mbar_init(&qkv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);
    }
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S[2] = {tmem_base, tmem_base + 128};
  const uint32_t tmem_O = tmem_base + 256;

  const int items = p.tiles * p.heads;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
        const int st = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        const int head = w / p.tiles, tile = w % p.tiles;  // head-major: bias is re-staged rarely
        mbar_wait(&qkv_empty[st], ph ^ 1);
        uint8_t* base = smem + st * SM::STAGE_BYTES;
        mbar_expect_tx(&qkv_full[st], SM::STAGE_BYTES);
        const int row0 = tile * p.gpt * S;
        tma_load_2d(base, &tmap_qkv, &qkv_full[st], head * HDP, row0);
        tma_load_2d(base + SM::TILE_BYTES, &tmap_qkv, &qkv_full[st], (p.heads + head) * HDP, row0);
        tma_load_2d(base + 2 * SM::TILE_BYTES, &tmap_qkv, &qkv_full[st], (2 * p.heads + head) * HDP, row0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc_f16(128, 128, 0, 0);
      const uint32_t idesc2 = make_idesc_f16(128, HDP, 0, 1);  // B = V, MN-major
      const uint32_t sP = smem_u32(smem + SM::P_OFF);
      auto issue_pv = [&](int j) {
        // O = P V for item j (local index): waits for the softmax warps' P
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint32_t sV = smem_u32(smem + (j & 1) * SM::STAGE_BYTES + 2 * SM::TILE_BYTES);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {  // 128 keys in steps of 16
          const uint64_t ad = make_smem_desc(sP + (ks >> 2) * (AT_ROWS * 128) + (ks & 3) * 32, 16, 1024, SWZ_128B);
          const uint64_t bd = make_smem_desc(sV + ks * 16 * ROW_BYTES, 0, SBO_QKV, SWZ);
          umma_f16_ss(tmem_O, ad, bd, idesc2, ks > 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(&qkv_empty[j & 1]);
      };
      int it = 0;
      for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
        const int st = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(&qkv_full[st], ph);
        mbar_wait(&s_empty[st], ph ^ 1);
        tc_fence_after();
        const uint32_t sQ = smem_u32(smem + st * SM::STAGE_BYTES);
        const uint32_t sK = sQ + SM::TILE_BYTES;
#pragma unroll
        for (int k = 0; k < HDP / 16; ++k) {
          const uint64_t ad = make_smem_desc(sQ + k * 32, 16, SBO_QKV, SWZ);
          const uint64_t bd = make_smem_desc(sK + k * 32, 16, SBO_QKV, SWZ);
          umma_f16_ss(tmem_S[st], ad, bd, idesc1, k > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[st]);
        if (it > 0) issue_pv(it - 1);
      }
      if (it > 0) issue_pv(it - 1);
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue warps
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // row in tile
    const int tid = threadIdx.x - 64;  // 0..127 among the softmax threads
    const int gi = r / S;
    const int lo = gi * S, hi = lo + S;  // this row's key range inside the tile
    const int wlo = ((quad * 32) / S) * S;
    const int whi = min(128, ((quad * 32 + 31) / S) * S + S);  // union of the warp's key ranges
    uint8_t* sP = smem + SM::P_OFF;
    // columns outside [wlo, whi) are never written by this warp's rows again: zero them once
    for (int j = 0; j < 16; ++j) {
      const int c = j * 8;
      if (c + 8 <= wlo || c >= whi)
        *reinterpret_cast<uint4*>(sP + (c >> 6) * (AT_ROWS * 128) + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4)) =
            make_uint4(0, 0, 0, 0);
    }
    int it = 0;
    int staged_head = -1;
    for (int w = blockIdx.x; w < items; w += gridDim.x, ++it) {
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int head = w / p.tiles, tile = w % p.tiles;
      const bool row_ok = gi < p.gpt && (tile * p.gpt + gi) < p.groups;
      // stage this head's bias (pre-multiplied by log2 e) in shared memory when the head changes;
      // loads are issued in batches of 8 so their latencies overlap
      if (p.bias && head != staged_head) {
        named_bar_sync(1, 128);  // previous head's readers are done
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
        named_bar_sync(1, 128);
        staged_head = head;
      }

      mbar_wait(&s_full[st], ph);
      tc_fence_after();
      const uint32_t ts = tmem_S[st] + ((uint32_t)(quad * 32) << 16);
      const float* brow = bias_s + (r - lo) * S - lo;  // brow[c] for c in [lo, hi)
      // pass 1: row maximum
      float mx = -INFINITY;
      for (int c0 = (wlo / 32) * 32; c0 < whi; c0 += 32) {
        uint32_t raw[32];
        tmem_ld32(ts + c0, raw);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = c0 + j;
            if (c >= lo && c < hi) {
              float s = __uint_as_float(raw[j]) * p.scale_log2e;
              if (p.bias) s += brow[c];
              mx = fmaxf(mx, s);
            }
          }
        }
      }
      // pass 2: probabilities (unnormalised) -> fp16 operand tile, row sum
      float sum = 0.f;
      for (int c0 = (wlo / 32) * 32; c0 < whi; c0 += 32) {
        uint32_t raw[32];
        tmem_ld32(ts + c0, raw);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float e[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int c = c0 + j + u;
            float v = 0.f;
            if (row_ok && c >= lo && c < hi) {
              float s = __uint_as_float(raw[j + u]) * p.scale_log2e;
              if (p.bias) s += brow[c];
              v = exp2f(s - mx);
            }
            e[u] = v;
          }
          const __half2 h = __floats2half2_rn(e[0], e[1]);
          // accumulate what the tensor core will actually multiply (fp16-rounded probabilities)
          sum += __low2float(h) + __high2float(h);
          pk[j >> 1] = *reinterpret_cast<const uint32_t*>(&h);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = c0 + q * 8;
          *reinterpret_cast<uint4*>(sP + (c >> 6) * (AT_ROWS * 128) + r * 128 +
                                    ((((c & 63) >> 3) ^ (r & 7)) << 4)) =
              make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      // P is complete for this warp's rows: publish to the async proxy, release S, signal the MMA warp
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(p_full);
        mbar_arrive(&s_empty[st]);
      }
      // O epilogue
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      const uint32_t to = tmem_O + ((uint32_t)(quad * 32) << 16);
      const float inv = row_ok ? 1.f / sum : 0.f;
      const long long grow = (long long)tile * p.gpt * S + r;
      __half* orow = p.out + grow * p.ldo + head * HDP;
#pragma unroll
      for (int c0 = 0; c0 < HDP; c0 += 32) {
        uint32_t raw[32];
        tmem_ld32(to + c0, raw);
        tmem_ld_wait();
        if (row_ok && grow < p.rows_total) {
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
      tc_fence_before();  // O reads retire before this thread's next p_full arrive
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static std::mutex g_at_mu;
static std::unordered_map<uint64_t, CUtensorMap> g_at_cache;

template <int HDP>
static int launch_attn(const CUtensorMap& tm, const AttnParams& p, size_t smem, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    FVIT_CUDA(cudaFuncSetAttribute(attn_tc_kernel<HDP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int items = p.tiles * p.heads;
  const int sms = num_sms();
  attn_tc_kernel<HDP><<<items < sms ? items : sms, AT_THREADS, smem, st>>>(tm, p);
  return post_launch("attn_tc_kernel");
}

}  // namespace fvit

using namespace fvit;

extern "C" int fvit_attn_tc_fwd(const void* qkv, int64_t ldq, int32_t groups, int32_t S, int32_t heads,
                                int32_t hdp, const float* bias, float scale, void* out, int64_t ldo,
                                void* stream) {
  FVIT_CHECK(qkv && out && groups > 0 && heads > 0, "fvit_attn_tc_fwd: bad arguments");
  FVIT_CHECK(S >= 1 && S <= 128, "fvit_attn_tc_fwd: S=%d unsupported (1..128)", S);
  FVIT_CHECK(hdp == 32 || hdp == 64, "fvit_attn_tc_fwd: padded head dim %d unsupported (32 or 64)", hdp);
  FVIT_CHECK(ldq % 8 == 0 && ldo % 8 == 0 && ldq >= 3 * heads * hdp && ldo >= heads * hdp,
             "fvit_attn_tc_fwd: bad leading dimensions");
  FVIT_CHECK((reinterpret_cast<uintptr_t>(out) & 15) == 0, "fvit_attn_tc_fwd: out must be 16-byte aligned");
  AttnParams p;
  p.groups = groups;
  p.S = S;
  p.heads = heads;
  p.gpt = AT_ROWS / S;
  p.tiles = (groups + p.gpt - 1) / p.gpt;
  p.rows_total = groups * S;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.bias = bias;
  p.out = (__half*)out;
  p.ldo = ldo;
  CUtensorMap tm;
  {
    const uint64_t key = reinterpret_cast<uint64_t>(qkv) ^ ((uint64_t)ldq << 40) ^ ((uint64_t)p.rows_total << 8) ^
                         (uint64_t)hdp ^ ((uint64_t)heads << 52);
    std::lock_guard<std::mutex> g(g_at_mu);
    auto it = g_at_cache.find(key);
    if (it != g_at_cache.end()) {
      tm = it->second;
    } else {
      uint64_t dims[2] = {(uint64_t)(3 * heads * hdp), (uint64_t)p.rows_total};
      uint64_t strides[1] = {(uint64_t)ldq * 2};
      uint32_t box[2] = {(uint32_t)hdp, AT_ROWS};
      int rc = encode_tmap_16bit(&tm, qkv, 2, dims, strides, box,
                                 hdp == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
      if (g_at_cache.size() > 4096) g_at_cache.clear();
      g_at_cache.emplace(key, tm);
    }
  }
  const size_t smem = 1024 + (hdp == 64 ? AttnSmem<64>::BIAS_OFF : AttnSmem<32>::BIAS_OFF) +
                      ((size_t)S * S * 4 + 15) / 16 * 16 + 256;
  FVIT_CHECK(smem <= 227 * 1024, "fvit_attn_tc_fwd: needs %zu B of shared memory", smem);
  if (hdp == 64) return launch_attn<64>(tm, p, smem, (cudaStream_t)stream);
  return launch_attn<32>(tm, p, smem, (cudaStream_t)stream);
}
