// Optimizer step on the flat gradient buffer: the work either side of loss.backward() in the reference's
// training loop (train.py:879-899): AMP unscale + non-finite check, clip-grad-norm (`--clip-grad 5.0
// --clip-mode norm`), AdamW (`--opt adamw`, fv0-3) or LAMB (`--opt lamb`, fv4-6; TRAINING.md:28,105),
// and the ModelEmaV2 update (train.py:898-899). In the reference these are hundreds of foreach / per-tensor
// ATen launches; here they are 3-4 HBM-bound launches over one "chunk table".
//
// Layout: every parameter is a *segment*. Gradients and both moments live in flat fp32 buffers (segment s at
// element offset seg_off[s]; the backward pass already produces gradients in that layout), parameters (and the
// EMA copies) stay where torch allocated them and are addressed through per-segment pointer tables. A chunk
// {seg, start, count, 0} is a run of <= chunk-size elements of one segment, processed by one CTA, so per-segment
// hyper-parameters (lr, weight decay) and the LAMB per-tensor norms are CTA-uniform.
//
// Algorithmic bytes per parameter element: sqnorm 4 (read g); AdamW 28 (read p,g,m,v; write p,m,v), +8 with the
// fused EMA; LAMB 28 + 12; EMA alone 12. Nothing here synchronises with the host: the step counter, bias
// corrections, clip coefficient and the skip-on-overflow flag are device scalars (scal[]).
#include <math.h>

#include "../../include/fvit.h"
#include "common.h"

namespace fvit {

constexpr int OPT_THREADS = 256;

__device__ __forceinline__ float opt_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum of `v` over the 256 threads of the CTA, valid in thread 0
__device__ __forceinline__ float opt_block_sum(float v, float* red /* [8] */) {
  v = opt_warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (warp == 0) {
    t = lane < OPT_THREADS / 32 ? red[lane] : 0.f;
    t = opt_warp_sum(t);
  }
  __syncthreads();
  return t;
}
__device__ __forceinline__ bool opt_aligned16(const void* a, const void* b = nullptr, const void* c = nullptr,
                                              const void* d = nullptr, const void* e = nullptr) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
           reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(e)) & 15) == 0;
}

// flat[seg_off[seg] + start + i] = src_seg[start + i]   (gradients that autograd did not leave in one buffer)
__global__ void __launch_bounds__(OPT_THREADS)
optim_gather_kernel(const int4* __restrict__ chunks, const long long* __restrict__ seg_src,
                    const long long* __restrict__ seg_off, float* __restrict__ flat) {
  const int4 ck = chunks[blockIdx.x];
  const float* src = reinterpret_cast<const float*>(seg_src[ck.x]) + ck.y;
  float* dst = flat + seg_off[ck.x] + ck.y;
  const int n = ck.z;
  if (opt_aligned16(src, dst)) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += OPT_THREADS)
      reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += OPT_THREADS) dst[i] = src[i];
  } else {
    for (int i = threadIdx.x; i < n; i += OPT_THREADS) dst[i] = src[i];
  }
}

// partials[2c] = sum g^2 over chunk c, partials[2c+1] = number of non-finite values in it
__global__ void __launch_bounds__(OPT_THREADS)
optim_sqnorm_kernel(const int4* __restrict__ chunks, const long long* __restrict__ seg_off,
                    const float* __restrict__ g, float* __restrict__ partials) {
  __shared__ float red[8];
  const int4 ck = chunks[blockIdx.x];
  const float* gp = g + seg_off[ck.x] + ck.y;
  const int n = ck.z;
  float s = 0.f, bad = 0.f;
  if (opt_aligned16(gp)) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(gp) + i);
      s = fmaf(t.x, t.x, s), s = fmaf(t.y, t.y, s), s = fmaf(t.z, t.z, s), s = fmaf(t.w, t.w, s);
      bad += (isfinite(t.x) ? 0.f : 1.f) + (isfinite(t.y) ? 0.f : 1.f) + (isfinite(t.z) ? 0.f : 1.f) +
             (isfinite(t.w) ? 0.f : 1.f);
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += OPT_THREADS) {
      const float t = gp[i];
      s = fmaf(t, t, s);
      bad += isfinite(t) ? 0.f : 1.f;
    }
  } else {
    for (int i = threadIdx.x; i < n; i += OPT_THREADS) {
      const float t = gp[i];
      s = fmaf(t, t, s);
      bad += isfinite(t) ? 0.f : 1.f;
    }
  }
  s = opt_block_sum(s, red);
  bad = opt_block_sum(bad, red);
  if (threadIdx.x == 0) {
    partials[2 * blockIdx.x] = s;
    partials[2 * blockIdx.x + 1] = bad;
  }
}

// One CTA: total gradient norm, overflow flag, clip coefficient, step counter and bias corrections.
//   scal[0] = ||g|| / grad_scale (the true gradient norm; 0 when nchunks == 0)
//   scal[1] = 1 if any gradient is non-finite (or *found_inf_in != 0): every update kernel then skips the step
//   scal[2] = multiplier applied to g: (1 / grad_scale) * min(1, max_norm / (norm + clip_eps))
//   scal[3] = number of steps taken (incremented here unless skipped), scal[4] = 1 - beta1^step, scal[5] = 1 - beta2^step
__global__ void __launch_bounds__(1024)
optim_prepare_kernel(const float* __restrict__ partials, int nchunks, const float* __restrict__ grad_scale,
                     const float* __restrict__ found_inf_in, float max_norm, float clip_eps, double beta1, double beta2,
                     float* __restrict__ scal) {
  __shared__ double rs[32], rb[32];
  double s = 0.0, bad = 0.0;
  for (int i = threadIdx.x; i < nchunks; i += blockDim.x) {
    s += (double)partials[2 * i];
    bad += (double)partials[2 * i + 1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if ((threadIdx.x & 31) == 0) rs[threadIdx.x >> 5] = s, rb[threadIdx.x >> 5] = bad;
  __syncthreads();
  if (threadIdx.x == 0) {
    s = 0.0, bad = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += rs[i], bad += rb[i];
    const float inv = grad_scale ? 1.f / __ldg(grad_scale) : 1.f;
    const bool skip = bad > 0.0 || !isfinite(s) || !isfinite(inv) || (found_inf_in && __ldg(found_inf_in) != 0.f);
    const float norm = (float)(sqrt(s) * (double)inv);
    float clip = 1.f;
    if (max_norm > 0.f) clip = fminf(1.f, max_norm / (norm + clip_eps));
    scal[0] = norm;
    scal[1] = skip ? 1.f : 0.f;
    scal[2] = inv * clip;
    if (!skip) {
      const float step = scal[3] + 1.f;
      scal[3] = step;
      scal[4] = (float)(1.0 - pow(beta1, (double)step));
      scal[5] = (float)(1.0 - pow(beta2, (double)step));
    }
  }
}

struct AdamwCoef {
  float gm, b1, b2, eps, decay, step_size, rs2, ema_decay;
};
// torch.optim.AdamW (decoupled decay, bias-corrected): p *= 1 - lr wd; m = lerp(m, g, 1-b1);
// v = b2 v + (1-b2) g^2; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
__device__ __forceinline__ void adamw_elem(float& p, float g, float& m, float& v, const AdamwCoef& c) {
  g *= c.gm;
  p *= c.decay;
  m = fmaf(g - m, 1.f - c.b1, m);
  v = fmaf(v, c.b2, (1.f - c.b2) * g * g);
  const float denom = sqrtf(v) / c.rs2 + c.eps;
  p = fmaf(-c.step_size, m / denom, p);
}

__global__ void __launch_bounds__(OPT_THREADS)
optim_adamw_kernel(const int4* __restrict__ chunks, const long long* __restrict__ seg_p,
                   const long long* __restrict__ seg_off, const float* __restrict__ seg_hp,
                   const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, float beta1, float beta2,
                   float eps, const float* __restrict__ scal, const long long* __restrict__ seg_ema, float ema_decay) {
  // overflow: skip the step (GradScaler semantics); a fused EMA still blends the unchanged parameters in
  const bool skip = __ldg(scal + 1) != 0.f;
  if (skip && !seg_ema) return;
  const int4 ck = chunks[blockIdx.x];
  const int seg = ck.x, n = ck.z;
  float* pp = reinterpret_cast<float*>(seg_p[seg]) + ck.y;
  float* ep = seg_ema ? reinterpret_cast<float*>(seg_ema[seg]) + ck.y : nullptr;
  const long long fo = seg_off[seg] + ck.y;
  const float* gp = g + fo;
  float* mp = m + fo;
  float* vp = v + fo;
  const float lr = __ldg(seg_hp + 2 * seg), wd = __ldg(seg_hp + 2 * seg + 1);
  AdamwCoef c;
  c.gm = __ldg(scal + 2), c.b1 = beta1, c.b2 = beta2, c.eps = eps;
  c.decay = 1.f - lr * wd;
  c.step_size = lr / __ldg(scal + 4);
  c.rs2 = sqrtf(__ldg(scal + 5));
  c.ema_decay = ema_decay;
  int i0 = 0;
  if (opt_aligned16(pp, gp, mp, vp, ep)) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
      float4 p4 = reinterpret_cast<float4*>(pp)[i];
      if (!skip) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gp) + i);
        float4 m4 = reinterpret_cast<float4*>(mp)[i], v4 = reinterpret_cast<float4*>(vp)[i];
        adamw_elem(p4.x, g4.x, m4.x, v4.x, c), adamw_elem(p4.y, g4.y, m4.y, v4.y, c);
        adamw_elem(p4.z, g4.z, m4.z, v4.z, c), adamw_elem(p4.w, g4.w, m4.w, v4.w, c);
        reinterpret_cast<float4*>(pp)[i] = p4;
        reinterpret_cast<float4*>(mp)[i] = m4;
        reinterpret_cast<float4*>(vp)[i] = v4;
      }
      if (ep) {
        float4 e4 = reinterpret_cast<float4*>(ep)[i];
        e4.x = e4.x * ema_decay + (1.f - ema_decay) * p4.x, e4.y = e4.y * ema_decay + (1.f - ema_decay) * p4.y;
        e4.z = e4.z * ema_decay + (1.f - ema_decay) * p4.z, e4.w = e4.w * ema_decay + (1.f - ema_decay) * p4.w;
        reinterpret_cast<float4*>(ep)[i] = e4;
      }
    }
    i0 = n4 << 2;
  }
  for (int i = i0 + threadIdx.x; i < n; i += OPT_THREADS) {
    float p = pp[i];
    if (!skip) {
      float mm = mp[i], vv = vp[i];
      adamw_elem(p, gp[i], mm, vv, c);
      pp[i] = p, mp[i] = mm, vp[i] = vv;
    }
    if (ep) ep[i] = ep[i] * ema_decay + (1.f - ema_decay) * p;
  }
}

// LAMB (timm.optim.Lamb, the `--opt lamb` of TRAINING.md:105), stage 1: moments + un-trusted update
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; u = (m/bc1) / (sqrt(v)/sqrt(bc2) + eps) + wd p
// u goes to its own flat buffer; per-segment sum p^2 / sum u^2 are accumulated into seg_norms[2 seg + {0,1}].
__device__ __forceinline__ float lamb_elem(float p, float g, float& m, float& v, float gm, float b1, float b2, float eps,
                                           float inv_bc1, float rs2, float wd) {
  g *= gm;
  m = fmaf(m, b1, (1.f - b1) * g);
  v = fmaf(v, b2, (1.f - b2) * g * g);
  const float denom = sqrtf(v) / rs2 + eps;
  return fmaf(wd, p, (m * inv_bc1) / denom);
}

__global__ void __launch_bounds__(OPT_THREADS)
optim_lamb1_kernel(const int4* __restrict__ chunks, const long long* __restrict__ seg_p,
                   const long long* __restrict__ seg_off, const float* __restrict__ seg_hp,
                   const float* __restrict__ g, float* __restrict__ u, float* __restrict__ m, float* __restrict__ v,
                   float beta1, float beta2, float eps, const float* __restrict__ scal, float* __restrict__ seg_norms) {
  __shared__ float red[8];
  if (__ldg(scal + 1) != 0.f) return;
  const int4 ck = chunks[blockIdx.x];
  const int seg = ck.x, n = ck.z;
  const float* pp = reinterpret_cast<const float*>(seg_p[seg]) + ck.y;
  const long long fo = seg_off[seg] + ck.y;
  const float* gp = g + fo;
  float* up = u + fo;
  float* mp = m + fo;
  float* vp = v + fo;
  const float wd = __ldg(seg_hp + 2 * seg + 1);
  const float gm = __ldg(scal + 2), inv_bc1 = 1.f / __ldg(scal + 4), rs2 = sqrtf(__ldg(scal + 5));
  float sp = 0.f, su = 0.f;
  int i0 = 0;
  if (opt_aligned16(pp, gp, mp, vp, up)) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
      const float4 p4 = __ldg(reinterpret_cast<const float4*>(pp) + i);
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(gp) + i);
      float4 m4 = reinterpret_cast<float4*>(mp)[i], v4 = reinterpret_cast<float4*>(vp)[i], u4;
      u4.x = lamb_elem(p4.x, g4.x, m4.x, v4.x, gm, beta1, beta2, eps, inv_bc1, rs2, wd);
      u4.y = lamb_elem(p4.y, g4.y, m4.y, v4.y, gm, beta1, beta2, eps, inv_bc1, rs2, wd);
      u4.z = lamb_elem(p4.z, g4.z, m4.z, v4.z, gm, beta1, beta2, eps, inv_bc1, rs2, wd);
      u4.w = lamb_elem(p4.w, g4.w, m4.w, v4.w, gm, beta1, beta2, eps, inv_bc1, rs2, wd);
      reinterpret_cast<float4*>(mp)[i] = m4;
      reinterpret_cast<float4*>(vp)[i] = v4;
      reinterpret_cast<float4*>(up)[i] = u4;
      sp += p4.x * p4.x + p4.y * p4.y + p4.z * p4.z + p4.w * p4.w;
      su += u4.x * u4.x + u4.y * u4.y + u4.z * u4.z + u4.w * u4.w;
    }
    i0 = n4 << 2;
  }
  for (int i = i0 + threadIdx.x; i < n; i += OPT_THREADS) {
    const float p = pp[i];
    float mm = mp[i], vv = vp[i];
    const float uu = lamb_elem(p, gp[i], mm, vv, gm, beta1, beta2, eps, inv_bc1, rs2, wd);
    mp[i] = mm, vp[i] = vv, up[i] = uu;
    sp = fmaf(p, p, sp);
    su = fmaf(uu, uu, su);
  }
  sp = opt_block_sum(sp, red);
  su = opt_block_sum(su, red);
  if (threadIdx.x == 0) {
    atomicAdd(seg_norms + 2 * seg, sp);
    atomicAdd(seg_norms + 2 * seg + 1, su);
  }
}

// LAMB stage 2: p -= lr * trust * u with trust = ||p|| / ||u|| per tensor (1 when either norm is 0; only for
// segments with weight decay unless always_adapt; optionally clipped to <= 1), plus the optional fused EMA.
__global__ void __launch_bounds__(OPT_THREADS)
optim_lamb2_kernel(const int4* __restrict__ chunks, const long long* __restrict__ seg_p,
                   const long long* __restrict__ seg_off, const float* __restrict__ seg_hp,
                   const float* __restrict__ u, const float* __restrict__ seg_norms, int trust_clip, int always_adapt,
                   const float* __restrict__ scal, const long long* __restrict__ seg_ema, float ema_decay) {
  const bool skip = __ldg(scal + 1) != 0.f;  // skipped step: only the fused EMA (of the unchanged parameters) runs
  if (skip && !seg_ema) return;
  const int4 ck = chunks[blockIdx.x];
  const int seg = ck.x, n = ck.z;
  float* pp = reinterpret_cast<float*>(seg_p[seg]) + ck.y;
  float* ep = seg_ema ? reinterpret_cast<float*>(seg_ema[seg]) + ck.y : nullptr;
  const float* up = u + seg_off[seg] + ck.y;
  const float lr = __ldg(seg_hp + 2 * seg), wd = __ldg(seg_hp + 2 * seg + 1);
  float trust = 1.f;
  if (wd != 0.f || always_adapt) {
    const float wn = sqrtf(__ldg(seg_norms + 2 * seg)), un = sqrtf(__ldg(seg_norms + 2 * seg + 1));
    if (wn > 0.f && un > 0.f) trust = wn / un;
    if (trust_clip) trust = fminf(trust, 1.f);
  }
  const float a = -lr * trust;
  int i0 = 0;
  if (opt_aligned16(pp, up, ep)) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
      float4 p4 = reinterpret_cast<float4*>(pp)[i];
      if (!skip) {
        const float4 u4 = __ldg(reinterpret_cast<const float4*>(up) + i);
        p4.x = fmaf(a, u4.x, p4.x), p4.y = fmaf(a, u4.y, p4.y), p4.z = fmaf(a, u4.z, p4.z), p4.w = fmaf(a, u4.w, p4.w);
        reinterpret_cast<float4*>(pp)[i] = p4;
      }
      if (ep) {
        float4 e4 = reinterpret_cast<float4*>(ep)[i];
        e4.x = e4.x * ema_decay + (1.f - ema_decay) * p4.x, e4.y = e4.y * ema_decay + (1.f - ema_decay) * p4.y;
        e4.z = e4.z * ema_decay + (1.f - ema_decay) * p4.z, e4.w = e4.w * ema_decay + (1.f - ema_decay) * p4.w;
        reinterpret_cast<float4*>(ep)[i] = e4;
      }
    }
    i0 = n4 << 2;
  }
  for (int i = i0 + threadIdx.x; i < n; i += OPT_THREADS) {
    float p = pp[i];
    if (!skip) {
      p = fmaf(a, up[i], p);
      pp[i] = p;
    }
    if (ep) ep[i] = ep[i] * ema_decay + (1.f - ema_decay) * p;
  }
}

// ModelEmaV2.update (train.py:898-899; timm utils/model_ema.py): ema = decay * ema + (1 - decay) * model, for
// every floating-point entry of the state_dict (parameters and BatchNorm running statistics alike).
__global__ void __launch_bounds__(OPT_THREADS)
optim_ema_kernel(const int4* __restrict__ chunks, const long long* __restrict__ seg_ema,
                 const long long* __restrict__ seg_src, float decay) {
  const int4 ck = chunks[blockIdx.x];
  float* ep = reinterpret_cast<float*>(seg_ema[ck.x]) + ck.y;
  const float* sp = reinterpret_cast<const float*>(seg_src[ck.x]) + ck.y;
  const int n = ck.z;
  const float w = 1.f - decay;
  int i0 = 0;
  if (opt_aligned16(ep, sp)) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
      float4 e4 = reinterpret_cast<float4*>(ep)[i];
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(sp) + i);
      e4.x = e4.x * decay + w * s4.x, e4.y = e4.y * decay + w * s4.y, e4.z = e4.z * decay + w * s4.z,
      e4.w = e4.w * decay + w * s4.w;
      reinterpret_cast<float4*>(ep)[i] = e4;
    }
    i0 = n4 << 2;
  }
  for (int i = i0 + threadIdx.x; i < n; i += OPT_THREADS) ep[i] = ep[i] * decay + w * sp[i];
}

}  // namespace fvit

using namespace fvit;

extern "C" {

int fvit_optim_gather_f32(const int32_t* chunks, int32_t nchunks, const int64_t* seg_src, const int64_t* seg_off,
                          float* flat, void* stream) {
  FVIT_CHECK(chunks && seg_src && seg_off && flat && nchunks > 0, "fvit_optim_gather_f32: bad arguments");
  optim_gather_kernel<<<nchunks, OPT_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const int4*>(chunks), reinterpret_cast<const long long*>(seg_src),
      reinterpret_cast<const long long*>(seg_off), flat);
  return post_launch("optim_gather_kernel");
}

int fvit_optim_sqnorm(const int32_t* chunks, int32_t nchunks, const int64_t* seg_off, const float* g, float* partials,
                      void* stream) {
  FVIT_CHECK(chunks && seg_off && g && partials && nchunks > 0, "fvit_optim_sqnorm: bad arguments");
  optim_sqnorm_kernel<<<nchunks, OPT_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const int4*>(chunks), reinterpret_cast<const long long*>(seg_off), g, partials);
  return post_launch("optim_sqnorm_kernel");
}

int fvit_optim_prepare(const float* partials, int32_t nchunks, const float* grad_scale, const float* found_inf,
                       float max_norm, float clip_eps, double beta1, double beta2, float* scal, void* stream) {
  FVIT_CHECK(scal && nchunks >= 0 && (nchunks == 0 || partials), "fvit_optim_prepare: bad arguments");
  optim_prepare_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(partials, nchunks, grad_scale, found_inf, max_norm,
                                                             clip_eps, beta1, beta2, scal);
  return post_launch("optim_prepare_kernel");
}

int fvit_optim_adamw(const int32_t* chunks, int32_t nchunks, const int64_t* seg_p, const int64_t* seg_off,
                     const float* seg_hp, const float* g, float* m, float* v, float beta1, float beta2, float eps,
                     const float* scal, const int64_t* seg_ema, float ema_decay, void* stream) {
  FVIT_CHECK(chunks && seg_p && seg_off && seg_hp && g && m && v && scal && nchunks > 0,
             "fvit_optim_adamw: bad arguments");
  optim_adamw_kernel<<<nchunks, OPT_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const int4*>(chunks), reinterpret_cast<const long long*>(seg_p),
      reinterpret_cast<const long long*>(seg_off), seg_hp, g, m, v, beta1, beta2, eps, scal,
      reinterpret_cast<const long long*>(seg_ema), ema_decay);
  return post_launch("optim_adamw_kernel");
}

int fvit_optim_lamb_stage1(const int32_t* chunks, int32_t nchunks, const int64_t* seg_p, const int64_t* seg_off,
                           const float* seg_hp, const float* g, float* u, float* m, float* v, float beta1,
                           float beta2, float eps, const float* scal, float* seg_norms, void* stream) {
  FVIT_CHECK(chunks && seg_p && seg_off && seg_hp && g && u && m && v && scal && seg_norms && nchunks > 0,
             "fvit_optim_lamb_stage1: bad arguments");
  optim_lamb1_kernel<<<nchunks, OPT_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const int4*>(chunks), reinterpret_cast<const long long*>(seg_p),
      reinterpret_cast<const long long*>(seg_off), seg_hp, g, u, m, v, beta1, beta2, eps, scal, seg_norms);
  return post_launch("optim_lamb1_kernel");
}

int fvit_optim_lamb_stage2(const int32_t* chunks, int32_t nchunks, const int64_t* seg_p, const int64_t* seg_off,
                           const float* seg_hp, const float* u, const float* seg_norms, int32_t trust_clip,
                           int32_t always_adapt, const float* scal, const int64_t* seg_ema, float ema_decay,
                           void* stream) {
  FVIT_CHECK(chunks && seg_p && seg_off && seg_hp && u && seg_norms && scal && nchunks > 0,
             "fvit_optim_lamb_stage2: bad arguments");
  optim_lamb2_kernel<<<nchunks, OPT_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const int4*>(chunks), reinterpret_cast<const long long*>(seg_p),
      reinterpret_cast<const long long*>(seg_off), seg_hp, u, seg_norms, trust_clip, always_adapt, scal,
      reinterpret_cast<const long long*>(seg_ema), ema_decay);
  return post_launch("optim_lamb2_kernel");
}

int fvit_optim_ema(const int32_t* chunks, int32_t nchunks, const int64_t* seg_ema, const int64_t* seg_src, float decay,
                   void* stream) {
  FVIT_CHECK(chunks && seg_ema && seg_src && nchunks > 0, "fvit_optim_ema: bad arguments");
  optim_ema_kernel<<<nchunks, OPT_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const int4*>(chunks), reinterpret_cast<const long long*>(seg_ema),
      reinterpret_cast<const long long*>(seg_src), decay);
  return post_launch("optim_ema_kernel");
}

}  // extern "C"
