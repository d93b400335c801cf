// Fused sampler-step kernels (HBM-bound, fp32, bit-exact with the reference op order).
//
// Replaces the elementwise tails of EODiffusion.sampling / _reverse_diffusion(_with_clip)
// (reference diffusion/model.py:58-60, 94-98, 110-122, 133-150) and of
// DDIMSampler.p_sample_ddim (diffusion/ddim.py:187-207).
//
// Every arithmetic op is written with the round-to-nearest intrinsics so nvcc cannot
// contract a mul+add into an FMA: the reference evaluates each torch op separately and
// rounds after each one, and the parity bar for this kernel is bit equality.
//
// Algorithmic bytes per image-step (C=3, 256x256): mix 4C+1, step 4C, fused step+mix
// 6C+1 planes of H*W*4 bytes ("sampler step").
#include "common.cuh"

namespace eo {

struct Coefs {
  float v[EO_DDPM_NCOEF];
};

__device__ __forceinline__ Coefs load_coefs(const float* __restrict__ table, long long t) {
  Coefs c;
  const float4* p = reinterpret_cast<const float4*>(table + t * EO_DDPM_NCOEF);
#pragma unroll
  for (int i = 0; i < EO_DDPM_NCOEF / 4; ++i) {
    float4 q = __ldg(p + i);
    c.v[4 * i + 0] = q.x; c.v[4 * i + 1] = q.y; c.v[4 * i + 2] = q.z; c.v[4 * i + 3] = q.w;
  }
  return c;
}

// mask*(sa*gt + sb*noise) + (1-mask)*x      (model.py:59-60 with :97-98 inlined)
__device__ __forceinline__ float mix1(float x, float gt, float m, float nz, float sa, float sb) {
  float gtn = __fadd_rn(__fmul_rn(sa, gt), __fmul_rn(sb, nz));
  return __fadd_rn(__fmul_rn(m, gtn), __fmul_rn(__fsub_rn(1.0f, m), x));
}

// model.py:137-150 (clip) / :114,122 (no clip)
template <bool CLIP>
__device__ __forceinline__ float step1(float x, float e, float nz, const Coefs& c, bool pos) {
  float mean, std;
  if (CLIP) {
    float x0 = __fsub_rn(__fmul_rn(c.v[EO_COEF_SQRT_RECIP_ACP], x),
                         __fmul_rn(c.v[EO_COEF_SQRT_RECIPM1_ACP], e));
    x0 = x0 < -1.0f ? -1.0f : (x0 > 1.0f ? 1.0f : x0);  // clamp_ (NaN passes through)
    if (pos) {
      mean = __fadd_rn(__fmul_rn(c.v[EO_COEF_MEAN_X0], x0), __fmul_rn(c.v[EO_COEF_MEAN_XT], x));
      std = c.v[EO_COEF_STD];
    } else {
      mean = __fmul_rn(c.v[EO_COEF_MEAN_X0_T0], x0);
      std = 0.0f;
    }
  } else {
    mean = __fmul_rn(c.v[EO_COEF_RECIP_SQRT_ALPHA],
                     __fsub_rn(x, __fmul_rn(c.v[EO_COEF_EPS_NOCLIP], e)));
    std = pos ? c.v[EO_COEF_STD] : 0.0f;
  }
  return __fadd_rn(mean, __fmul_rn(std, nz));
}

// One thread handles VEC consecutive elements along HW of one (b, c) plane.
template <int VEC>
struct Pack;
template <>
struct Pack<4> {
  float4 d;
  __device__ __forceinline__ void load(const float* p) { d = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = d; }
  __device__ __forceinline__ float& at(int i) { return (&d.x)[i]; }
};
template <>
struct Pack<1> {
  float d;
  __device__ __forceinline__ void load(const float* p) { d = *p; }
  __device__ __forceinline__ void store(float* p) const { *p = d; }
  __device__ __forceinline__ float& at(int) { return d; }
};

template <int VEC>
__global__ void __launch_bounds__(256)
k_sum_mix(const float* x_t, const float* __restrict__ gt,          // x_t and x_out may alias (eo_b200.h): no __restrict__
          const float* __restrict__ mask, const float* __restrict__ noise,
          const long long* __restrict__ ts, const float* __restrict__ table,
          float* x_out, int C, int HWv, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long plane = i / HWv;
    int hw = (int)(i - plane * HWv);
    int b = (int)(plane / C);
    Coefs c = load_coefs(table, ts[b]);
    Pack<VEC> px, pg, pm, pn, po;
    px.load(x_t + i * VEC); pg.load(gt + i * VEC); pn.load(noise + i * VEC);
    pm.load(mask + ((long long)b * HWv + hw) * VEC);
#pragma unroll
    for (int k = 0; k < VEC; ++k)
      po.at(k) = mix1(px.at(k), pg.at(k), pm.at(k), pn.at(k), c.v[EO_COEF_SQRT_ACP],
                      c.v[EO_COEF_SQRT_1M_ACP]);
    po.store(x_out + i * VEC);
  }
}

template <int VEC, bool CLIP, bool MIX>
__global__ void __launch_bounds__(256)
k_step(const float* x_t, const float* __restrict__ eps,            // x_t and x_out may alias: no __restrict__
       const float* __restrict__ noise, const long long* __restrict__ ts,
       const float* __restrict__ gt, const float* __restrict__ mask,
       const float* __restrict__ noise_next, const long long* __restrict__ ts_next,
       const float* __restrict__ table, float* x_out, int C, int HWv,
       long long total, int all_pos) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long plane = i / HWv;
    int hw = (int)(i - plane * HWv);
    int b = (int)(plane / C);
    Coefs c = load_coefs(table, ts[b]);
    Pack<VEC> px, pe, pn, po;
    px.load(x_t + i * VEC); pe.load(eps + i * VEC); pn.load(noise + i * VEC);
#pragma unroll
    for (int k = 0; k < VEC; ++k)
      po.at(k) = step1<CLIP>(px.at(k), pe.at(k), pn.at(k), c, all_pos != 0);
    if (MIX) {
      Coefs cn = load_coefs(table, ts_next[b]);
      Pack<VEC> pg, pm, pnn;
      pg.load(gt + i * VEC); pnn.load(noise_next + i * VEC);
      pm.load(mask + ((long long)b * HWv + hw) * VEC);
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        po.at(k) = mix1(po.at(k), pg.at(k), pm.at(k), pnn.at(k), cn.v[EO_COEF_SQRT_ACP],
                        cn.v[EO_COEF_SQRT_1M_ACP]);
    }
    po.store(x_out + i * VEC);
  }
}

// ddim.py:198-206
template <int VEC>
__global__ void __launch_bounds__(256)
k_ddim_step(const float* x, const float* __restrict__ e_t,       // x_prev / pred_x0 may alias x: no __restrict__
            const float* __restrict__ noise, float* x_prev,
            float* pred_x0, float sqrt_a_t, float sqrt_1m_a_t, float sqrt_a_prev,
            float dir_coef, float sigma_t, float temperature, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    Pack<VEC> px, pe, pn, po, pp;
    px.load(x + i * VEC); pe.load(e_t + i * VEC);
    if (noise) pn.load(noise + i * VEC);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float p0 = __fdiv_rn(__fsub_rn(px.at(k), __fmul_rn(sqrt_1m_a_t, pe.at(k))), sqrt_a_t);
      float dir = __fmul_rn(dir_coef, pe.at(k));
      float nz = noise ? __fmul_rn(__fmul_rn(sigma_t, pn.at(k)), temperature) : 0.0f;
      pp.at(k) = p0;
      po.at(k) = __fadd_rn(__fadd_rn(__fmul_rn(sqrt_a_prev, p0), dir), nz);
    }
    po.store(x_prev + i * VEC);
    pp.store(pred_x0 + i * VEC);
  }
}

template <int VEC>
__global__ void __launch_bounds__(256)
k_cfg_combine(const float* eu, const float* ec, float s, float* out, long long total) {   // out may alias either input
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    Pack<VEC> a, b, o;
    a.load(eu + i * VEC); b.load(ec + i * VEC);
#pragma unroll
    for (int k = 0; k < VEC; ++k)
      o.at(k) = __fadd_rn(a.at(k), __fmul_rn(s, __fsub_rn(b.at(k), a.at(k))));
    o.store(out + i * VEC);
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// grid: enough CTAs for >= 2 waves of 148 SMs x 8 resident CTAs, capped by the work
static inline int grid_for(long long total) {
  long long need = ceil_div(total, 256);
  long long cap = (long long)num_sms() * 16;
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

}  // namespace eo

using namespace eo;

extern "C" int eo_ddpm_sum_mix(const float* x_t, const float* gt, const float* mask,
                               const float* noise, const int64_t* timesteps, const float* table,
                               float* x_out, int B, int C, int HW, void* stream) {
  EO_REQUIRE(x_t && gt && mask && noise && timesteps && table && x_out, EO_ERR_ARG,
             "eo_ddpm_sum_mix: null pointer");
  EO_REQUIRE(B > 0 && C > 0 && HW > 0, EO_ERR_ARG, "eo_ddpm_sum_mix: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  bool v4 = (HW % 4 == 0) && aligned16(x_t) && aligned16(gt) && aligned16(mask) &&
            aligned16(noise) && aligned16(x_out);
  const long long* ts = reinterpret_cast<const long long*>(timesteps);
  if (v4) {
    long long total = (long long)B * C * (HW / 4);
    k_sum_mix<4><<<grid_for(total), 256, 0, st>>>(x_t, gt, mask, noise, ts, table, x_out, C,
                                                   HW / 4, total);
  } else {
    long long total = (long long)B * C * HW;
    k_sum_mix<1><<<grid_for(total), 256, 0, st>>>(x_t, gt, mask, noise, ts, table, x_out, C, HW,
                                                   total);
  }
  EO_CHECK_LAUNCH();
  return EO_OK;
}

template <bool MIX>
static int launch_step(const float* x_t, const float* eps, const float* noise,
                       const int64_t* timesteps, const float* gt, const float* mask,
                       const float* noise_next, const int64_t* timesteps_next,
                       const float* table, float* x_out, int B, int C, int HW, int clip,
                       int all_pos, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  bool v4 = (HW % 4 == 0) && aligned16(x_t) && aligned16(eps) && aligned16(noise) &&
            aligned16(x_out);
  if (MIX) v4 = v4 && aligned16(gt) && aligned16(mask) && aligned16(noise_next);
  const long long* ts = reinterpret_cast<const long long*>(timesteps);
  const long long* tn = reinterpret_cast<const long long*>(timesteps_next);
#define EO_LAUNCH_STEP(VEC, CLIP)                                                          \
  do {                                                                                     \
    long long total = (long long)B * C * (HW / VEC);                                       \
    k_step<VEC, CLIP, MIX><<<grid_for(total), 256, 0, st>>>(                               \
        x_t, eps, noise, ts, gt, mask, noise_next, tn, table, x_out, C, HW / VEC, total,   \
        all_pos);                                                                          \
  } while (0)
  if (v4) { if (clip) EO_LAUNCH_STEP(4, true); else EO_LAUNCH_STEP(4, false); }
  else    { if (clip) EO_LAUNCH_STEP(1, true); else EO_LAUNCH_STEP(1, false); }
#undef EO_LAUNCH_STEP
  EO_CHECK_LAUNCH();
  return EO_OK;
}

extern "C" int eo_ddpm_step(const float* x_t, const float* eps, const float* noise,
                            const int64_t* timesteps, const float* table, float* x_out, int B,
                            int C, int HW, int clip, int all_t_positive, void* stream) {
  EO_REQUIRE(x_t && eps && noise && timesteps && table && x_out, EO_ERR_ARG,
             "eo_ddpm_step: null pointer");
  EO_REQUIRE(B > 0 && C > 0 && HW > 0, EO_ERR_ARG, "eo_ddpm_step: bad shape");
  return launch_step<false>(x_t, eps, noise, timesteps, nullptr, nullptr, nullptr, nullptr,
                            table, x_out, B, C, HW, clip, all_t_positive, stream);
}

extern "C" int eo_ddpm_step_mix(const float* x_t, const float* eps, const float* noise,
                                const int64_t* timesteps, const float* gt, const float* mask,
                                const float* noise_next, const int64_t* timesteps_next,
                                const float* table, float* x_out, int B, int C, int HW, int clip,
                                int all_t_positive, void* stream) {
  EO_REQUIRE(x_t && eps && noise && timesteps && gt && mask && noise_next && timesteps_next &&
                 table && x_out, EO_ERR_ARG, "eo_ddpm_step_mix: null pointer");
  EO_REQUIRE(B > 0 && C > 0 && HW > 0, EO_ERR_ARG, "eo_ddpm_step_mix: bad shape");
  return launch_step<true>(x_t, eps, noise, timesteps, gt, mask, noise_next, timesteps_next,
                           table, x_out, B, C, HW, clip, all_t_positive, stream);
}

extern "C" int eo_ddim_step(const float* x, const float* e_t, const float* noise, float* x_prev,
                            float* pred_x0, float sqrt_a_t, float sqrt_1m_a_t, float sqrt_a_prev,
                            float dir_coef, float sigma_t, float temperature, int64_t n_elems,
                            void* stream) {
  EO_REQUIRE(x && e_t && x_prev && pred_x0, EO_ERR_ARG, "eo_ddim_step: null pointer");
  EO_REQUIRE(n_elems > 0, EO_ERR_ARG, "eo_ddim_step: bad size");
  EO_REQUIRE(noise || sigma_t == 0.0f, EO_ERR_ARG, "eo_ddim_step: noise required when sigma_t != 0");
  cudaStream_t st = (cudaStream_t)stream;
  bool v4 = (n_elems % 4 == 0) && aligned16(x) && aligned16(e_t) && aligned16(x_prev) &&
            aligned16(pred_x0) && (!noise || aligned16(noise));
  if (v4)
    k_ddim_step<4><<<grid_for(n_elems / 4), 256, 0, st>>>(x, e_t, noise, x_prev, pred_x0, sqrt_a_t,
                                                          sqrt_1m_a_t, sqrt_a_prev, dir_coef,
                                                          sigma_t, temperature, n_elems / 4);
  else
    k_ddim_step<1><<<grid_for(n_elems), 256, 0, st>>>(x, e_t, noise, x_prev, pred_x0, sqrt_a_t,
                                                      sqrt_1m_a_t, sqrt_a_prev, dir_coef, sigma_t,
                                                      temperature, n_elems);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

extern "C" int eo_cfg_combine(const float* e_uncond, const float* e_cond, float scale,
                              float* e_out, int64_t n_elems, void* stream) {
  EO_REQUIRE(e_uncond && e_cond && e_out && n_elems > 0, EO_ERR_ARG, "eo_cfg_combine: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  bool v4 = (n_elems % 4 == 0) && aligned16(e_uncond) && aligned16(e_cond) && aligned16(e_out);
  if (v4)
    k_cfg_combine<4><<<grid_for(n_elems / 4), 256, 0, st>>>(e_uncond, e_cond, scale, e_out,
                                                            n_elems / 4);
  else
    k_cfg_combine<1><<<grid_for(n_elems), 256, 0, st>>>(e_uncond, e_cond, scale, e_out, n_elems);
  EO_CHECK_LAUNCH();
  return EO_OK;
}
