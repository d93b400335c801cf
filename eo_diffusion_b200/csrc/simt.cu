// CUDA-core kernels of libeo_b200:
//   * the fp32 parity-mode path (generic implicit-GEMM conv, attention),
//   * the HBM-bound kernels shared by both modes (GroupNorm statistics / apply, timestep
//     embedding linears, layout pre-passes, the small-N output conv),
//   * weight packing.
// Reference semantics are cited per kernel (paths under the reference repo root).
#include "kernels.h"

namespace eo {

// =======================================================================================
// generic SIMT implicit-GEMM convolution
//   reference: nn.Conv2d via conv_nd (backbones/unet_openai.py:16-26) with the producers'
//   GroupNorm32+SiLU (:11-13, :314) folded into the operand load, th.cat (:773) as a
//   second K segment, F.interpolate nearest x2 (:236) / stride 2 (:262-265) as index maps,
//   and the `h + emb_out` (:382), `skip_connection(x) + h` (:385), `x + h` (:433) adds in
//   the epilogue.
// =======================================================================================
namespace {
constexpr int BM = 64, BN = 64, BK = 16;

__device__ __forceinline__ float load_elem(const ConvSrc& s, long long idx) {
  if (s.dt == DT_F32) return reinterpret_cast<const float*>(s.ptr)[idx];
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(s.ptr)[idx]);
}

__global__ void __launch_bounds__(256) k_conv_simt(const ConvSimtParams p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long M = (long long)p.B * p.Hout * p.Wout;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // operand-load roles
  const int lm = tid >> 2, kq = (tid & 3) * 4;
  const long long gm = m0 + lm;
  int pb = 0, poh = 0, pow_ = 0;
  const bool m_ok = gm < M;
  if (m_ok) {
    pb = (int)(gm / ((long long)p.Hout * p.Wout));
    int r = (int)(gm - (long long)pb * p.Hout * p.Wout);
    poh = r / p.Wout;
    pow_ = r - poh * p.Wout;
  }
  const int kr = tid >> 4, nc = (tid & 15) * 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int s = 0; s < p.nsrc; ++s) {
    const ConvSrc& src = p.src[s];
    const int C = src.C;
    const int KK = src.ksize * src.ksize * C;
    const bool uniform = (C % BK) == 0;
    for (int k0 = 0; k0 < KK; k0 += BK) {
      // ---- A tile: 64 pixels x 16 k -------------------------------------------------
      float av[4] = {0.f, 0.f, 0.f, 0.f};
      if (m_ok) {
        int tap_u = 0, c_u = 0;
        if (uniform) { tap_u = k0 / C; c_u = k0 - tap_u * C + kq; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int k = k0 + kq + j;
          if (k >= KK) continue;
          int tap, c;
          if (uniform) { tap = tap_u; c = c_u + j; }
          else { tap = k / C; c = k - tap * C; }
          int ih, iw;
          bool ok = true;
          if (src.ksize == 3) {
            int dh = tap / 3 - 1, dw = tap - (tap / 3) * 3 - 1;
            if (p.up) {
              int uh = poh + dh, uw = pow_ + dw;
              ok = uh >= 0 && uh < 2 * p.Hin && uw >= 0 && uw < 2 * p.Win;
              ih = uh >> 1; iw = uw >> 1;
            } else {
              ih = poh * p.stride + dh; iw = pow_ * p.stride + dw;
              ok = ih >= 0 && ih < p.Hin && iw >= 0 && iw < p.Win;
            }
          } else { ih = poh; iw = pow_; }
          if (ok) {
            long long idx = src.nchw
                ? (((long long)pb * C + c) * p.Hin + ih) * p.Win + iw
                : (((long long)pb * p.Hin + ih) * p.Win + iw) * C + c;
            float v = load_elem(src, idx);
            if (src.gn_scale) {
              v = v * src.gn_scale[(long long)pb * src.gn_ld + c] + src.gn_shift[(long long)pb * src.gn_ld + c];
            }
            if (src.silu) v = silu_acc(v);
            av[j] = v;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) As[kq + j][lm] = av[j];
      // ---- B tile: 16 k x 64 n --------------------------------------------------------
      {
        int k = k0 + kr;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int n = n0 + nc + j;
          Bs[kr][nc + j] = (k < KK && n < p.Cout)
              ? __ldg(p.W + (long long)(src.w_off + k) * p.Cout + n) : 0.f;
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue -----------------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    int b = (int)(m / ((long long)p.Hout * p.Wout));
    int r = (int)(m - (long long)b * p.Hout * p.Wout);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= p.Cout) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      if (p.bias_nc) v += p.bias_nc[(long long)b * p.ld_bias_nc + n];
      long long o = m * p.Cout + n;
      if (p.residual) {
        v += (p.out_dt == DT_F32)
            ? reinterpret_cast<const float*>(p.residual)[o]
            : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.residual)[o]);
      }
      if (p.out_nchw) {
        reinterpret_cast<float*>(p.out)[((long long)b * p.Cout + n) * p.Hout * p.Wout + r] = v;
      } else if (p.out_dt == DT_F32) {
        reinterpret_cast<float*>(p.out)[o] = v;
      } else {
        reinterpret_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16_rn(v);
      }
    }
  }
}
}  // namespace

int launch_conv_simt(const ConvSimtParams& p, cudaStream_t st) {
  EO_REQUIRE(p.nsrc >= 1 && p.nsrc <= 3, EO_ERR_ARG, "conv_simt: nsrc");
  EO_REQUIRE(!(p.up && p.stride != 1), EO_ERR_ARG, "conv_simt: up with stride");
  for (int s = 0; s < p.nsrc; ++s) {
    EO_REQUIRE(p.src[s].ksize == 3 || (p.src[s].ksize == 1 && p.stride == 1 && !p.up),
               EO_ERR_ARG, "conv_simt: 1x1 source with stride/up");
  }
  long long M = (long long)p.B * p.Hout * p.Wout;
  dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(p.Cout, BN));
  k_conv_simt<<<grid, 256, 0, st>>>(p);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

// =======================================================================================
// output conv for small Cout (bf16 mode): out = conv3x3(SiLU(GN(x))) , NHWC bf16 -> NCHW fp32
//   reference: UNetModel.out (unet_openai.py:739-743, :780)
// one thread per output pixel, all Cout accumulators in registers; weights and the sample's
// GN scale/shift staged in shared memory; 16-byte (8 x bf16) loads along C.
// =======================================================================================
namespace {
template <int NACC>
__global__ void __launch_bounds__(128) k_conv_small_n(const ConvSmallNParams p) {
  extern __shared__ float smem[];
  float* wsm = smem;                       // [9*C][Cout]
  float* sc = smem + 9 * p.C * p.Cout;     // [C]
  float* sh = sc + p.C;                    // [C]
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 9 * p.C * p.Cout; i += blockDim.x) wsm[i] = p.Wp[i];
  for (int i = threadIdx.x; i < p.C; i += blockDim.x) {
    sc[i] = p.gn_scale[(long long)b * p.C + i];
    sh[i] = p.gn_shift[(long long)b * p.C + i];
  }
  __syncthreads();
  const int HW = p.H * p.W;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  const int oh = pix / p.W, ow = pix - oh * p.W;
  float acc[NACC];
#pragma unroll
  for (int n = 0; n < NACC; ++n) acc[n] = 0.f;
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(p.x);
  for (int tap = 0; tap < 9; ++tap) {
    int ih = oh + tap / 3 - 1, iw = ow + tap % 3 - 1;
    if (ih < 0 || ih >= p.H || iw < 0 || iw >= p.W) continue;
    const uint4* row = reinterpret_cast<const uint4*>(x + (((long long)b * p.H + ih) * p.W + iw) * p.C);
    const float* wt = wsm + tap * p.C * p.Cout;
    for (int c8 = 0; c8 < p.C / 8; ++c8) {
      uint4 q = __ldg(row + c8);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f = __bfloat1622float2(h2[e]);
        int c = c8 * 8 + e * 2;
        float v0 = silu_f(f.x * sc[c] + sh[c]);
        float v1 = silu_f(f.y * sc[c + 1] + sh[c + 1]);
#pragma unroll
        for (int n = 0; n < NACC; ++n) {
          if (n < p.Cout) {
            acc[n] = fmaf(v0, wt[c * p.Cout + n], acc[n]);
            acc[n] = fmaf(v1, wt[(c + 1) * p.Cout + n], acc[n]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int n = 0; n < NACC; ++n)
    if (n < p.Cout) p.out_nchw[((long long)b * p.Cout + n) * HW + pix] = acc[n] + p.bias[n];
}
}  // namespace

int launch_conv_small_n(const ConvSmallNParams& p, cudaStream_t st) {
  EO_REQUIRE(p.dt == DT_BF16 && p.C % 8 == 0 && p.Cout <= 16, EO_ERR_ARG, "conv_small_n: shape");
  size_t smem = (size_t)(9 * p.C * p.Cout + 2 * p.C) * sizeof(float);
  EO_REQUIRE(smem <= 200 * 1024, EO_ERR_ARG, "conv_small_n: weights do not fit shared memory");
  dim3 grid((unsigned)ceil_div(p.H * p.W, 128), (unsigned)p.B);
  if (p.Cout <= 4) {
    EO_CHECK_CUDA(cudaFuncSetAttribute(k_conv_small_n<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_conv_small_n<4><<<grid, 128, smem, st>>>(p);
  } else {
    EO_CHECK_CUDA(cudaFuncSetAttribute(k_conv_small_n<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_conv_small_n<16><<<grid, 128, smem, st>>>(p);
  }
  EO_CHECK_LAUNCH();
  return EO_OK;
}

// =======================================================================================
// GroupNorm32 (unet_openai.py:11-13; nn.GroupNorm(32, C), eps 1e-5, biased variance)
// =======================================================================================
namespace {
struct GnSrcs { GnSrc s[2]; int n; };

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float v[4]) {
    float4 q = *reinterpret_cast<const float4*>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float v[4]) {
    uint2 q = *reinterpret_cast<const uint2*>(p);
    float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&q.x));
    float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&q.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
};

// grid (chunks, B); each CTA reduces `ppc` pixels of sample b over all channels of all
// sources into 32 (sum, sumsq) pairs, then adds them to sums[b] with 64 double atomics.
template <typename T, typename Acc>
__global__ void __launch_bounds__(256)
k_gn_stats(GnSrcs srcs, int HW, int ppc, int cpg, double* __restrict__ sums) {
  __shared__ double sm[64];
  const int tid = threadIdx.x, b = blockIdx.y;
  const int p0 = blockIdx.x * ppc, p1 = min(HW, p0 + ppc);
  if (tid < 64) sm[tid] = 0.0;
  __syncthreads();
  int coff = 0;
  for (int s = 0; s < srcs.n; ++s) {
    const int C = srcs.s[s].C;
    const T* base = reinterpret_cast<const T*>(srcs.s[s].ptr) + (long long)b * HW * C;
    const int ncol = C / 4;
    const int rows = blockDim.x / ncol;   // >= 1 (C <= 1024)
    const int col = tid % ncol, row = tid / ncol;
    if (row < rows) {
      Acc s4[4] = {0, 0, 0, 0}, q4[4] = {0, 0, 0, 0};
      for (int p = p0 + row; p < p1; p += rows) {
        float v[4];
        Vec4<T>::load(base + (long long)p * C + col * 4, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) { s4[j] += (Acc)v[j]; q4[j] += (Acc)v[j] * (Acc)v[j]; }
      }
      const int c0 = coff + col * 4;
      if ((cpg & 3) == 0) {
        int g = c0 / cpg;
        atomicAdd(&sm[2 * g], (double)s4[0] + (double)s4[1] + (double)s4[2] + (double)s4[3]);
        atomicAdd(&sm[2 * g + 1], (double)q4[0] + (double)q4[1] + (double)q4[2] + (double)q4[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int g = (c0 + j) / cpg;
          atomicAdd(&sm[2 * g], (double)s4[j]);
          atomicAdd(&sm[2 * g + 1], (double)q4[j]);
        }
      }
    }
    coff += C;
  }
  __syncthreads();
  if (tid < 64) atomicAdd(&sums[(long long)b * 64 + tid], sm[tid]);
}

__global__ void k_gn_finalize(const double* __restrict__ sums, const float* __restrict__ gamma,
                              const float* __restrict__ beta, int B, int C, int HW,
                              float* __restrict__ scale, float* __restrict__ shift) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  int b = i / C, c = i - b * C;
  int cpg = C / 32, g = c / cpg;
  double n = (double)cpg * (double)HW;
  double mean = sums[(long long)b * 64 + 2 * g] / n;
  double var = sums[(long long)b * 64 + 2 * g + 1] / n - mean * mean;
  if (var < 0.0) var = 0.0;
  double rstd = 1.0 / sqrt(var + 1e-5);
  float sc = (float)(rstd * (double)gamma[c]);
  scale[i] = sc;
  shift[i] = (float)((double)beta[c] - mean * rstd * (double)gamma[c]);
}

// bf16 -> bf16, 8 channels (16 bytes) per thread per pixel
__global__ void __launch_bounds__(256)
k_gn_apply(GnSrcs srcs, int HW, int ppc, int Ctot, const float* __restrict__ scale,
           const float* __restrict__ shift, int silu, __nv_bfloat16* __restrict__ dst) {
  const int tid = threadIdx.x, b = blockIdx.y;
  const int p0 = blockIdx.x * ppc, p1 = min(HW, p0 + ppc);
  int coff = 0;
  for (int s = 0; s < srcs.n; ++s) {
    const int C = srcs.s[s].C;
    const __nv_bfloat16* base =
        reinterpret_cast<const __nv_bfloat16*>(srcs.s[s].ptr) + (long long)b * HW * C;
    const int ncol = C / 8;
    const int rows = blockDim.x / ncol;
    const int col = tid % ncol, row = tid / ncol;
    if (row < rows) {
      float sc[8], sh[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sc[j] = scale[(long long)b * Ctot + coff + col * 8 + j];
        sh[j] = shift[(long long)b * Ctot + coff + col * 8 + j];
      }
      for (int p = p0 + row; p < p1; p += rows) {
        uint4 q = *reinterpret_cast<const uint4*>(base + (long long)p * C + col * 8);
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f = __bfloat1622float2(h2[e]);
          f.x = f.x * sc[2 * e] + sh[2 * e];
          f.y = f.y * sc[2 * e + 1] + sh[2 * e + 1];
          if (silu) { f.x = silu_f(f.x); f.y = silu_f(f.y); }
          h2[e] = __floats2bfloat162_rn(f.x, f.y);
        }
        *reinterpret_cast<uint4*>(dst + ((long long)b * HW + p) * Ctot + coff + col * 8) = q;
      }
    }
    coff += C;
  }
}

// pixels per CTA so the grid is a few waves of the 148 SMs
inline int pick_ppc(int B, int HW) {
  long long target = (long long)num_sms() * 8;
  long long ppc = ceil_div((long long)B * HW, target);
  if (ppc < 32) ppc = 32;
  if (ppc > HW) ppc = HW;
  return (int)ppc;
}
}  // namespace

int launch_gn_stats(const GnSrc* src, int nsrc, int dt, int B, int HW, double* sums,
                    cudaStream_t st) {
  EO_REQUIRE(nsrc >= 1 && nsrc <= 2, EO_ERR_ARG, "gn_stats: nsrc");
  GnSrcs s; s.n = nsrc;
  int Ctot = 0;
  for (int i = 0; i < nsrc; ++i) {
    s.s[i] = src[i];
    EO_REQUIRE(src[i].C % 4 == 0 && src[i].C <= 1024 && src[i].C > 0, EO_ERR_ARG,
               "gn_stats: channels must be a multiple of 4 and <= 1024 (got %d)", src[i].C);
    Ctot += src[i].C;
  }
  EO_REQUIRE(Ctot % 32 == 0, EO_ERR_ARG, "gn_stats: channels %d not divisible by 32", Ctot);
  int ppc = pick_ppc(B, HW);
  dim3 grid((unsigned)ceil_div(HW, ppc), (unsigned)B);
  if (dt == DT_F32) k_gn_stats<float, double><<<grid, 256, 0, st>>>(s, HW, ppc, Ctot / 32, sums);
  else k_gn_stats<__nv_bfloat16, float><<<grid, 256, 0, st>>>(s, HW, ppc, Ctot / 32, sums);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

int launch_gn_finalize(const double* sums, const float* gamma, const float* beta, int B, int C,
                       int HW, float* scale, float* shift, cudaStream_t st) {
  k_gn_finalize<<<(unsigned)ceil_div((long long)B * C, 256), 256, 0, st>>>(sums, gamma, beta, B, C,
                                                                          HW, scale, shift);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

int launch_gn_apply(const GnSrc* src, int nsrc, int B, int HW, const float* scale,
                    const float* shift, int silu, void* dst, cudaStream_t st) {
  EO_REQUIRE(nsrc >= 1 && nsrc <= 2, EO_ERR_ARG, "gn_apply: nsrc");
  GnSrcs s; s.n = nsrc;
  int Ctot = 0;
  for (int i = 0; i < nsrc; ++i) {
    s.s[i] = src[i];
    EO_REQUIRE(src[i].C % 8 == 0 && src[i].C <= 2048, EO_ERR_ARG, "gn_apply: channels %% 8");
    Ctot += src[i].C;
  }
  int ppc = pick_ppc(B, HW);
  dim3 grid((unsigned)ceil_div(HW, ppc), (unsigned)B);
  k_gn_apply<<<grid, 256, 0, st>>>(s, HW, ppc, Ctot, scale, shift, silu,
                                   reinterpret_cast<__nv_bfloat16*>(dst));
  EO_CHECK_LAUNCH();
  return EO_OK;
}

// =======================================================================================
// timestep embedding (unet_openai.py:81-99) and nn.Linear layers of time_embed (:597-602)
// and ResBlock.emb_layers (:333-339)
// =======================================================================================
namespace {
__global__ void k_sinusoid(const long long* __restrict__ t, const float* __restrict__ freqs,
                           int B, int half, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  int b = i / half, j = i - b * half;
  float arg = (float)t[b] * freqs[j];   // timesteps[:, None].float() * freqs[None]
  out[(long long)b * 2 * half + j] = cosf(arg);
  out[(long long)b * 2 * half + half + j] = sinf(arg);
}

// one warp per output element
__global__ void __launch_bounds__(256)
k_linear(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ bias,
         const float* __restrict__ bias2, const float* __restrict__ emb_rows,
         const long long* __restrict__ idx, int silu_in, int B, int K, int N,
         float* __restrict__ out) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= B * N) return;
  int b = warp / N, n = warp - b * N;
  const float* x = in + (long long)b * K;
  const float* w = W + (long long)n * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) {
    float v = x[k];
    if (silu_in) v = silu_acc(v);
    acc = fmaf(v, w[k], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    if (bias) acc += bias[n];
    if (bias2) acc += bias2[n];
    if (emb_rows) acc += emb_rows[idx[b] * N + n];
    out[(long long)b * N + n] = acc;
  }
}
}  // namespace

int launch_sinusoid(const int64_t* t, const float* freqs, int B, int half, float* out,
                    cudaStream_t st) {
  k_sinusoid<<<(unsigned)ceil_div((long long)B * half, 128), 128, 0, st>>>(
      reinterpret_cast<const long long*>(t), freqs, B, half, out);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

int launch_linear(const float* in, const float* W, const float* bias, const float* bias2,
                  const float* emb_rows, const int64_t* idx, int silu_in, int B, int K, int N,
                  float* out, cudaStream_t st) {
  long long warps = (long long)B * N;
  k_linear<<<(unsigned)ceil_div(warps * 32, 256), 256, 0, st>>>(
      in, W, bias, bias2, emb_rows, reinterpret_cast<const long long*>(idx), silu_in, B, K, N, out);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

// =======================================================================================
// attention, fp32 SIMT: QKVAttentionLegacy.forward (unet_openai.py:465-481) /
// QKVAttention.forward (:497-515).  One thread per query row, online softmax in fp32,
// K/V tiles of 64 keys staged in shared memory (all threads read the same key -> broadcast).
// =======================================================================================
namespace {
constexpr int AT_Q = 128, AT_K = 64, AT_D = 64;

__global__ void __launch_bounds__(AT_Q)
k_attention_simt(const float* __restrict__ qkv, float* __restrict__ out, int T, int heads, int ch,
                 int ld, int head_stride, int part_stride, float scale) {
  __shared__ float Ks[AT_K][AT_D];
  __shared__ float Vs[AT_K][AT_D];
  const int b = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * AT_Q + threadIdx.x;
  const float* base = qkv + (long long)b * T * ld + (long long)h * head_stride;
  float q[AT_D], o[AT_D];
#pragma unroll
  for (int d = 0; d < AT_D; ++d) { q[d] = 0.f; o[d] = 0.f; }
  if (qi < T) {
    for (int d = 0; d < ch; ++d) q[d] = base[(long long)qi * ld + d] * scale;   // q * scale
  }
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < T; k0 += AT_K) {
    __syncthreads();
    for (int i = threadIdx.x; i < AT_K * AT_D; i += AT_Q) {
      int kk = i / AT_D, d = i - kk * AT_D;
      int kt = k0 + kk;
      float kv = 0.f, vv = 0.f;     // zero fill: the dot products run over all AT_D lanes
      if (kt < T && d < ch) {
        kv = base[(long long)kt * ld + part_stride + d] * scale;               // k * scale
        vv = base[(long long)kt * ld + 2 * part_stride + d];
      }
      Ks[kk][d] = kv; Vs[kk][d] = vv;
    }
    __syncthreads();
    const int nk = min(AT_K, T - k0);
    for (int c0 = 0; c0 < nk; c0 += 8) {
      float s[8];
      float cm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = -INFINITY;
        if (c0 + j < nk) {
          a = 0.f;
#pragma unroll
          for (int d = 0; d < AT_D; ++d) a = fmaf(q[d], Ks[c0 + j][d], a);
        }
        s[j] = a;
        cm = fmaxf(cm, a);
      }
      float mn = fmaxf(m, cm);
      float alpha = expf(m - mn);     // m == -inf on the first chunk -> 0
      l *= alpha;
#pragma unroll
      for (int d = 0; d < AT_D; ++d) o[d] *= alpha;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (c0 + j < nk) {
          float pj = expf(s[j] - mn);
          l += pj;
#pragma unroll
          for (int d = 0; d < AT_D; ++d) o[d] = fmaf(pj, Vs[c0 + j][d], o[d]);
        }
      }
      m = mn;
    }
  }
  if (qi < T) {
    float inv = 1.0f / l;
    float* op = out + ((long long)b * T + qi) * (heads * ch) + h * ch;
    for (int d = 0; d < ch; ++d) op[d] = o[d] * inv;
  }
}
}  // namespace

int launch_attention_simt(const float* qkv, float* out, int B, int T, int heads, int ch, int ld,
                          int head_stride, int part_stride, cudaStream_t st) {
  EO_REQUIRE(ch <= AT_D, EO_ERR_ARG,
             "attention: head dimension %d > %d is not supported by this build", ch, AT_D);
  float scale = 1.0f / sqrtf(sqrtf((float)ch));
  dim3 grid((unsigned)ceil_div(T, AT_Q), (unsigned)heads, (unsigned)B);
  k_attention_simt<<<grid, AT_Q, 0, st>>>(qkv, out, T, heads, ch, ld, head_stride, part_stride,
                                           scale);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

// =======================================================================================
// layout pre-passes (bf16, 16-byte vectors)
// =======================================================================================
namespace {
// F.interpolate(scale_factor=2, mode="nearest") (unet_openai.py:236)
__global__ void __launch_bounds__(256)
k_upsample2x(const uint4* __restrict__ src, uint4* __restrict__ dst, int H, int W, int C8,
             long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C8);
    long long r = i / C8;
    int ow = (int)(r % (2 * W)); r /= (2 * W);
    int oh = (int)(r % (2 * H));
    long long b = r / (2 * H);
    dst[i] = __ldg(src + ((b * H + (oh >> 1)) * W + (ow >> 1)) * C8 + c);
  }
}
// stride-2 conv operand regrouping: 4 parity planes so every tap is a unit-stride window
__global__ void __launch_bounds__(256)
k_space_to_depth(const uint4* __restrict__ src, uint4* __restrict__ dst, int Bstride, int H, int W,
                 int C8, long long total) {
  const int H2 = H / 2, W2 = W / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C8);
    long long r = i / C8;
    int w = (int)(r % W); r /= W;
    int h = (int)(r % H);
    long long b = r / H;
    int plane = (h & 1) * 2 + (w & 1);
    dst[((((long long)plane * Bstride + b) * H2 + (h >> 1)) * W2 + (w >> 1)) * C8 + c] = __ldg(src + i);
  }
}
__global__ void __launch_bounds__(256)
k_nhwc_to_nchw(const void* __restrict__ src, int dt, float* __restrict__ dst, int HW, int C,
               long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // i indexes dst (NCHW)
    int p = (int)(i % HW);
    long long r = i / HW;
    int c = (int)(r % C);
    long long b = r / C;
    long long s = (b * HW + p) * C + c;
    dst[i] = dt == DT_F32 ? reinterpret_cast<const float*>(src)[s]
                          : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[s]);
  }
}
inline unsigned ew_grid(long long total) {
  long long need = ceil_div(total, 256), cap = (long long)num_sms() * 16;
  return (unsigned)(need < cap ? (need < 1 ? 1 : need) : cap);
}
}  // namespace

int launch_upsample2x(const void* src, void* dst, int B, int H, int W, int C, cudaStream_t st) {
  EO_REQUIRE(C % 8 == 0, EO_ERR_ARG, "upsample2x: C %% 8");
  long long total = (long long)B * 4 * H * W * (C / 8);
  k_upsample2x<<<ew_grid(total), 256, 0, st>>>(reinterpret_cast<const uint4*>(src),
                                               reinterpret_cast<uint4*>(dst), H, W, C / 8, total);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

int launch_space_to_depth(const void* src, void* dst, int B, int Bstride, int H, int W, int C,
                          cudaStream_t st) {
  EO_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, EO_ERR_ARG, "space_to_depth: shape");
  long long total = (long long)B * H * W * (C / 8);
  k_space_to_depth<<<ew_grid(total), 256, 0, st>>>(reinterpret_cast<const uint4*>(src),
                                                   reinterpret_cast<uint4*>(dst), Bstride, H, W, C / 8,
                                                   total);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

int launch_nhwc_to_nchw_f32(const void* src, int dt, float* dst, int B, int HW, int C,
                            cudaStream_t st) {
  long long total = (long long)B * HW * C;
  k_nhwc_to_nchw<<<ew_grid(total), 256, 0, st>>>(src, dt, dst, HW, C, total);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

// =======================================================================================
// weight packing
// =======================================================================================
namespace {
__global__ void k_pack_conv_weight(const float* __restrict__ w, int Cin_total, int ksize,
                                   int cin_off, int C, void* __restrict__ dst, int dst_dt,
                                   long long stride_n, long long stride_k, int k_off, int Nout,
                                   const int* __restrict__ row_map) {
  const int taps = ksize * ksize;
  long long total = (long long)Nout * taps * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long r = i / C;
    int tap = (int)(r % taps);
    int n = (int)(r / taps);
    int row = row_map ? row_map[n] : n;
    float v = 0.f;
    if (row >= 0) v = w[((long long)row * Cin_total + cin_off + c) * taps + tap];
    long long o = (long long)n * stride_n + (long long)(k_off + tap * C + c) * stride_k;
    if (dst_dt == DT_F32) reinterpret_cast<float*>(dst)[o] = v;
    else reinterpret_cast<__nv_bfloat16*>(dst)[o] = __float2bfloat16_rn(v);
  }
}
__global__ void k_pack_bias(const float* a, const float* b, float* dst, int Nout,
                            const int* row_map) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Nout) return;
  int row = row_map ? row_map[n] : n;
  float v = 0.f;
  if (row >= 0) { if (a) v += a[row]; if (b) v += b[row]; }
  dst[n] = v;
}
}  // namespace

int launch_pack_conv_weight(const float* w, int Cin_total, int ksize, int cin_off, int C,
                            void* dst, int dst_dt, long long stride_n, long long stride_k,
                            int k_off, int Nout, const int* row_map, cudaStream_t st) {
  long long total = (long long)Nout * ksize * ksize * C;
  k_pack_conv_weight<<<ew_grid(total), 256, 0, st>>>(w, Cin_total, ksize, cin_off, C, dst, dst_dt,
                                                     stride_n, stride_k, k_off, Nout, row_map);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

int launch_pack_bias(const float* a, const float* b, float* dst, int Nout, const int* row_map,
                     cudaStream_t st) {
  k_pack_bias<<<(unsigned)ceil_div(Nout, 128), 128, 0, st>>>(a, b, dst, Nout, row_map);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

}  // namespace eo
