// CUDA-core kernels of libeo_b200:
//   * the fp32 parity-mode path (generic implicit-GEMM conv, attention),
//   * the HBM-bound kernels shared by both modes (GroupNorm statistics / apply, timestep
//     embedding linears, layout pre-passes, the small-N output conv),
//   * weight packing.
// Reference semantics are cited per kernel (paths under the reference repo root).
#include "kernels.h"
#include <algorithm>

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
constexpr int BK = 16;

__device__ __forceinline__ float load_elem(const ConvSrc& s, long long idx) {
  if (s.dt == DT_F32) return reinterpret_cast<const float*>(s.ptr)[idx];
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(s.ptr)[idx]);
}

// BM x BN output tile per CTA (256 threads as 16 x 16, TM x TN outputs each), K in steps of 16.
// <64,64>: small layers / narrow outputs; <128,128>: 64 FFMA per 4 LDS.128 for the big ones.
template <int BM, int BN>
__global__ void __launch_bounds__(256) k_conv_simt(const ConvSimtParams p) {
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int AK = BM * BK / 256;      // k values of one pixel loaded per thread (4 or 8)
  constexpr int TPP = BK / AK;           // threads per pixel
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long M = (long long)p.B * p.Hout * p.Wout;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // operand-load roles
  const int lm = tid / TPP, kq = (tid % TPP) * AK;
  const long long gm = m0 + lm;
  int pb = 0, poh = 0, pow_ = 0;
  const bool m_ok = gm < M;
  if (m_ok) {
    pb = (int)(gm / ((long long)p.Hout * p.Wout));
    int r = (int)(gm - (long long)pb * p.Hout * p.Wout);
    poh = r / p.Wout;
    pow_ = r - poh * p.Wout;
  }
  const int kr = tid >> 4, nc = (tid & 15) * TN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int s = 0; s < p.nsrc; ++s) {
    const ConvSrc& src = p.src[s];
    const int C = src.C;
    const int KK = src.ksize * src.ksize * C;
    const bool uniform = (C % BK) == 0;
    for (int k0 = 0; k0 < KK; k0 += BK) {
      // ---- A tile: BM pixels x 16 k -------------------------------------------------
      float av[AK];
#pragma unroll
      for (int j = 0; j < AK; ++j) av[j] = 0.f;
      if (m_ok) {
        int tap_u = 0, c_u = 0;
        if (uniform) { tap_u = k0 / C; c_u = k0 - tap_u * C + kq; }
#pragma unroll
        for (int j = 0; j < AK; ++j) {
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
      for (int j = 0; j < AK; ++j) As[kq + j][lm] = av[j];
      // ---- B tile: 16 k x BN n --------------------------------------------------------
      {
        int k = k0 + kr;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          int n = n0 + nc + j;
          Bs[kr][nc + j] = (k < KK && n < p.Cout)
              ? __ldg(p.W + (long long)(src.w_off + k) * p.Cout + n) : 0.f;
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; i += 4)
          *reinterpret_cast<float4*>(&a[i]) = *reinterpret_cast<const float4*>(&As[kk][ty * TM + i]);
#pragma unroll
        for (int j = 0; j < TN; j += 4)
          *reinterpret_cast<float4*>(&b[j]) = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN + j]);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue -----------------------------------------------------------------------
  const bool vec = !p.out_nchw && (p.Cout % TN) == 0;     // TN contiguous channels per pixel: vector stores
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    long long m = m0 + ty * TM + i;
    if (m >= M) continue;
    int b = (int)(m / ((long long)p.Hout * p.Wout));
    int r = (int)(m - (long long)b * p.Hout * p.Wout);
    const int nb = n0 + tx * TN;
    if (vec) {
      if (nb >= p.Cout) continue;
      float v[TN];
      const long long o = m * p.Cout + nb;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        v[j] = acc[i][j];
        if (p.bias) v[j] += p.bias[nb + j];
        if (p.bias_nc) v[j] += p.bias_nc[(long long)b * p.ld_bias_nc + nb + j];
      }
      if (p.out_dt == DT_F32) {
        if (p.residual) {
#pragma unroll
          for (int j = 0; j < TN; j += 4) {
            const float4 r4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + o + j);
            v[j] += r4.x; v[j + 1] += r4.y; v[j + 2] += r4.z; v[j + 3] += r4.w;
          }
        }
#pragma unroll
        for (int j = 0; j < TN; j += 4)
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
        if (p.residual) {
#pragma unroll
          for (int j = 0; j < TN; j += 2) {
            const float2 r2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(
                reinterpret_cast<const __nv_bfloat16*>(p.residual) + o + j));
            v[j] += r2.x; v[j + 1] += r2.y;
          }
        }
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          uint2 q;
          *reinterpret_cast<__nv_bfloat162*>(&q.x) = __floats2bfloat162_rn(v[j], v[j + 1]);
          *reinterpret_cast<__nv_bfloat162*>(&q.y) = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o + j) = q;
        }
      }
      continue;
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = nb + j;
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
  const bool big = p.Cout > 64 && ceil_div(M, 128) * ceil_div(p.Cout, 128) >= num_sms();
  if (big) {
    dim3 grid((unsigned)ceil_div(M, 128), (unsigned)ceil_div(p.Cout, 128));
    k_conv_simt<128, 128><<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid((unsigned)ceil_div(M, 64), (unsigned)ceil_div(p.Cout, 64));
    k_conv_simt<64, 64><<<grid, 256, 0, st>>>(p);
  }
  EO_CHECK_LAUNCH();
  return EO_OK;
}

// =======================================================================================
// stem conv: conv3x3 over cat(x, cond) given as NCHW fp32 (the reference's input layout), few
// input channels (K = 9*Cin = 27 ... 252), NHWC output.  reference: input_blocks[0]
// (unet_openai.py:608) and th.cat([x, cond], 1) (:754-756).
// One thread = one pixel x 64 output channels: inputs come straight from global memory
// (coalesced along W), the K x 64 weight slice is broadcast from shared memory (LDS.128).
// =======================================================================================
namespace {
template <typename OutT>
__global__ void __launch_bounds__(128) k_conv_stem(const float* __restrict__ x, int Cx,
                                                   const float* __restrict__ cond, int Cc,
                                                   const float* __restrict__ Wp, const float* __restrict__ bias,
                                                   OutT* __restrict__ out, int B, int H, int W, int Cout) {
  extern __shared__ __align__(16) float wsm[];     // [K][64]
  const int K = 9 * (Cx + Cc);
  const int n0 = blockIdx.y * 64;
  for (int i = threadIdx.x; i < K * 64; i += blockDim.x) {
    const int n = n0 + (i & 63);
    wsm[i] = n < Cout ? Wp[(long long)(i >> 6) * Cout + n] : 0.f;
  }
  __syncthreads();
  const long long HW = (long long)H * W;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)B * HW) return;
  const int b = (int)(pix / HW);
  const int r = (int)(pix - (long long)b * HW);
  const int oh = r / W, ow = r - oh * W;
  float acc[64];
#pragma unroll
  for (int n = 0; n < 64; ++n) acc[n] = (n0 + n < Cout) ? bias[n0 + n] : 0.f;
  int kbase = 0;
  for (int s = 0; s < 2; ++s) {
    const float* src = s == 0 ? x : cond;
    const int C = s == 0 ? Cx : Cc;
    if (C == 0) continue;
    for (int tap = 0; tap < 9; ++tap) {
      const int ih = oh + tap / 3 - 1, iw = ow + tap % 3 - 1;
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
        const float* sp = src + ((long long)b * C * H + ih) * W + iw;
        for (int c = 0; c < C; ++c) {
          const float v = __ldg(sp + (long long)c * HW);
          const float4* wr = reinterpret_cast<const float4*>(wsm + (kbase + tap * C + c) * 64);
#pragma unroll
          for (int n4 = 0; n4 < 16; ++n4) {
            const float4 w4 = wr[n4];
            acc[4 * n4] = fmaf(v, w4.x, acc[4 * n4]);
            acc[4 * n4 + 1] = fmaf(v, w4.y, acc[4 * n4 + 1]);
            acc[4 * n4 + 2] = fmaf(v, w4.z, acc[4 * n4 + 2]);
            acc[4 * n4 + 3] = fmaf(v, w4.w, acc[4 * n4 + 3]);
          }
        }
      }
    }
    kbase += 9 * C;
  }
  OutT* op = out + pix * Cout + n0;
  if (n0 + 64 <= Cout && (Cout % 8) == 0) {
    if (sizeof(OutT) == 2) {
#pragma unroll
      for (int n = 0; n < 64; n += 8) {
        uint4 q;
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
        for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(acc[n + 2 * e], acc[n + 2 * e + 1]);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(op) + n) = q;
      }
    } else {
#pragma unroll
      for (int n = 0; n < 64; n += 4)
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(op) + n) = make_float4(acc[n], acc[n + 1], acc[n + 2], acc[n + 3]);
    }
  } else {
#pragma unroll
    for (int n = 0; n < 64; ++n)
      if (n0 + n < Cout) op[n] = from_float<OutT>(acc[n]);
  }
}
}  // namespace

int launch_conv_stem(const float* x, int Cx, const float* cond, int Cc, const float* Wp, const float* bias,
                     void* out, int out_dt, int B, int H, int W, int Cout, cudaStream_t st) {
  const int K = 9 * (Cx + Cc);
  size_t smem = (size_t)K * 64 * sizeof(float);
  EO_REQUIRE(smem <= 200 * 1024, EO_ERR_ARG, "conv_stem: %d input channels do not fit the shared-memory weight slice",
             Cx + Cc);
  dim3 grid((unsigned)ceil_div((long long)B * H * W, 128), (unsigned)ceil_div(Cout, 64));
  if (out_dt == DT_BF16) {
    EO_CHECK_CUDA(cudaFuncSetAttribute(k_conv_stem<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_conv_stem<__nv_bfloat16><<<grid, 128, smem, st>>>(x, Cx, cond, Cc, Wp, bias, reinterpret_cast<__nv_bfloat16*>(out),
                                                       B, H, W, Cout);
  } else {
    EO_CHECK_CUDA(cudaFuncSetAttribute(k_conv_stem<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_conv_stem<float><<<grid, 128, smem, st>>>(x, Cx, cond, Cc, Wp, bias, reinterpret_cast<float*>(out), B, H, W, Cout);
  }
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
__device__ __forceinline__ void unpack8(const uint4& q, float f[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {          // bf16 -> fp32 is a 16-bit shift
    f[2 * e] = __uint_as_float(w[e] << 16);
    f[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
  }
}

// One thread computes TWO horizontally adjacent output pixels x NACC output channels.  The 3x3
// windows of the pair share a 3 x 4 pixel patch (12 instead of 18 16-byte loads per 8 channels);
// the weights sit in shared memory as [tap][c][NACC] so one broadcast LDS.128 feeds 4 output
// channels of both pixels.
template <int NACC>
__global__ void __launch_bounds__(128) k_conv_small_n(const ConvSmallNParams p) {
  extern __shared__ __align__(16) float smem[];
  float* wsm = smem;                        // [9*C][NACC], zero padded
  float* sc = smem + 9 * p.C * NACC;        // [C]
  float* sh = sc + p.C;                     // [C]
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 9 * p.C * NACC; i += blockDim.x) {
    const int n = i % NACC, k = i / NACC;
    wsm[i] = n < p.Cout ? p.Wp[k * p.Cout + n] : 0.f;
  }
  const bool fold = p.gn_scale != nullptr;   // GroupNorm + SiLU folded into the load (else: pre-applied)
  for (int i = threadIdx.x; i < p.C && fold; i += blockDim.x) {
    sc[i] = p.gn_scale[(long long)b * p.C + i];
    sh[i] = p.gn_shift[(long long)b * p.C + i];
  }
  __syncthreads();
  const int HW = p.H * p.W;
  const int Wp2 = (p.W + 1) >> 1;
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= p.H * Wp2) return;
  const int oh = pair / Wp2, ow = (pair - oh * Wp2) * 2;
  float acc[2][NACC];
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int n = 0; n < NACC; ++n) acc[q][n] = 0.f;
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(p.x) + (long long)b * HW * p.C;
  const int C8 = p.C / 8;
  for (int r = 0; r < 3; ++r) {
    const int ih = oh + r - 1;
    if (ih < 0 || ih >= p.H) continue;
    const uint4* rowp = reinterpret_cast<const uint4*>(x + (long long)ih * p.W * p.C);
    for (int c8 = 0; c8 < C8; ++c8) {
      float f[4][8];
#pragma unroll
      for (int col = 0; col < 4; ++col) {
        const int iw = ow - 1 + col;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (iw >= 0 && iw < p.W) q = __ldg(rowp + (long long)iw * C8 + c8);
        unpack8(q, f[col]);
        if (fold && iw >= 0 && iw < p.W) {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[col][e] = silu_f(fmaf(f[col][e], sc[c8 * 8 + e], sh[c8 * 8 + e]));
        }
      }
#pragma unroll
      for (int dw = 0; dw < 3; ++dw) {
        const float* wt = wsm + ((r * 3 + dw) * p.C + c8 * 8) * NACC;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
#pragma unroll
          for (int n4 = 0; n4 < NACC; n4 += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wt + e * NACC + n4);
            const float a0 = f[dw][e], a1 = f[dw + 1][e];
            acc[0][n4] = fmaf(a0, w4.x, acc[0][n4]);         acc[1][n4] = fmaf(a1, w4.x, acc[1][n4]);
            acc[0][n4 + 1] = fmaf(a0, w4.y, acc[0][n4 + 1]); acc[1][n4 + 1] = fmaf(a1, w4.y, acc[1][n4 + 1]);
            acc[0][n4 + 2] = fmaf(a0, w4.z, acc[0][n4 + 2]); acc[1][n4 + 2] = fmaf(a1, w4.z, acc[1][n4 + 2]);
            acc[0][n4 + 3] = fmaf(a0, w4.w, acc[0][n4 + 3]); acc[1][n4 + 3] = fmaf(a1, w4.w, acc[1][n4 + 3]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (ow + q >= p.W) continue;
#pragma unroll
    for (int n = 0; n < NACC; ++n)
      if (n < p.Cout)
        p.out_nchw[((long long)b * p.Cout + n) * HW + oh * p.W + ow + q] = acc[q][n] + p.bias[n];
  }
}
}  // namespace

int launch_conv_small_n(const ConvSmallNParams& p, cudaStream_t st) {
  EO_REQUIRE(p.dt == DT_BF16 && p.C % 8 == 0 && p.Cout <= 16, EO_ERR_ARG, "conv_small_n: shape");
  const int nacc = p.Cout <= 4 ? 4 : 16;
  size_t smem = (size_t)(9 * p.C * nacc + 2 * p.C) * sizeof(float);
  EO_REQUIRE(smem <= 200 * 1024, EO_ERR_ARG, "conv_small_n: weights do not fit shared memory");
  dim3 grid((unsigned)ceil_div(p.H * ((p.W + 1) / 2), 128), (unsigned)p.B);
  if (nacc == 4) {
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

// 16-byte vector of T as floats: 4 x fp32 or 8 x bf16
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float v[4]) {
    float4 q = *reinterpret_cast<const float4*>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float v[8]) {
    uint4 q = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[2 * e] = __uint_as_float(w[e] << 16);
      v[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
    }
  }
};

// grid (chunks, B); each CTA reduces `ppc` pixels of sample b over all channels of all
// sources into 32 (sum, sumsq) pairs, then adds them to sums[b] with 64 double atomics.
// A thread owns one 16-byte channel vector and walks the pixels with 4 loads in flight.
template <typename T, typename Acc>
__global__ void __launch_bounds__(256)
k_gn_stats(GnSrcs srcs, int HW, int ppc, int cpg, double* __restrict__ sums) {
  constexpr int N = Vec16<T>::N;
  __shared__ double sm[64];
  const int tid = threadIdx.x, b = blockIdx.y;
  const int p0 = blockIdx.x * ppc, p1 = min(HW, p0 + ppc);
  if (tid < 64) sm[tid] = 0.0;
  __syncthreads();
  int coff = 0;
  for (int s = 0; s < srcs.n; ++s) {
    const int C = srcs.s[s].C;
    const T* base = reinterpret_cast<const T*>(srcs.s[s].ptr) + (long long)b * HW * C;
    const int ncol = C / N;
    const int rows = blockDim.x / ncol;   // >= 1 (C <= 1024)
    const int col = tid % ncol, row = tid / ncol;
    if (row < rows) {
      Acc sv[N], qv[N];
#pragma unroll
      for (int j = 0; j < N; ++j) { sv[j] = 0; qv[j] = 0; }
      const T* colp = base + col * N;
      int p = p0 + row;
      for (; p + 3 * rows < p1; p += 4 * rows) {       // 4 loads in flight per thread
        float v[4][N];
#pragma unroll
        for (int u = 0; u < 4; ++u) Vec16<T>::load(colp + (long long)(p + u * rows) * C, v[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int j = 0; j < N; ++j) { sv[j] += (Acc)v[u][j]; qv[j] += (Acc)v[u][j] * (Acc)v[u][j]; }
      }
      for (; p < p1; p += rows) {
        float v[N];
        Vec16<T>::load(colp + (long long)p * C, v);
#pragma unroll
        for (int j = 0; j < N; ++j) { sv[j] += (Acc)v[j]; qv[j] += (Acc)v[j] * (Acc)v[j]; }
      }
      // channels -> groups: merge runs of channels that fall into the same group
      const int c0 = coff + col * N;
      double rs = 0.0, rq = 0.0;
      int g = c0 / cpg;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const int gj = (c0 + j) / cpg;
        if (gj != g) {
          atomicAdd(&sm[2 * g], rs); atomicAdd(&sm[2 * g + 1], rq);
          rs = 0.0; rq = 0.0; g = gj;
        }
        rs += (double)sv[j]; rq += (double)qv[j];
      }
      atomicAdd(&sm[2 * g], rs); atomicAdd(&sm[2 * g + 1], rq);
    }
    coff += C;
  }
  __syncthreads();
  if (tid < 64) atomicAdd(&sums[(long long)b * 64 + tid], sm[tid]);
}

// Per-channel (sum, sum of squares) of one NHWC bf16 tensor -> stats[b][c][2] (fp32 atomics), for
// tensors that are not written by the tensor-core conv (whose epilogue emits the same sums).
__global__ void __launch_bounds__(256)
k_chan_stats(const __nv_bfloat16* __restrict__ x, int C, int HW, int ppc, double* __restrict__ stats) {
  extern __shared__ double sm[];    // [C][2]; double so that the atomic order does not show in the result
  const int tid = threadIdx.x, b = blockIdx.y;
  const int p0 = blockIdx.x * ppc, p1 = min(HW, p0 + ppc);
  for (int i = tid; i < 2 * C; i += blockDim.x) sm[i] = 0.0;
  __syncthreads();
  const __nv_bfloat16* base = x + (long long)b * HW * C;
  const int ncol = C / 8;
  const int rows = blockDim.x / ncol;
  const int col = tid % ncol, row = tid / ncol;
  if (row < rows) {
    float sv[8], qv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sv[j] = 0.f; qv[j] = 0.f; }
    for (int p = p0 + row; p < p1; p += rows) {
      float v[8];
      Vec16<__nv_bfloat16>::load(base + (long long)p * C + col * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { sv[j] += v[j]; qv[j] = fmaf(v[j], v[j], qv[j]); }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sm[(col * 8 + j) * 2], (double)sv[j]);
      atomicAdd(&sm[(col * 8 + j) * 2 + 1], (double)qv[j]);
    }
  }
  __syncthreads();
  for (int i = tid; i < 2 * C; i += blockDim.x) atomicAdd(stats + (long long)b * C * 2 + i, sm[i]);
}

// GroupNorm scale/shift from per-channel sums of one or two concatenated tensors.  One thread per channel, a block
// holds gpb whole groups of cpg channels: every thread fetches its own (sum, sum of squares) -- ONE load per thread,
// all in flight together; the kernel used to walk the cpg channels of its group with dependent loads, 7 us per launch
// 56 times per forward -- parks it in shared memory, and adds up its group in channel order (the order of the
// sequential sum it replaces: bit-identical rows).
__global__ void __launch_bounds__(256)
k_gn_finalize_ch(const double* __restrict__ sa, int Ca, const double* __restrict__ sb, int Cb,
                 const float* __restrict__ gamma, const float* __restrict__ beta, int B,
                 int HW, int cpg, int gpb, float* __restrict__ scale, float* __restrict__ shift) {
  __shared__ double2 sm[256];
  pdl_wait(); pdl_trigger();
  const int C = Ca + Cb;
  const int tid = threadIdx.x;
  const int gl = tid / cpg, k = tid - gl * cpg;          // group inside the block, channel inside the group
  const long long gi = (long long)blockIdx.x * gpb + gl;   // (image, group) pair
  const bool on = gl < gpb && gi < (long long)B * 32;
  const int b = (int)(gi >> 5), c = (int)(gi & 31) * cpg + k;
  if (on) {
    const double* src = c < Ca ? sa + ((long long)b * Ca + c) * 2 : sb + ((long long)b * Cb + (c - Ca)) * 2;
    sm[tid] = *reinterpret_cast<const double2*>(src);
  }
  __syncthreads();
  if (!on) return;
  double s = 0.0, q = 0.0;
  const double2* grp = sm + gl * cpg;
  for (int j = 0; j < cpg; ++j) { s += grp[j].x; q += grp[j].y; }
  const double n = (double)cpg * (double)HW;
  const double mean = s / n;
  double var = q / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const double rstd = 1.0 / sqrt(var + 1e-5);
  const long long i = (long long)b * C + c;
  scale[i] = (float)(rstd * (double)gamma[c]);
  shift[i] = (float)((double)beta[c] - mean * rstd * (double)gamma[c]);
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
  pdl_wait(); pdl_trigger();
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
      auto xform = [&](uint4 q) -> uint4 {
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f = __bfloat1622float2(h2[e]);
          f.x = fmaf(f.x, sc[2 * e], sh[2 * e]);
          f.y = fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1]);
          if (silu) { f.x = silu_f(f.x); f.y = silu_f(f.y); }
          h2[e] = __floats2bfloat162_rn(f.x, f.y);
        }
        return q;
      };
      const __nv_bfloat16* src = base + col * 8;
      __nv_bfloat16* out = dst + (long long)b * HW * Ctot + coff + col * 8;
      int p = p0 + row;
      for (; p + 3 * rows < p1; p += 4 * rows) {       // 4 loads in flight per thread
        uint4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) q[u] = *reinterpret_cast<const uint4*>(src + (long long)(p + u * rows) * C);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          *reinterpret_cast<uint4*>(out + (long long)(p + u * rows) * Ctot) = xform(q[u]);
      }
      for (; p < p1; p += rows)
        *reinterpret_cast<uint4*>(out + (long long)p * Ctot) =
            xform(*reinterpret_cast<const uint4*>(src + (long long)p * C));
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
    const int vec = dt == DT_F32 ? 4 : 8;
    EO_REQUIRE(src[i].C % vec == 0 && src[i].C <= 256 * vec && src[i].C > 0, EO_ERR_ARG,
               "gn_stats: channels must be a multiple of %d and <= %d (got %d)", vec, 256 * vec, src[i].C);
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

int launch_chan_stats(const void* x_bf16, int B, int HW, int C, double* stats, cudaStream_t st) {
  EO_REQUIRE(C % 8 == 0 && C <= 2048, EO_ERR_ARG, "chan_stats: channels");
  int ppc = pick_ppc(B, HW);
  dim3 grid((unsigned)ceil_div(HW, ppc), (unsigned)B);
  k_chan_stats<<<grid, 256, 2 * C * sizeof(double), st>>>(reinterpret_cast<const __nv_bfloat16*>(x_bf16), C, HW, ppc, stats);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

int launch_gn_finalize_ch(const double* sa, int Ca, const double* sb, int Cb, const float* gamma, const float* beta,
                          int B, int HW, float* scale, float* shift, cudaStream_t st) {
  const int C = Ca + Cb;
  EO_REQUIRE(C % 32 == 0 && C / 32 <= 256, EO_ERR_ARG, "gn_finalize_ch: %d channels (a multiple of 32, at most 8192)", C);
  const int cpg = C / 32, gpb = 256 / cpg;      // whole groups per 256-thread block
  EO_CHECK_CUDA(launch_chain(k_gn_finalize_ch, dim3((unsigned)ceil_div((long long)B * 32, gpb)), dim3(256), 0, st, sa, Ca,
                             sb, Cb, gamma, beta, B, HW, cpg, gpb, scale, shift));
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
  EO_CHECK_CUDA(launch_chain(k_gn_apply, grid, dim3(256), 0, st, s, HW, ppc, Ctot, scale, shift, silu,
                             reinterpret_cast<__nv_bfloat16*>(dst)));
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

namespace {
// out[b, :] = table[t[b], :] for 0 <= t[b] < n, NaN otherwise (a timestep outside the precomputed table must not
// pass silently); rows of `ld` floats, ld % 4 == 0
__global__ void __launch_bounds__(256)
k_gather_rows(const float4* __restrict__ table, int n, int ld4, const long long* __restrict__ t, float4* __restrict__ out) {
  pdl_wait(); pdl_trigger();
  const int b = blockIdx.y;
  const long long tv = t[b];
  const bool ok = tv >= 0 && tv < n;
  const float qn = __int_as_float(0x7fc00000);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ld4; i += gridDim.x * blockDim.x)
    out[(long long)b * ld4 + i] = ok ? __ldg(table + tv * ld4 + i) : make_float4(qn, qn, qn, qn);
}
__global__ void k_iota64(long long* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}
}  // namespace

int launch_gather_rows(const float* table, int n, int ld, const int64_t* t, int B, float* out, cudaStream_t st) {
  EO_REQUIRE(ld % 4 == 0, EO_ERR_ARG, "gather_rows: row length %d is not a multiple of 4", ld);
  dim3 grid((unsigned)std::min<long long>(ceil_div(ld / 4, 256), 8), (unsigned)B);
  EO_CHECK_CUDA(launch_chain(k_gather_rows, grid, dim3(256), 0, st, reinterpret_cast<const float4*>(table), n, ld / 4,
                             reinterpret_cast<const long long*>(t), reinterpret_cast<float4*>(out)));
  return EO_OK;
}

int launch_iota64(int64_t* out, int n, cudaStream_t st) {
  k_iota64<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(reinterpret_cast<long long*>(out), n);
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

// ---------------------------------------------------------------------------------------
// Wide heads (head dimension > 64): what the reference's own scripts build -- num_heads = 1, so the
// middle-block attention has one head of 512 (train.py:50) or 1024 (inference.py:59) channels.
// One WARP per query row: lane l owns channels l, l + 32, ... of q and of the output accumulator;
// a block of 8 warps shares K / V tiles of KT keys staged in shared memory as fp32.  Per tile: KT dot
// products (a 5-step warp reduction each; lane j keeps the logit of key j), one online-softmax update,
// KT rank-1 updates of o.  Same arithmetic order for fp32 and bf16 inputs (fp32 throughout).
// ---------------------------------------------------------------------------------------
constexpr int AW_WARPS = 8;

template <typename TIO, int NC>
__global__ void __launch_bounds__(AW_WARPS * 32)
k_attention_wide(const TIO* __restrict__ qkv, TIO* __restrict__ out, int T, int heads, int ch, int ld,
                 int head_stride, int part_stride, float scale, int KT) {
  extern __shared__ float aw_smem[];
  float* Ks = aw_smem;                       // [KT][ch]
  float* Vs = aw_smem + (size_t)KT * ch;     // [KT][ch]
  const int b = blockIdx.z, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qi = blockIdx.x * AW_WARPS + warp;
  const TIO* base = qkv + (long long)b * T * ld + (long long)h * head_stride;
  float q[NC], o[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int d = lane + 32 * i;
    q[i] = (qi < T && d < ch) ? to_float(base[(long long)qi * ld + d]) * scale : 0.f;   // q * scale (:475)
    o[i] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < T; k0 += KT) {
    const int nk = min(KT, T - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < nk * ch; i += AW_WARPS * 32) {
      const int kk = i / ch, d = i - kk * ch;
      const TIO* row = base + (long long)(k0 + kk) * ld;
      Ks[i] = to_float(row[part_stride + d]) * scale;                                    // k * scale
      Vs[i] = to_float(row[2 * part_stride + d]);
    }
    __syncthreads();
    float sj = -INFINITY;                    // lane j: logit of key k0 + j
    for (int j = 0; j < nk; ++j) {
      const float* kr = Ks + (size_t)j * ch + lane;
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < NC; ++i)
        if (lane + 32 * i < ch) a = fmaf(q[i], kr[32 * i], a);
      a = warp_sum(a);
      if (lane == j) sj = a;
    }
    const float mn = fmaxf(m, warp_max(sj));
    const float alpha = expf(m - mn);        // first tile: exp(-inf) = 0
    const float pj = lane < nk ? expf(sj - mn) : 0.f;
    l = l * alpha + warp_sum(pj);
    m = mn;
#pragma unroll
    for (int i = 0; i < NC; ++i) o[i] *= alpha;
    for (int j = 0; j < nk; ++j) {
      const float p = __shfl_sync(0xffffffffu, pj, j);
      const float* vr = Vs + (size_t)j * ch + lane;
#pragma unroll
      for (int i = 0; i < NC; ++i)
        if (lane + 32 * i < ch) o[i] = fmaf(p, vr[32 * i], o[i]);
    }
  }
  if (qi < T) {
    const float inv = 1.0f / l;
    TIO* op = out + ((long long)b * T + qi) * ((long long)heads * ch) + (long long)h * ch;
#pragma unroll
    for (int i = 0; i < NC; ++i)
      if (lane + 32 * i < ch) op[lane + 32 * i] = from_float<TIO>(o[i] * inv);
  }
}

template <typename TIO>
int launch_attention_wide_t(const TIO* qkv, TIO* out, int B, int T, int heads, int ch, int ld, int head_stride,
                            int part_stride, cudaStream_t st) {
  EO_REQUIRE(ch <= 1024, EO_ERR_ARG, "attention: head dimension %d > 1024 is not supported", ch);
  const float scale = 1.0f / sqrtf(sqrtf((float)ch));
  int KT = 8192 / ch;                          // K and V tiles of <= 32 KB each
  KT = KT > 32 ? 32 : KT;
  const size_t smem = (size_t)2 * KT * ch * sizeof(float);
  dim3 grid((unsigned)ceil_div(T, AW_WARPS), (unsigned)heads, (unsigned)B);
#define EO_AW(NCV)                                                                                                  \
  do {                                                                                                              \
    static bool attr_set = false;       /* once per instantiation, outside any stream capture that follows */      \
    if (!attr_set) {                                                                                                \
      EO_CHECK_CUDA(cudaFuncSetAttribute(k_attention_wide<TIO, NCV>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                         64 * 1024));                                                               \
      attr_set = true;                                                                                              \
    }                                                                                                               \
    k_attention_wide<TIO, NCV><<<grid, AW_WARPS * 32, smem, st>>>(qkv, out, T, heads, ch, ld, head_stride,         \
                                                                   part_stride, scale, KT);                         \
  } while (0)
  if (ch <= 128) EO_AW(4); else if (ch <= 256) EO_AW(8); else if (ch <= 512) EO_AW(16); else EO_AW(32);
#undef EO_AW
  EO_CHECK_LAUNCH();
  return EO_OK;
}
}  // namespace

int launch_attention_wide(const void* qkv, void* out, int dt, int B, int T, int heads, int ch, int ld, int head_stride,
                          int part_stride, cudaStream_t st) {
  if (dt == DT_F32)
    return launch_attention_wide_t(reinterpret_cast<const float*>(qkv), reinterpret_cast<float*>(out), B, T, heads, ch, ld,
                                   head_stride, part_stride, st);
  return launch_attention_wide_t(reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), B, T,
                                 heads, ch, ld, head_stride, part_stride, st);
}

int launch_attention_simt(const float* qkv, float* out, int B, int T, int heads, int ch, int ld,
                          int head_stride, int part_stride, cudaStream_t st) {
  if (ch > AT_D) return launch_attention_wide(qkv, out, DT_F32, B, T, heads, ch, ld, head_stride, part_stride, st);
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


int launch_nhwc_to_nchw_f32(const void* src, int dt, float* dst, int B, int HW, int C,
                            cudaStream_t st) {
  long long total = (long long)B * HW * C;
  k_nhwc_to_nchw<<<ew_grid(total), 256, 0, st>>>(src, dt, dst, HW, C, total);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

// =======================================================================================
// ResBlock variants of the reference's UNet / UNetBig / UNetSmall factories (unet_openai.py:783-922):
// up/down-sampling blocks (:321-328, :366-371) and scale-shift (FiLM) conditioning (:377-381)
// =======================================================================================
namespace {
template <typename T> struct VecIO;
template <> struct VecIO<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float v[4]) {
    const float4 q = *reinterpret_cast<const float4*>(p); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
  static __device__ __forceinline__ void store(float* p, const float v[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  static __device__ __forceinline__ float act(float x) { return silu_acc(x); }
};
template <> struct VecIO<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float v[8]) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) { v[2 * e] = __uint_as_float(w[e] << 16); v[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u); }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float v[8]) {
    uint4 q;
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
    for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4*>(p) = q;
  }
  static __device__ __forceinline__ float act(float x) { return silu_f(x); }
};

// dst[b, ho, wo, coff + c] = resample(act(src[b, hi, wi, c] * scale[b, gc + c] + shift[b, gc + c]))
//   mode 0: identity grid; 1: nearest x2 (F.interpolate, Upsample without conv); 2: 2x2 average (AvgPool2d,
//   Downsample without conv), taken AFTER the affine + activation like the reference (h_upd follows SiLU)
template <typename T>
__global__ void __launch_bounds__(256)
k_resample(const T* __restrict__ src, int Cs, T* __restrict__ dst, int Cd, int coff, int Hi, int Wi, int Ho, int Wo, int mode,
           const float* __restrict__ scale, const float* __restrict__ shift, int gld, int gc, int silu, long long total) {
  constexpr int N = VecIO<T>::N;
  const int cv = Cs / N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * N;
    long long r = i / cv;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const long long b = r / Ho;
    float sc[N], sh[N];
    if (scale) {
#pragma unroll
      for (int j = 0; j < N; ++j) { sc[j] = scale[b * gld + gc + c + j]; sh[j] = shift[b * gld + gc + c + j]; }
    }
    auto fetch = [&](int hi, int wi, float v[N]) {
      VecIO<T>::load(src + ((b * Hi + hi) * Wi + wi) * Cs + c, v);
      if (scale) {
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
      }
      if (silu) {
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = VecIO<T>::act(v[j]);
      }
    };
    float o[N];
    if (mode == 2) {
      float v0[N], v1[N], v2[N], v3[N];
      fetch(2 * ho, 2 * wo, v0); fetch(2 * ho, 2 * wo + 1, v1); fetch(2 * ho + 1, 2 * wo, v2); fetch(2 * ho + 1, 2 * wo + 1, v3);
#pragma unroll
      for (int j = 0; j < N; ++j) o[j] = ((v0[j] + v1[j]) + (v2[j] + v3[j])) * 0.25f;
    } else {
      fetch(mode == 1 ? ho >> 1 : ho, mode == 1 ? wo >> 1 : wo, o);
    }
    VecIO<T>::store(dst + ((b * Ho + ho) * Wo + wo) * Cd + coff + c, o);
  }
}

// GroupNorm folded to x * scale + shift, then FiLM: (x * scale + shift) * (1 + s) + t with (s | t) = the block's
// row of the per-step embedding table -> scale *= (1 + s); shift = shift * (1 + s) + t        (unet_openai.py:378-380)
__global__ void k_gn_modulate(float* __restrict__ scale, float* __restrict__ shift, const float* __restrict__ tb, int ld, int off,
                              int B, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C;
  const float s1 = 1.0f + tb[(long long)b * ld + off + c], t = tb[(long long)b * ld + off + C + c];
  const float sc = scale[i], sh = shift[i];
  scale[i] = sc * s1;
  shift[i] = fmaf(sh, s1, t);
}
}  // namespace

int launch_resample(const void* src, int Cs, void* dst, int Cd, int coff, int dt, int B, int Hi, int Wi, int mode,
                    const float* scale, const float* shift, int gld, int gc, int silu, cudaStream_t st) {
  const int N = dt == DT_F32 ? 4 : 8;
  EO_REQUIRE(Cs % N == 0 && Cd % N == 0 && coff % N == 0, EO_ERR_ARG, "resample: channel counts must be multiples of %d", N);
  EO_REQUIRE(mode != 2 || (Hi % 2 == 0 && Wi % 2 == 0), EO_ERR_ARG, "resample: odd feature map %dx%d", Hi, Wi);
  const int Ho = mode == 1 ? 2 * Hi : mode == 2 ? Hi / 2 : Hi, Wo = mode == 1 ? 2 * Wi : mode == 2 ? Wi / 2 : Wi;
  const long long total = (long long)B * Ho * Wo * (Cs / N);
  if (dt == DT_F32)
    k_resample<float><<<ew_grid(total), 256, 0, st>>>(reinterpret_cast<const float*>(src), Cs, reinterpret_cast<float*>(dst), Cd, coff,
                                                       Hi, Wi, Ho, Wo, mode, scale, shift, gld, gc, silu, total);
  else
    k_resample<__nv_bfloat16><<<ew_grid(total), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(src), Cs,
                                                               reinterpret_cast<__nv_bfloat16*>(dst), Cd, coff, Hi, Wi, Ho, Wo, mode,
                                                               scale, shift, gld, gc, silu, total);
  EO_CHECK_LAUNCH();
  return EO_OK;
}
int launch_gn_modulate(float* scale, float* shift, const float* tb, int ld, int off, int B, int C, cudaStream_t st) {
  k_gn_modulate<<<ceil_div((long long)B * C, 256), 256, 0, st>>>(scale, shift, tb, ld, off, B, C);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

// =======================================================================================
// tensor-core stem and head (bf16 mode): layout passes either side of k_conv_tc
// =======================================================================================
namespace {
// im2col of the few-channel network input for the stem conv (unet_openai.py:608, :754-756): one 64-channel
// bf16 NHWC pixel per output pixel, channels [0, 9C) = the 3x3 window of cat(x, cond) ordered (tap, channel)
// rounded to bf16, channels [9C, 18C) = the rounding residuals (x - bf16(x), exact to 2^-17), rest zero.  The
// stem is then a 64-deep 1x1 convolution on the tensor cores with the weights repeated for both halves.
template <int C>
__global__ void __launch_bounds__(256)
k_stem_im2col(const float* __restrict__ x, int Cx, const float* __restrict__ cond, int Cc, __nv_bfloat16* __restrict__ dst,
              int H, int W, long long npix) {
  // one thread per pixel: 9 C coalesced loads (neighbouring threads read neighbouring pixels of an NCHW row),
  // 64 bf16 built in registers; the 128-byte pixel rows go through shared memory (16-byte chunk j of pixel p at
  // position j ^ (p & 7)) so that a warp's store instruction covers 512 contiguous bytes (four whole pixels) instead
  // of 32 half-filled sectors 128 bytes apart
  __shared__ uint4 stage[256 * 8];
  const long long HW = (long long)H * W;
  const int tid = threadIdx.x;
  for (long long pix0 = blockIdx.x * (long long)blockDim.x; pix0 < npix; pix0 += (long long)gridDim.x * blockDim.x) {
    const long long pix = pix0 + tid < npix ? pix0 + tid : npix - 1;      // the tail recomputes the last pixel, stores are guarded
    const int b = (int)(pix / HW);
    const int r = (int)(pix - (long long)b * HW);
    const int oh = r / W, ow = r - oh * W;
    float v[9 * C];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int ih = oh + tap / 3 - 1, iw = ow + tap % 3 - 1;
      const bool in = ih >= 0 && ih < H && iw >= 0 && iw < W;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float* src = c < Cx ? x + ((long long)b * Cx + c) * HW : cond + ((long long)b * Cc + (c - Cx)) * HW;
        v[tap * C + c] = in ? __ldg(src + (long long)ih * W + iw) : 0.f;
      }
    }
    __align__(16) __nv_bfloat16 e[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) {
      float f = 0.f;
      if (k < 9 * C) f = v[k];
      else if (k < 18 * C) f = v[k - 9 * C] - __bfloat162float(__float2bfloat16_rn(v[k - 9 * C]));
      e[k] = __float2bfloat16_rn(f);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) stage[tid * 8 + (j ^ (tid & 7))] = reinterpret_cast<const uint4*>(e)[j];
    __syncthreads();
    uint4* o = reinterpret_cast<uint4*>(dst + pix0 * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int p = i * 32 + (tid >> 3), j = tid & 7;       // pixel inside the block, chunk
      if (pix0 + p < npix) o[p * 8 + j] = stage[p * 8 + (j ^ (p & 7))];
    }
    __syncthreads();
  }
}
// 4..32 input channels: cat(x, cond) (NCHW fp32) -> one 64-channel bf16 NHWC pixel, channels [0, C) the values
// rounded to bf16, [C, 2C) their rounding residuals, rest zero; the stem is then an ordinary 3x3 tensor-core conv
__global__ void __launch_bounds__(256)
k_stem_nhwc(const float* __restrict__ x, int Cx, const float* __restrict__ cond, int Cc, __nv_bfloat16* __restrict__ dst,
            long long HW, long long npix) {
  const int C = Cx + Cc;
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < npix; pix += (long long)gridDim.x * blockDim.x) {
    const long long b = pix / HW, r = pix - b * HW;
    __align__(16) __nv_bfloat16 e[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) e[k] = __float2bfloat16_rn(0.f);
    for (int c = 0; c < C; ++c) {
      const float v = c < Cx ? __ldg(x + (b * Cx + c) * HW + r) : __ldg(cond + (b * Cc + (c - Cx)) * HW + r);
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      e[c] = hi;
      e[C + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
    uint4* o = reinterpret_cast<uint4*>(dst + pix * 64);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = reinterpret_cast<const uint4*>(e)[j];
  }
}
// w [Cout][C][3][3] -> w2 [Cout][64][3][3] matching k_stem_nhwc's channel order
__global__ void k_stem_weight3(const float* __restrict__ w, int Cout, int C, float* __restrict__ w2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * 64 * 9) return;
  const int tap = i % 9, c = (i / 9) & 63, n = i / (9 * 64);
  const int cs = c < C ? c : c - C;
  w2[i] = c < 2 * C ? w[((long long)n * C + cs) * 9 + tap] : 0.f;
}
// w [Cout][C][3][3] -> w2 [Cout][64] matching k_stem_im2col's channel order
__global__ void k_stem_weight(const float* __restrict__ w, int Cout, int C, float* __restrict__ w2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * 64) return;
  const int n = i >> 6;
  int k = i & 63;
  if (k >= 9 * C) k -= 9 * C;
  float v = 0.f;
  if (k < 9 * C) { const int tap = k / C, c = k - tap * C; v = w[((long long)n * C + c) * 9 + tap]; }
  w2[i] = v;
}
}  // namespace

int launch_stem_im2col(const float* x, int Cx, const float* cond, int Cc, void* dst, int B, int H, int W, cudaStream_t st) {
  EO_REQUIRE(18 * (Cx + Cc) <= 64, EO_ERR_ARG, "stem_im2col: %d input channels do not fit one 64-deep K block", Cx + Cc);
  const long long npix = (long long)B * H * W;
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst);
  switch (Cx + Cc) {
    case 1: k_stem_im2col<1><<<ew_grid(npix), 256, 0, st>>>(x, Cx, cond, Cc, d, H, W, npix); break;
    case 2: k_stem_im2col<2><<<ew_grid(npix), 256, 0, st>>>(x, Cx, cond, Cc, d, H, W, npix); break;
    default: k_stem_im2col<3><<<ew_grid(npix), 256, 0, st>>>(x, Cx, cond, Cc, d, H, W, npix); break;
  }
  EO_CHECK_LAUNCH();
  return EO_OK;
}
int launch_stem_nhwc(const float* x, int Cx, const float* cond, int Cc, void* dst, int B, int H, int W, cudaStream_t st) {
  EO_REQUIRE(2 * (Cx + Cc) <= 64, EO_ERR_ARG, "stem_nhwc: %d input channels do not fit one 64-channel block", Cx + Cc);
  const long long npix = (long long)B * H * W;
  k_stem_nhwc<<<ew_grid(npix), 256, 0, st>>>(x, Cx, cond, Cc, reinterpret_cast<__nv_bfloat16*>(dst), (long long)H * W, npix);
  EO_CHECK_LAUNCH();
  return EO_OK;
}
int launch_stem_weight3(const float* w, int Cout, int C, float* w2, cudaStream_t st) {
  k_stem_weight3<<<ceil_div(Cout * 64 * 9, 256), 256, 0, st>>>(w, Cout, C, w2);
  EO_CHECK_LAUNCH();
  return EO_OK;
}
int launch_stem_weight(const float* w, int Cout, int C, float* w2, cudaStream_t st) {
  k_stem_weight<<<ceil_div(Cout * 64, 256), 256, 0, st>>>(w, Cout, C, w2);
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
                                   const int* __restrict__ row_map, const float* __restrict__ row_scale) {
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
    if (row_scale) v *= row_scale[n];
    long long o = (long long)n * stride_n + (long long)(k_off + tap * C + c) * stride_k;
    if (dst_dt == DT_F32) reinterpret_cast<float*>(dst)[o] = v;
    else reinterpret_cast<__nv_bfloat16*>(dst)[o] = __float2bfloat16_rn(v);
  }
}
__global__ void k_pack_bias(const float* a, const float* b, float* dst, int Nout,
                            const int* row_map, const float* row_scale) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Nout) return;
  int row = row_map ? row_map[n] : n;
  float v = 0.f;
  if (row >= 0) { if (a) v += a[row]; if (b) v += b[row]; }
  if (row_scale) v *= row_scale[n];
  dst[n] = v;
}
// x[r][col0 + c] = bf16(x[r][col0 + c] * scale), c < ncols (the attention self-test's stand-in for a scaled q projection)
__global__ void k_scale_cols_bf16(__nv_bfloat16* x, long long rows, int ld, int col0, int ncols, float scale) {
  const long long total = rows * ncols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    __nv_bfloat16* p = x + (i / ncols) * ld + col0 + (int)(i % ncols);
    *p = __float2bfloat16_rn(__bfloat162float(*p) * scale);
  }
}
}  // namespace

int launch_pack_conv_weight(const float* w, int Cin_total, int ksize, int cin_off, int C,
                            void* dst, int dst_dt, long long stride_n, long long stride_k,
                            int k_off, int Nout, const int* row_map, cudaStream_t st, const float* row_scale) {
  long long total = (long long)Nout * ksize * ksize * C;
  k_pack_conv_weight<<<ew_grid(total), 256, 0, st>>>(w, Cin_total, ksize, cin_off, C, dst, dst_dt,
                                                     stride_n, stride_k, k_off, Nout, row_map, row_scale);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

namespace {
// Nearest x2 upsampling followed by a 3x3 convolution (Upsample.forward, unet_openai.py:229-242) equals,
// for output pixel (2h+a, 2w+b), a 2x2 convolution over the LOW-resolution input with taps at rows
// {a-1, a} and columns {b-1, b}: the 3x3 taps that read the same source pixel are summed.
//   a = 0: row -1 <- kh 0, row 0 <- kh 1+2;   a = 1: row 0 <- kh 0+1, row +1 <- kh 2   (same for columns)
__global__ void k_fold_upsample_weight(const float* __restrict__ w, long long n_pairs, int a, int b,
                                       float* __restrict__ wf) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_pairs;
       i += (long long)gridDim.x * blockDim.x) {
    const float* s = w + i * 9;
    float rows[2][3];
    for (int kw = 0; kw < 3; ++kw) {
      rows[0][kw] = a == 0 ? s[kw] : s[kw] + s[3 + kw];
      rows[1][kw] = a == 0 ? s[3 + kw] + s[6 + kw] : s[6 + kw];
    }
    for (int r = 0; r < 2; ++r) {
      wf[i * 4 + r * 2 + 0] = b == 0 ? rows[r][0] : rows[r][0] + rows[r][1];
      wf[i * 4 + r * 2 + 1] = b == 0 ? rows[r][1] + rows[r][2] : rows[r][2];
    }
  }
}
}  // namespace

int launch_fold_upsample_weight(const float* w, int Cout, int Cin, int a, int b, float* wf, cudaStream_t st) {
  const long long n = (long long)Cout * Cin;
  k_fold_upsample_weight<<<ew_grid(n), 256, 0, st>>>(w, n, a, b, wf);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

int launch_pack_bias(const float* a, const float* b, float* dst, int Nout, const int* row_map,
                     cudaStream_t st, const float* row_scale) {
  k_pack_bias<<<(unsigned)ceil_div(Nout, 128), 128, 0, st>>>(a, b, dst, Nout, row_map, row_scale);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

int launch_scale_cols_bf16(void* x, long long rows, int ld, int col0, int ncols, float scale, cudaStream_t st) {
  k_scale_cols_bf16<<<ew_grid(rows * ncols), 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(x), rows, ld, col0, ncols, scale);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

}  // namespace eo
