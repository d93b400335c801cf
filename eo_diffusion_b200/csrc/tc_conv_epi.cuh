// Epilogue shared by the tcgen05 convolution kernels (tc_conv.cu, tc_conv3.cu):
// TMEM accumulator -> registers -> (+bias, +per-sample timestep-embedding row, +residual) -> NHWC
// store, plus the per-channel GroupNorm partial sums of the tensor being written.
// Reference: the adds are `h + emb_out` (backbones/unet_openai.py:382), `skip_connection(x) + h`
// (:385), `x + h` (:433); the statistics feed nn.GroupNorm (:11-13) of the consumer.
#pragma once
#include "tc_common.cuh"

namespace eo { namespace tc {

struct Epi {
  const float* bias;        // [Cout] or null
  const float* bias_nc;     // [B, ld_bias_nc] or null
  int ld_bias_nc;
  const void* residual;     // NHWC [B,H,W,Cout], fp32 if res_f32 else bf16, or null
  void* out;                // NHWC [B,H,W,Cout], fp32 if out_f32 else bf16
  double* stats;            // [B, Cout, 2] (sum, sum of squares) accumulated with atomics, or null
  int Cout, res_f32, out_f32;
  long long* trace;         // development aid (eo_debug_conv_trace): [n][8] per-CTA phase stamps, or null
  int trace_n;
};

// phase stamp for the per-CTA timeline: slot 0 = globaltimer at entry, 1..6 = clock64 at a phase, 7 = smid
__device__ __forceinline__ void trace_stamp(const Epi& ep, int slot) {
  if (ep.trace && (int)blockIdx.x < ep.trace_n && blockIdx.y == 0) {
    long long v;
    if (slot == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); v = (long long)t; }
    else if (slot == 7) { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); v = s; }
    else v = clock64();
    ep.trace[(long long)blockIdx.x * 8 + slot] = v;
  }
}

// column sums over the 32 lanes of a warp: on return lane j holds sum_over_lanes(f[j]).
// Recursive halving: 16 + 8 + 4 + 2 + 1 = 31 shuffles instead of 32 x 5.
__device__ __forceinline__ float warp_column_sums(float (&f)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float keep = upper ? f[i + off] : f[i];
      const float send = upper ? f[i] : f[i + off];
      f[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return f[0];
}

// Run by the four epilogue warps of a CTA after the accumulator is complete.
//   q            TMEM lane quadrant of this warp (warp_id % 4); this thread owns accumulator row q*32+lane
//   valid        the row's image index is inside the batch
//   pix          linear NHWC pixel index of the row
//   single_image all 128 rows of the tile lie in image n_img: the four warps combine their partial
//                sums in `sstat` ([4][BN][2] floats of shared memory) before the atomics
//   et           index of this thread among the 128 epilogue threads
template <int BN>
__device__ __forceinline__ void conv_epilogue(uint32_t tmem_base, int q, int lane, int et, int nbase, bool valid,
                                              int n_img, long long pix, const Epi& ep, bool single_image,
                                              float* sstat) {
  const float* bnc = (ep.bias_nc && valid) ? ep.bias_nc + (long long)n_img * ep.ld_bias_nc : nullptr;
  const int Cout = ep.Cout;
  // statistics need a warp's 32 rows inside one image (the host guarantees it when stats != null)
  const bool do_stats = ep.stats != nullptr;
  const bool warp_valid = __shfl_sync(0xffffffffu, valid ? 1 : 0, 0) != 0;
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
    const int n = nbase + c0;
    if (n >= Cout) continue;             // warp-uniform
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    if (valid) {
      if (ep.bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n + j));
          f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
        }
      }
      if (bnc) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 b4 = __ldg(reinterpret_cast<const float4*>(bnc + n + j));
          f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
        }
      }
      const long long o = pix * Cout + n;
      if (ep.residual) {
        if (ep.res_f32) {
          const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ep.residual) + o);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 r = __ldg(rp + j);
            f[4 * j] += r.x; f[4 * j + 1] += r.y; f[4 * j + 2] += r.z; f[4 * j + 3] += r.w;
          }
        } else {
          const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(ep.residual) + o);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 r = __ldg(rp + j);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float2 t = __bfloat1622float2(h2[e]);
              f[j * 8 + e * 2] += t.x; f[j * 8 + e * 2 + 1] += t.y;
            }
          }
        }
      }
      if (ep.out_f32) {
        float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + o);
#pragma unroll
        for (int j = 0; j < 8; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
      } else {
        uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 w;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            h2[e] = __floats2bfloat162_rn(f[j * 8 + e * 2], f[j * 8 + e * 2 + 1]);
          op[j] = w;
        }
      }
    }
    if (do_stats && warp_valid) {
      // per-channel sum and sum of squares of this warp's 32 pixel rows (GroupNorm statistics of
      // the tensor being written; reference nn.GroupNorm reduces them per group later)
      float sq[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) sq[j] = f[j] * f[j];
      const float cs = warp_column_sums(f, lane);
      const float cq = warp_column_sums(sq, lane);
      // Deterministic: every partial sum is formed in a fixed order in fp32; only the final
      // accumulation across tiles is atomic, and that one is in double (order effects ~1e-16).
      if (single_image) {               // the 4 warps meet in smem first
        sstat[(q * BN + c0 + lane) * 2] = cs;
        sstat[(q * BN + c0 + lane) * 2 + 1] = cq;
      } else {
        double* dst = ep.stats + ((long long)n_img * Cout + n + lane) * 2;
        atomicAdd(dst, (double)cs);
        atomicAdd(dst + 1, (double)cq);
      }
    }
  }
  if (do_stats && single_image) {
    asm volatile("bar.sync 1, 128;" ::: "memory");      // the 4 epilogue warps
    const bool tile_valid = __shfl_sync(0xffffffffu, valid ? 1 : 0, 0) != 0;   // same for all rows of the tile
    if (tile_valid) {
      for (int i = et; i < BN * 2; i += 128) {
        const int c = nbase + (i >> 1);
        if (c < Cout) {
          const float tsum = (sstat[i] + sstat[BN * 2 + i]) + (sstat[2 * BN * 2 + i] + sstat[3 * BN * 2 + i]);
          atomicAdd(ep.stats + ((long long)n_img * Cout + c) * 2 + (i & 1), (double)tsum);
        }
      }
    }
  }
}

} }  // namespace eo::tc
