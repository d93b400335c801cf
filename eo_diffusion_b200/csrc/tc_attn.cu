// Flash-style QKV attention on tcgen05 / TMEM (bf16 operands, fp32 softmax and accumulate).
//
// Replaces QKVAttentionLegacy.forward / QKVAttention.forward (reference
// backbones/unet_openai.py:465-481, :497-515): softmax_fp32((q*s)^T (k*s)) v with
// s = ch^-1/4 (applied here once as ch^-1/2 on the fp32 logits), never materialising the
// [T, T] score matrix in HBM.
//
// Layout: qkv is [B, T, heads*3*64] bf16 -- the qkv 1x1 convolution writes each head's q, k
// and v padded to 64 channels (zero weight rows), so every head dimension <= 64 runs as
// d = 64.  One CTA handles 128 queries of one (batch, head) and loops over 128-key blocks:
//   S_j = Q K_j^T          tcgen05.mma  M=128 N=128 K=64   -> TMEM (double buffered)
//   P_j = exp2(S_j - m_j)  128 softmax threads, one query row each (tcgen05.ld 32x32b),
//                          written as bf16 into a 128B-swizzled K-major smem tile
//   O_j = P_j V_j          tcgen05.mma  M=128 N=64  K=128, V consumed MN-major straight from
//                          the [key][d] TMA tile -> TMEM scratch (double buffered)
//   acc = (acc + O_{j-1}) * exp2(m_{j-1} - m_j)   in registers (fp32)
// Warp roles (192 threads): warp 0 TMA loader, warp 1 TMEM owner + MMA issuer, warps 2..5
// softmax / output.  All hand-offs are mbarriers; there is no __syncthreads in the loop.
//
// Roofline: tensor pipe + MUFU.  Algorithmic FLOPs per launch = 4 * B * heads * T^2 * ch.
#include "kernels.h"
#include "tc_common.cuh"

namespace eo {

namespace {

constexpr int QT = 128;     // queries per CTA
constexpr int KT = 128;     // keys per block
constexpr int HD = 64;      // padded head dim
constexpr int KV_STAGES = 3;
constexpr int TILE_BYTES = 128 * HD * 2;   // 16 KB

constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + TILE_BYTES;
constexpr int OFF_V = OFF_K + KV_STAGES * TILE_BYTES;
constexpr int OFF_P = OFF_V + KV_STAGES * TILE_BYTES;       // 2 buffers x 32 KB
constexpr int OFF_BAR = OFF_P + 2 * 2 * TILE_BYTES;
constexpr int N_BARS = 1 + 2 * KV_STAGES + 2 + 2 + 2;
constexpr int ATTN_SMEM = OFF_BAR + N_BARS * 8 + 16 + 1024;

constexpr uint32_t TM_S0 = 0, TM_S1 = 128, TM_O0 = 256, TM_O1 = 320, TM_COLS = 512;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(192, 1)
k_attn_tc(const __grid_constant__ CUtensorMap map_qkv, __nv_bfloat16* __restrict__ out, int T,
          int heads, int ch, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (tc::smem_u32(smem_raw) & 1023)) & 1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + KV_STAGES;
  uint64_t* s_full = kv_empty + KV_STAGES;   // [2]
  uint64_t* p_full = s_full + 2;             // [2], 128 arrivals
  uint64_t* o_full = p_full + 2;             // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * QT, h = blockIdx.y, b = blockIdx.z;
  const int nblk = (T + KT - 1) / KT;
  const int cq = h * 3 * HD, ck = cq + HD, cv = cq + 2 * HD;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&map_qkv);
    tc::mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) { tc::mbar_init(&kv_full[s], 1); tc::mbar_init(&kv_empty[s], 1); }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&s_full[i], 1); tc::mbar_init(&p_full[i], 128); tc::mbar_init(&o_full[i], 1);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) { tc::tmem_alloc(tmem_ptr, TM_COLS); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tc::tma_load_3d(smem + OFF_Q, &map_qkv, q_full, cq, q0, b);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        tc::mbar_wait(&kv_empty[s], ph ^ 1);
        tc::mbar_arrive_expect_tx(&kv_full[s], 2 * TILE_BYTES);
        tc::tma_load_3d(smem + OFF_K + s * TILE_BYTES, &map_qkv, &kv_full[s], ck, j * KT, b);
        tc::tma_load_3d(smem + OFF_V + s * TILE_BYTES, &map_qkv, &kv_full[s], cv, j * KT, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = tc::make_idesc_bf16(128, KT, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = tc::make_idesc_bf16(128, HD, 0, 1);   // P (K-major) x V (MN-major)
      const uint64_t qdesc = tc::make_sw128_desc(tc::smem_u32(smem + OFF_Q));
      auto issue_s = [&](int j) {
        const int s = j % KV_STAGES;
        tc::mbar_wait(&kv_full[s], (j / KV_STAGES) & 1);
        tc::tc_fence_after();
        const uint64_t kdesc = tc::make_sw128_desc(tc::smem_u32(smem + OFF_K + s * TILE_BYTES));
        const uint32_t d = tmem + ((j & 1) ? TM_S1 : TM_S0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          tc::umma_f16_ss(d, tc::desc_advance(qdesc, k * 32), tc::desc_advance(kdesc, k * 32),
                          idesc_s, k != 0 ? 1u : 0u);
        tc::umma_commit(&s_full[j & 1]);
      };
      tc::mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) issue_s(j + 1);
        tc::mbar_wait(&p_full[j & 1], (j >> 1) & 1);
        tc::tc_fence_after();
        const int s = j % KV_STAGES;
        const uint64_t vdesc = tc::make_sw128_desc(tc::smem_u32(smem + OFF_V + s * TILE_BYTES));
        const uint32_t pbase = tc::smem_u32(smem + OFF_P + (j & 1) * 2 * TILE_BYTES);
        const uint32_t d = tmem + ((j & 1) ? TM_O1 : TM_O0);
#pragma unroll
        for (int k = 0; k < KT / 16; ++k) {
          // A: P tile, chunk k/4 (64 keys = one 128 B row segment), 32 B per 16 keys
          const uint64_t pdesc = tc::make_sw128_desc(pbase + (k >> 2) * TILE_BYTES + (k & 3) * 32);
          // B: V tile [key][d], 16 keys = 16 rows of 128 B
          tc::umma_f16_ss(d, pdesc, tc::desc_advance(vdesc, k * 16 * 128), idesc_o, k != 0 ? 1u : 0u);
        }
        tc::umma_commit(&kv_empty[s]);
        tc::umma_commit(&o_full[j & 1]);
      }
    }
  } else {
    // ---- softmax / output warps: thread <-> query row
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    float acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = 0.f;
    float m = -INFINITY, l = 0.f;
    const bool ragged = (T % KT) != 0;

    auto add_o = [&](int j) {   // acc += O_j
      tc::mbar_wait(&o_full[j & 1], (j >> 1) & 1);
      tc::tc_fence_after();
      const uint32_t src = tmem + lane_addr + ((j & 1) ? TM_O1 : TM_O0);
#pragma unroll
      for (int c = 0; c < HD; c += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32(src + c, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[c + i] += __uint_as_float(v[i]);
      }
    };

    for (int j = 0; j < nblk; ++j) {
      __syncwarp();
      tc::mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc::tc_fence_after();
      const uint32_t src = tmem + lane_addr + ((j & 1) ? TM_S1 : TM_S0);
      const int kbase = j * KT;
      // pass 1: row maximum of the scaled logits (log2 domain)
      float bm = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < KT; c += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32(src + c, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float sv = __uint_as_float(v[i]) * scale_log2;
          if (ragged && kbase + c + i >= T) sv = -INFINITY;
          bm = fmaxf(bm, sv);
        }
      }
      const float m_new = fmaxf(m, bm);
      const float alpha = ex2(m - m_new);          // first block: ex2(-inf) = 0
      // pass 2: P = exp2(S - m_new) -> bf16 -> swizzled smem; row sum
      float rs = 0.f;
      uint8_t* prow = smem + OFF_P + (j & 1) * 2 * TILE_BYTES + row * 128;
#pragma unroll 1
      for (int c = 0; c < KT; c += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32(src + c, v);
        tc::tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float s0 = __uint_as_float(v[i]) * scale_log2;
          float s1 = __uint_as_float(v[i + 1]) * scale_log2;
          if (ragged) {
            if (kbase + c + i >= T) s0 = -INFINITY;
            if (kbase + c + i + 1 >= T) s1 = -INFINITY;
          }
          float p0 = ex2(s0 - m_new), p1 = ex2(s1 - m_new);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
          // the row sum uses the bf16-rounded probabilities that the PV product will see
          float2 r = __bfloat1622float2(h2);
          rs += r.x + r.y;
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h2);
        }
        // 32 keys = 64 B = four 16 B pieces; piece index within the 128 B row: (c%64)/8 + jj
        uint8_t* chunk = prow + (c >> 6) * TILE_BYTES;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int piece = ((c & 63) >> 3) + jj;
          uint4 w = make_uint4(pk[jj * 4], pk[jj * 4 + 1], pk[jj * 4 + 2], pk[jj * 4 + 3]);
          *reinterpret_cast<uint4*>(chunk + ((piece ^ (row & 7)) << 4)) = w;
        }
      }
      // hand P_j (and the consumed S buffer) to the MMA warp
      tc::fence_proxy_async_smem();
      tc::tc_fence_before();
      tc::mbar_arrive(&p_full[j & 1]);
      // fold the previous block's PV result, then rescale to the new maximum
      if (j > 0) add_o(j - 1);
      l = l * alpha + rs;
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[d] *= alpha;
      m = m_new;
    }
    add_o(nblk - 1);
    tc::tc_fence_before();
    const int qi = q0 + row;
    if (qi < T) {
      const float inv = 1.0f / l;
      __nv_bfloat16* op = out + ((long long)b * T + qi) * (heads * ch) + h * ch;
#pragma unroll
      for (int d = 0; d < HD; d += 8) {     // fully unrolled: acc[] must stay in registers
        if (d < ch) {
          uint4 w;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            h2[e] = __floats2bfloat162_rn(acc[d + 2 * e] * inv, acc[d + 2 * e + 1] * inv);
          *reinterpret_cast<uint4*>(op + d) = w;
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, TM_COLS);
  }
}

}  // namespace

struct TcAttnPlan {
  CUtensorMap map;
  TcAttnParams p;
};

int tc_attn_plan_create(const TcAttnParams& p, TcAttnPlan** out) {
  EO_REQUIRE(p.ch <= HD && p.ch % 8 == 0, EO_ERR_ARG,
             "tc_attn: head dimension %d unsupported (must be a multiple of 8, <= 64)", p.ch);
  TcAttnPlan* pl = new TcAttnPlan();
  pl->p = p;
  uint64_t ld = (uint64_t)p.heads * 3 * HD;
  uint64_t dims[3] = {ld, (uint64_t)p.T, (uint64_t)p.B};
  uint64_t str[2] = {ld * 2, (uint64_t)p.T * ld * 2};
  uint32_t box[3] = {(uint32_t)HD, 128u, 1u};
  int rc = encode_tmap_bf16(&pl->map, p.qkv, 3, dims, str, box);
  if (rc != EO_OK) { delete pl; return rc; }
  *out = pl;
  return EO_OK;
}

void tc_attn_plan_destroy(TcAttnPlan* p) { delete p; }

int tc_attn_launch(const TcAttnPlan* pl, int B, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    EO_CHECK_CUDA(cudaFuncSetAttribute(k_attn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       ATTN_SMEM));
    attr_set = true;
  }
  const TcAttnParams& p = pl->p;
  // logits = (q . k) * ch^-1/2 ; softmax evaluated with exp2
  float scale_log2 = (1.0f / sqrtf((float)p.ch)) * 1.4426950408889634f;
  dim3 grid((unsigned)ceil_div(p.T, QT), (unsigned)p.heads, (unsigned)B);
  k_attn_tc<<<grid, 192, ATTN_SMEM, st>>>(pl->map, reinterpret_cast<__nv_bfloat16*>(p.out), p.T,
                                          p.heads, p.ch, scale_log2);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

}  // namespace eo
