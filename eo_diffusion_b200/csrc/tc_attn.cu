// Flash-style QKV attention on tcgen05 / TMEM (bf16 operands, fp32 softmax and accumulate).
//
// Replaces QKVAttentionLegacy.forward / QKVAttention.forward (reference
// backbones/unet_openai.py:465-481, :497-515): softmax_fp32((q*s)^T (k*s)) v with
// s = ch^-1/4 (applied here once as ch^-1/2 on the fp32 logits), never materialising the
// [T, T] score matrix in HBM.
//
// Layout: qkv is [B, T, heads*3*64] bf16 -- the qkv 1x1 convolution writes each head's q, k
// and v padded to 64 channels (zero weight rows), so every head dimension <= 64 runs as
// d = 64.  One CTA handles up to TWO tiles of 128 queries of one (batch, head) and streams
// 128-key blocks past them; the two tiles share every K/V load and ping-pong on the tensor
// core, so that one tile's softmax (the MUFU-bound part: 128 exp2 per row per block against
// 512 tensor cycles) overlaps the other tile's MMAs:
//   S_t,j = Q_t K_j^T        tcgen05.mma  M=128 N=128 K=64   -> TMEM S_t
//   P_t,j = exp2(S_t,j - m)  128 softmax threads per tile, one query row each; ONE pass over
//                            TMEM in 32-column chunks using a lazily updated reference
//                            maximum m (rescale only when a logit exceeds m by 2^8);
//                            written as bf16 into a 128B-swizzled K-major smem tile
//   O_t,j = P_t,j V_j        tcgen05.mma  M=128 N=64 K=128, V consumed MN-major straight from
//                            the [key][d] TMA tile -> TMEM O_t, folded into registers
// Warp roles (320 threads): warp 0 TMA loader, warp 1 TMEM owner + MMA issuer, warps 2..5
// softmax of tile 0, warps 6..9 softmax of tile 1.  All hand-offs are mbarriers.
//
// Roofline: tensor pipe / MUFU.  Algorithmic FLOPs per launch = 4 * B * heads * T^2 * ch.
#include "kernels.h"
#include "tc_common.cuh"

namespace eo {

namespace {

constexpr int QT = 128;     // queries per tile
constexpr int KT = 128;     // keys per block
constexpr int HD = 64;      // padded head dim
constexpr int KV_STAGES = 3;
constexpr int TILE_BYTES = 128 * HD * 2;   // 16 KB

constexpr int OFF_Q = 0;                                    // 2 tiles
constexpr int OFF_K = OFF_Q + 2 * TILE_BYTES;
constexpr int OFF_V = OFF_K + KV_STAGES * TILE_BYTES;
constexpr int OFF_P = OFF_V + KV_STAGES * TILE_BYTES;       // 2 tiles x 32 KB
constexpr int OFF_BAR = OFF_P + 2 * 2 * TILE_BYTES;
constexpr int N_BARS = 1 + 2 * KV_STAGES + 2 + 2 + 2;
constexpr int ATTN_SMEM = OFF_BAR + N_BARS * 8 + 16 + 1024;

constexpr uint32_t TM_S0 = 0, TM_S1 = 128, TM_O0 = 256, TM_O1 = 320, TM_COLS = 512;
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 units: P stays <= 2^8

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One 128-key block of one query row: P = exp2(S*scale - m) as bf16 into the swizzled smem row,
// rs = row sum.  m is the lazily updated reference maximum: when some row of the warp sees a
// logit above m + 2^8 the warp takes the exact row maximum as the new reference (alpha carries
// the rescale factor of everything accumulated so far) and redoes the block.  All TMEM loads
// are warp-collective, so every decision here is warp-uniform.
template <bool MASKED>
__device__ __forceinline__ void softmax_block(uint32_t s_addr, uint8_t* prow, int row, int nvalid,
                                              float scale_log2, float& m, float& alpha, float& rs) {
  bool redo = true;
  while (redo) {
    redo = false;
    rs = 0.f;
#pragma unroll 1
    for (int c = 0; c < KT; c += 32) {
      uint32_t v[32];
      tc::tmem_ld_32x32(s_addr + c, v);
      tc::tmem_ld_wait();
      if (MASKED) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c + i >= nvalid) v[i] = 0xff800000u;   // -inf
      }
      float cm = __uint_as_float(v[0]);
#pragma unroll
      for (int i = 1; i < 32; ++i) cm = fmaxf(cm, __uint_as_float(v[i]));
      if (__any_sync(0xffffffffu, cm * scale_log2 > m + RESCALE_THRESHOLD)) {
        float full = -INFINITY;
#pragma unroll 1
        for (int c2 = 0; c2 < KT; c2 += 32) {
          uint32_t w[32];
          tc::tmem_ld_32x32(s_addr + c2, w);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (!MASKED || c2 + i < nvalid) full = fmaxf(full, __uint_as_float(w[i]));
        }
        const float m_new = fmaxf(m, full * scale_log2);
        alpha *= ex2(m - m_new);          // first block: ex2(-inf) = 0
        m = m_new;
        redo = true;
        break;
      }
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float p0 = ex2(fmaf(__uint_as_float(v[i]), scale_log2, -m));
        const float p1 = ex2(fmaf(__uint_as_float(v[i + 1]), scale_log2, -m));
        rs += p0 + p1;
        __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
        pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h2);
      }
      // 32 keys = 64 B = four 16 B pieces; piece index within the 128 B row: (c%64)/8 + jj
      uint8_t* half = prow + (c >> 6) * TILE_BYTES;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int piece = ((c & 63) >> 3) + jj;
        uint4 w4 = make_uint4(pk[jj * 4], pk[jj * 4 + 1], pk[jj * 4 + 2], pk[jj * 4 + 3]);
        *reinterpret_cast<uint4*>(half + ((piece ^ (row & 7)) << 4)) = w4;
      }
    }
  }
}

__global__ void __launch_bounds__(320, 1)
k_attn_tc(const __grid_constant__ CUtensorMap map_qkv, __nv_bfloat16* __restrict__ out, int T,
          int heads, int ch, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (tc::smem_u32(smem_raw) & 1023)) & 1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + KV_STAGES;
  uint64_t* s_full = kv_empty + KV_STAGES;   // [2] per tile
  uint64_t* p_full = s_full + 2;             // [2] per tile, 128 arrivals
  uint64_t* o_full = p_full + 2;             // [2] per tile
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * QT, h = blockIdx.y, b = blockIdx.z;
  const int ntiles = (T - q0 > QT) ? 2 : 1;
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
      tc::mbar_arrive_expect_tx(q_full, ntiles * TILE_BYTES);
      for (int t = 0; t < ntiles; ++t)
        tc::tma_load_3d(smem + OFF_Q + t * TILE_BYTES, &map_qkv, q_full, cq, q0 + t * QT, b);
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
      // S_t,j = Q_t K_j^T into TMEM S_t (single buffer: issued only after P_t,j-1 was handed over)
      auto issue_s = [&](int t, int j) {
        const int s = j % KV_STAGES;
        const uint64_t qdesc = tc::make_sw128_desc(tc::smem_u32(smem + OFF_Q + t * TILE_BYTES));
        const uint64_t kdesc = tc::make_sw128_desc(tc::smem_u32(smem + OFF_K + s * TILE_BYTES));
        const uint32_t d = tmem + (t ? TM_S1 : TM_S0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          tc::umma_f16_ss(d, tc::desc_advance(qdesc, k * 32), tc::desc_advance(kdesc, k * 32),
                          idesc_s, k != 0 ? 1u : 0u);
        tc::umma_commit(&s_full[t]);
      };
      // O_t,j = P_t,j V_j into TMEM O_t (overwrite; the softmax threads fold it into registers)
      auto issue_pv = [&](int t, int j) {
        const int s = j % KV_STAGES;
        const uint64_t vdesc = tc::make_sw128_desc(tc::smem_u32(smem + OFF_V + s * TILE_BYTES));
        const uint32_t pbase = tc::smem_u32(smem + OFF_P + t * 2 * TILE_BYTES);
        const uint32_t d = tmem + (t ? TM_O1 : TM_O0);
#pragma unroll
        for (int k = 0; k < KT / 16; ++k) {
          // A: P tile, half k/4 (64 keys = one 128 B row segment), 32 B per 16 keys
          const uint64_t pdesc = tc::make_sw128_desc(pbase + (k >> 2) * TILE_BYTES + (k & 3) * 32);
          // B: V tile [key][d], 16 keys = 16 rows of 128 B
          tc::umma_f16_ss(d, pdesc, tc::desc_advance(vdesc, k * 16 * 128), idesc_o, k != 0 ? 1u : 0u);
        }
        tc::umma_commit(&o_full[t]);
      };
      tc::mbar_wait(q_full, 0);
      tc::mbar_wait(&kv_full[0], 0);
      tc::tc_fence_after();
      for (int t = 0; t < ntiles; ++t) issue_s(t, 0);
      for (int j = 0; j < nblk; ++j) {
        const bool more = j + 1 < nblk;
        if (more) {
          tc::mbar_wait(&kv_full[(j + 1) % KV_STAGES], ((j + 1) / KV_STAGES) & 1);
          tc::tc_fence_after();
        }
        for (int t = 0; t < ntiles; ++t) {
          tc::mbar_wait(&p_full[t], j & 1);
          tc::tc_fence_after();
          issue_pv(t, j);
          if (t == ntiles - 1) tc::umma_commit(&kv_empty[j % KV_STAGES]);   // K_j, V_j consumed
          if (more) issue_s(t, j + 1);
        }
      }
    }
  } else if (warp < 2 + 4 * ntiles) {
    // ---- softmax / output warps: thread <-> query row of tile t
    const int t = (warp - 2) >> 2;
    const int qd = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    const uint32_t s_addr = tmem + lane_addr + (t ? TM_S1 : TM_S0);
    const uint32_t o_addr = tmem + lane_addr + (t ? TM_O1 : TM_O0);
    uint8_t* prow = smem + OFF_P + t * 2 * TILE_BYTES + row * 128;
    float acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = 0.f;
    float m = -INFINITY, l = 0.f;       // m: reference maximum (log2 domain) of everything folded so far
    const int last_valid = T - (nblk - 1) * KT;   // valid keys in the last block

    for (int j = 0; j < nblk; ++j) {
      tc::mbar_wait(&s_full[t], j & 1);
      tc::tc_fence_after();
      if (j > 0) {
        // S_t,j is issued after PV_t,j-1, so O_t,j-1 is complete: fold it (scale of m as of block j-1)
        tc::mbar_wait(&o_full[t], (j - 1) & 1);
        tc::tc_fence_after();
#pragma unroll
        for (int c = 0; c < HD; c += 32) {
          uint32_t v[32];
          tc::tmem_ld_32x32(o_addr + c, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[c + i] += __uint_as_float(v[i]);
        }
      }
      float alpha = 1.f, rs = 0.f;
      if (j == nblk - 1 && last_valid < KT)
        softmax_block<true>(s_addr, prow, row, last_valid, scale_log2, m, alpha, rs);
      else
        softmax_block<false>(s_addr, prow, row, KT, scale_log2, m, alpha, rs);
      // hand P_t,j (and the consumed S_t) to the MMA warp
      tc::fence_proxy_async_smem();
      tc::tc_fence_before();
      tc::mbar_arrive(&p_full[t]);
      if (alpha != 1.f) {
#pragma unroll
        for (int d = 0; d < HD; ++d) acc[d] *= alpha;
        l *= alpha;
      }
      l += rs;
    }
    // last block's PV
    tc::mbar_wait(&o_full[t], (nblk - 1) & 1);
    tc::tc_fence_after();
    const int qi = q0 + t * QT + row;
    const float inv = 1.0f / l;
    __nv_bfloat16* op = out + ((long long)b * T + qi) * (heads * ch) + h * ch;
#pragma unroll
    for (int c = 0; c < HD; c += 32) {
      uint32_t v[32];
      tc::tmem_ld_32x32(o_addr + c, v);
      tc::tmem_ld_wait();
      if (qi < T) {
#pragma unroll
        for (int d = 0; d < 32; d += 8) {
          if (c + d < ch) {
            uint4 w4;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&w4);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              h2[e] = __floats2bfloat162_rn((acc[c + d + 2 * e] + __uint_as_float(v[d + 2 * e])) * inv,
                                            (acc[c + d + 2 * e + 1] + __uint_as_float(v[d + 2 * e + 1])) * inv);
            *reinterpret_cast<uint4*>(op + c + d) = w4;
          }
        }
      }
    }
    tc::tc_fence_before();
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
  dim3 grid((unsigned)ceil_div(p.T, 2 * QT), (unsigned)p.heads, (unsigned)B);
  k_attn_tc<<<grid, 320, ATTN_SMEM, st>>>(pl->map, reinterpret_cast<__nv_bfloat16*>(p.out), p.T,
                                          p.heads, p.ch, scale_log2);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

}  // namespace eo
