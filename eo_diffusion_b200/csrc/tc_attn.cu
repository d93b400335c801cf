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
//   P_t,j = exp2(S_t,j - m)  128 softmax threads per tile, one query row each: the whole 128-logit
//                            row is pulled into registers with one batch of tcgen05.ld, m is a
//                            lazily updated reference maximum (O and l are rescaled only when a
//                            logit exceeds m by 2^8); P is written as bf16 into a 128B-swizzled
//                            K-major smem tile
//   O_t  += P_t,j V_j        tcgen05.mma  M=128 N=64 K=128, V consumed MN-major straight from
//                            the [key][d] TMA tile; O_t accumulates in TMEM across all blocks
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

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t v[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
         "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
         "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
         "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
         "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
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
      // O_t (+)= P_t,j V_j in TMEM; the softmax threads rescale O_t in place when m moves
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
          tc::umma_f16_ss(d, pdesc, tc::desc_advance(vdesc, k * 16 * 128), idesc_o, (j | k) != 0 ? 1u : 0u);
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
          // p_full: P_t,j is in smem AND S_t has been consumed -> the next S first (the softmax
          // of block j+1 waits on it), then this block's PV
          tc::mbar_wait(&p_full[t], j & 1);
          tc::tc_fence_after();
          if (more) issue_s(t, j + 1);
          issue_pv(t, j);
          if (t == ntiles - 1) tc::umma_commit(&kv_empty[j % KV_STAGES]);   // K_j, V_j consumed
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
    float m = -INFINITY, l = 0.f;       // m: reference maximum (log2 domain) of everything accumulated so far
    const int last_valid = T - (nblk - 1) * KT;   // valid keys in the last block

    for (int j = 0; j < nblk; ++j) {
      tc::mbar_wait(&s_full[t], j & 1);
      tc::tc_fence_after();
      const int nvalid = (j == nblk - 1) ? last_valid : KT;   // < KT only in a ragged last block
      // ---- pass 1: row maximum.  32-column chunks, the next chunk's tcgen05.ld in flight while
      // this one is reduced (two register buffers; wait::ld covers the single outstanding load)
      uint32_t va[32], vb[32];
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      tc::tmem_ld_32x32(s_addr, va);
      tc::tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < KT; c += 32) {
        uint32_t* cur = (c & 32) ? vb : va;
        uint32_t* nxt = (c & 32) ? va : vb;
        if (c + 32 < KT) tc::tmem_ld_32x32(s_addr + c + 32, nxt);
        if (nvalid < KT) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i >= nvalid) cur[i] = 0xff800000u;   // -inf
        }
#pragma unroll
        for (int i = 0; i < 32; i += 8)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mx[k] = max3(mx[k], __uint_as_float(cur[i + 2 * k]), __uint_as_float(cur[i + 2 * k + 1]));
        tc::tmem_ld_wait();
      }
      const float cm = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * scale_log2;
      // warp-uniform decision (the TMEM accesses below are warp-collective)
      if (__any_sync(0xffffffffu, cm > m + RESCALE_THRESHOLD)) {
        const float m_new = fmaxf(m, cm);
        const float alpha = ex2(m - m_new);          // first block: ex2(-inf) = 0
        m = m_new;
        l *= alpha;
        if (j > 0) {
          // O_t holds blocks < j (PV_t,j-1 has completed: o_full); rescale it in place
          tc::mbar_wait(&o_full[t], (j - 1) & 1);
          tc::tc_fence_after();
#pragma unroll
          for (int c = 0; c < HD; c += 32) {
            uint32_t o[32];
            tc::tmem_ld_32x32(o_addr + c, o);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(o_addr + c, o);
          }
          tmem_st_wait();
        }
      }
      // P smem is free once PV_t,j-1 has read it (issued a whole softmax ago: no real wait)
      if (j > 0) tc::mbar_wait(&o_full[t], (j - 1) & 1);
      // ---- pass 2: P = exp2(S*scale - m) -> bf16 -> swizzled smem; row sum
      float rs0 = 0.f, rs1 = 0.f;
      tc::tmem_ld_32x32(s_addr, va);
      tc::tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < KT; c += 32) {
        uint32_t* cur = (c & 32) ? vb : va;
        uint32_t* nxt = (c & 32) ? va : vb;
        if (c + 32 < KT) tc::tmem_ld_32x32(s_addr + c + 32, nxt);
        if (nvalid < KT) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i >= nvalid) cur[i] = 0xff800000u;
        }
        // 32 keys = 64 B = four 16 B pieces of the 128 B row of the 64-key half c/64, XOR-swizzled
        uint8_t* half = prow + (c >> 6) * TILE_BYTES;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = jj * 8 + 2 * e;
            const float p0 = ex2(fmaf(__uint_as_float(cur[i]), scale_log2, -m));
            const float p1 = ex2(fmaf(__uint_as_float(cur[i + 1]), scale_log2, -m));
            rs0 += p0; rs1 += p1;
            __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
            pk[e] = *reinterpret_cast<uint32_t*>(&h2);
          }
          const int piece = ((c & 63) >> 3) + jj;
          *reinterpret_cast<uint4*>(half + ((piece ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        tc::tmem_ld_wait();
      }
      // hand P_t,j (and the consumed S_t, and the possibly rescaled O_t) to the MMA warp
      tc::fence_proxy_async_smem();
      tc::tc_fence_before();
      tc::mbar_arrive(&p_full[t]);
      l += rs0 + rs1;
    }
    tc::mbar_wait(&o_full[t], (nblk - 1) & 1);
    tc::tc_fence_after();
    const int qi = q0 + t * QT + row;
    const float inv = 1.0f / l;
    __nv_bfloat16* op = out + ((long long)b * T + qi) * (heads * ch) + h * ch;
#pragma unroll
    for (int c = 0; c < HD; c += 32) {
      uint32_t o[32];
      tc::tmem_ld_32x32(o_addr + c, o);
      tc::tmem_ld_wait();
      if (qi < T) {
#pragma unroll
        for (int d = 0; d < 32; d += 8) {
          if (c + d < ch) {
            uint4 w4;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&w4);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              h2[e] = __floats2bfloat162_rn(__uint_as_float(o[d + 2 * e]) * inv,
                                            __uint_as_float(o[d + 2 * e + 1]) * inv);
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
