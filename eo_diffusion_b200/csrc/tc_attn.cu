// Flash-style QKV attention on tcgen05 / TMEM (bf16 operands, fp32 softmax and accumulate).
//
// Replaces QKVAttentionLegacy.forward / QKVAttention.forward (reference
// backbones/unet_openai.py:465-481, :497-515): softmax_fp32((q*s)^T (k*s)) v with
// s = ch^-1/4 (applied once as ch^-1/2: on the fp32 logits for 64-channel heads, folded into the q
// projection for heads of <= 48 channels), never materialising the [T, T] score matrix in HBM.
//
// Layout: qkv is [B, T, heads*3*64] bf16 -- the qkv 1x1 convolution writes each head's q, k
// and v padded to 64 channels (zero weight rows); wider heads take simt.cu's k_attention_wide.
// The kernel (k_attn_tc6) is described at its definition.
//
// Roofline: MUFU + issue slots (one ex2 per logit, 4*ch FLOPs per logit); algorithmic FLOPs per launch =
// 4 * B * heads * T^2 * ch.
#include "kernels.h"
#include "tc_common.cuh"
#include <cstdlib>
#include <type_traits>

namespace eo {

namespace {

constexpr int QT = 128;     // queries per tile
constexpr int KT = 128;     // keys per K/V tile
constexpr int HD = 64;      // padded head dim
constexpr int TILE_BYTES = 128 * HD * 2;   // 16 KB

constexpr uint32_t TM_COLS = 512;
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 units: P stays <= 2^8

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA pipe (a share of the exponentials is taken off the MUFU pipe): x = n + f with n = round(x),
// f in [-0.5, 0.5]; 2^f by a degree-3 minimax polynomial (relative error 7.5e-5, 26 times below the bf16 rounding
// of P), 2^n added into the exponent field.  x <= 127 - 1; below -125 clamps.  ex2_fma2 below does a pair.
// Packed fp32 pairs (Blackwell FFMA2 / FADD2: one issue slot for two lanes' worth of a pair): the scale / subtract of
// every logit pair and the polynomial form of 2^x run on these, halving the issue slots they take next to the MUFU pipe.
using tc::pack2; using tc::unpack2; using tc::fma2; using tc::add2;
__device__ __forceinline__ void ex2_fma2(uint64_t x, float& p0, float& p1) {
  float x0, x1;
  unpack2(x, x0, x1);
  x = pack2(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
  const uint64_t t = add2(x, pack2(12582912.0f, 12582912.0f));
  const uint64_t r = add2(t, pack2(-12582912.0f, -12582912.0f));
  const uint64_t f = fma2(r, pack2(-1.0f, -1.0f), x);
  uint64_t p = fma2(pack2(0.0551716685295105f, 0.0551716685295105f), f, pack2(0.2426111251115799f, 0.2426111251115799f));
  p = fma2(p, f, pack2(0.6932609677314758f, 0.6932609677314758f));
  p = fma2(p, f, pack2(0.9999280571937561f, 0.9999280571937561f));
  float q0, q1, t0, t1;
  unpack2(p, q0, q1);
  unpack2(t, t0, t1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// development aid (eo_debug_conv_trace, TRACE instantiation, -DEO_DEVTOOLS builds): per-CTA clock64 totals
#define ATR_T0() const long long _t0 = TRACE ? clock64() : 0
#define ATR_ACC(var) do { if (TRACE) (var) += clock64() - _t0; } while (0)

constexpr int NTILE = 4;
constexpr int KQ = 32;                      // keys per quarter-block
constexpr int KV_STAGES = 4;
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + NTILE * TILE_BYTES;
constexpr int OFF_V = OFF_K + KV_STAGES * TILE_BYTES;
constexpr int OFF_BAR = OFF_V + KV_STAGES * TILE_BYTES;
constexpr int ATTN_THREADS = 768;
constexpr int SM_WARP0 = 8;
// POLY_NUM of every POLY_DEN pairs of exponentials run on the FMA pipe (0 = none)
#ifndef EO_ATTN_POLY_NUM
#define EO_ATTN_POLY_NUM 1
#endif
#ifndef EO_ATTN_POLY_DEN
#define EO_ATTN_POLY_DEN 4
#endif
constexpr int POLY_NUM = EO_ATTN_POLY_NUM, POLY_DEN = EO_ATTN_POLY_DEN;

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t v[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t v[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
         "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}

// =====================================================================================================
// k_attn_tc6: FOUR 128-query tiles per CTA, keys consumed in 32-key quarter-blocks (rounds).
//
// 768 threads (registers redistributed with setmaxnreg): warps 0..3 MMA issuers of tile 0..3 (one thread each), 4 TMA
// loader (K/V tiles of 128 keys, 4 stages; 5..7 idle), 8..23 softmax (tile (w-8)/4, TMEM lane quadrant w % 4, thread <->
// query row): every warp scheduler holds one softmax warp of each tile.  Per round and tile: S_q = Q K_q^T (M128 N32)
// into the tile's S columns; the softmax thread of a row takes P = exp2(S * scale - m) against an INTEGER-valued lazily
// raised reference maximum m (O and l rescaled in place by an exact power of two, only when a logit exceeds m by 2^8)
// and writes P as packed bf16 pairs into the tile's P columns; O += P V_q is a TS MMA (A = P from tensor memory, V
// MN-major straight from its TMA tile).  A share of the exponentials runs on the FMA pipe (ex2_fma2, packed fp32x2).
//
// What shaped it (round 2; tools/attn_timeline.py = a clock64 timeline of the softmax warps of one scheduler and of
// the issuers): the predecessor handed S and P back and forth through one aliased buffer pair and ran its four tiles
// in LOCKSTEP -- every softmax warp of a scheduler in the same phase at the same time (all four queueing for the MUFU
// pipe, then all four in barrier wait / tcgen05.ld / maximum / vote with the MUFU pipe idle), while the issuer warps,
// which get every fifth issue slot of their scheduler at best, needed 500-700 clk to get five MMAs out.  Hence:
//   * the softmax is software-pipelined across rounds: the row of S_q+1 is fetched into a second register set behind
//     the first exponentials of S_q (tcgen05.ld is asynchronous until its registers are read), barrier tests are
//     issued early and consumed late, the maximum / vote on S_q+1 covers the latency of the tcgen05.st of P_q -- a
//     warp's own arithmetic hides its latencies, lockstep or not;
//   * S has ONE 32-column buffer per tile, released when its row is in registers (s_free), and P a buffer of its
//     own; the issuer puts S_q+2 ahead of P V_q (S is what the softmax needs next, P V_q only has to land before
//     P_q+1 is written);
//   * the issuer is ONE thread running an unrolled loop: ~45 instructions per round instead of ~110;
//   * TSQ (head dimension <= 48, the T = 4096 blocks): the softmax threads copy their Q row into tensor memory once, so
//     S = Q K^T is a TS MMA (16 clk per K step instead of 40: an SS MMA at N = 32 re-reads Q from shared memory), O
//     is only round16(ch) columns wide, and one extra K step carries (-m) x 1.0: q arrives multiplied by the logit
//     scale (folded into the qkv projection, TcAttnParams::k_one), the thread keeps -m in a spare channel of its Q
//     row, k has 1.0 there -- the tensor core delivers S * scale - m and the softmax spends no FFMA2 per logit pair
//     on it.  A raised maximum is written to the Q column BEFORE s_free lets the next S be issued; the one row
//     already in registers gets the difference added (corr).  64-channel heads keep Q in shared memory, a 64-column
//     O and the FFMA2.
// Tensor memory per tile (128 columns): S 0..31, P 32..47 (packed bf16 pairs), O 48..48+ON, Q 96..96+ON/2+8 (TSQ).
// Barriers per tile: s_full (commit of S_q), s_free (128 threads: row read), p_full (128 threads: P_q written),
// pv_done (commit of P V_q; the softmax waits for P V_q-1 before it overwrites P or rescales O), q_ready (4 warps: Q
// copied).  The row sum l is accumulated by the softmax threads (packed fp32 adds).
// Measured (B200, 64 x 8 heads, T = 4096, ch = 48, kernel alone at ~1.9 GHz): predecessor 2.42 ms; pipelined softmax +
// lean issuer 2.28; + baked scale / maximum 2.12 (778 TFLOP/s algorithmic).  ncu: MUFU pipe 71 %, issue slots 62 %.
// =====================================================================================================
// q_full, kv_full[S], kv_empty[S], then per tile: s_full, s_free, p_full, pv_done, q_ready
constexpr int N_BARS = 1 + 2 * KV_STAGES + 5 * NTILE;
constexpr int ATTN_SMEM = OFF_BAR + N_BARS * 8 + 16 + 1024;
constexpr int TM6_S = 0, TM6_P = 32, TM6_O = 48, TM6_Q = 96;      // Q: up to 3 + 1 K steps of 8 columns

__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t v[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

// wait without the trap counter of tc::mbar_wait: the single issuer thread competes with four softmax warps for its
// scheduler's issue slots, every instruction in its loop counts
__device__ __forceinline__ void mbar_wait_lean(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra WAIT_%=;\n\t}"
      :: "r"(tc::smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void tmem_st_32x1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(taddr), "r"(v) : "memory");
}

template <bool TRACE, int ON>      // ON: columns of O = round16(ch); Q in tensor memory below 64
__global__ void __launch_bounds__(ATTN_THREADS, 1)
k_attn_tc6(const __grid_constant__ CUtensorMap map_qkv, __nv_bfloat16* __restrict__ out, int T,
           int heads, int ch, float scale_log2, long long* trace, int trace_n) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (tc::smem_u32(smem_raw) & 1023)) & 1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + KV_STAGES;
  uint64_t* s_full = kv_empty + KV_STAGES;   // [tile]
  uint64_t* s_free = s_full + NTILE;
  uint64_t* p_full = s_free + NTILE;
  uint64_t* pv_done = p_full + NTILE;
  uint64_t* q_ready = pv_done + NTILE;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(q_ready + NTILE);

  // warp index through a shuffle: the compiler then knows it (and the tile, quadrant, barrier and tensor-memory
  // addresses derived from it) to be warp-uniform and keeps them in uniform registers
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const long long t_entry = TRACE ? clock64() : 0;
  const int cta_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  long long* trc = (TRACE && cta_lin < trace_n) ? trace + (long long)cta_lin * 8 : nullptr;
  // timeline (TRACE only): rows behind the per-CTA totals, CTA TL_CTA, rounds TL_Q0 .. TL_Q0 + 15: softmax warps
  // [tile][quadrant][round][8] = loop top, first half done, next row requested, second half done, next row there,
  // P V_q-1 seen, P handed over, next maximum known; then issuers
  // [tile][round][4] = round starts, next S issued, saw P, P V issued
  constexpr int TL_CTA = 150, TL_Q0 = 48, TL_N = 16;
  long long* tl = (TRACE && cta_lin == TL_CTA && trace_n > TL_CTA) ? trace + (long long)trace_n * 8 : nullptr;
  const int q0 = blockIdx.x * NTILE * QT, h = blockIdx.y, b = blockIdx.z;
  const int ntiles = min(NTILE, (T - q0 + QT - 1) / QT);
  const int nblk = (T + KT - 1) / KT;
  const int nq = (T + KQ - 1) / KQ;          // quarter-blocks that hold at least one key
  const int cq = h * 3 * HD, ck = cq + HD, cv = cq + 2 * HD;
  constexpr bool TSQ = ON < HD;
  constexpr int on = ON;
  // K steps of S: the channels (the padded ones of q and k are zero) and, with Q in tensor memory, one more whose first
  // channel is -m in Q (written by the softmax thread of the row) and 1.0 in K (the qkv convolution's bias on that padded
  // row): the tensor core delivers S * scale - m, the softmax does not spend an FFMA2 per logit pair on it
  constexpr int ksteps = ON / 16 + (TSQ ? 1 : 0);

  if (warp == 4 && lane == 0) {
    tc::tma_prefetch_desc(&map_qkv);
    tc::mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) { tc::mbar_init(&kv_full[s], 1); tc::mbar_init(&kv_empty[s], ntiles); }
    for (int i = 0; i < NTILE; ++i) {
      tc::mbar_init(&s_full[i], 1); tc::mbar_init(&s_free[i], 128); tc::mbar_init(&p_full[i], 128);
      tc::mbar_init(&pv_done[i], 1); tc::mbar_init(&q_ready[i], 4);
    }
    tc::fence_barrier_init();
  }
  if (warp == 0) { tc::tmem_alloc(tmem_ptr, TM_COLS); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  pdl_wait();            // the qkv convolution may still have been running while this CTA set itself up (common.cuh)
  pdl_trigger();

  // 768 threads leave 80 registers each: the issue warpgroup drops to 40, the loader's to 24, the four softmax
  // warpgroups grow to 104 (40 + 24 + 4 * 104 = 480 = 6 * 80; the pool is what the CTA was launched with)
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    // ---------------------------------------------------------------- MMA issuer of tile `warp`: ONE thread runs the
    // whole loop.  The issuer shares its scheduler with four busy softmax warps and gets every fifth issue slot at
    // best, so every instruction in its loop counts: unrolled over the four quarter-blocks of a K/V tile, every
    // descriptor is a constant offset from one base and every parity a constant.
    if (warp < ntiles && tc::elect_one()) {
      const int t = warp;
      constexpr uint32_t idesc_s = tc::make_idesc_bf16(128, KQ, 0, 0);  // Q (K-major / tensor memory) x K (K-major)
      constexpr uint32_t idesc_o = tc::make_idesc_bf16(128, on, 0, 1);   // P (tensor memory) x V (MN-major)
      constexpr uint32_t TILE16 = TILE_BYTES >> 4, QUART16 = (KQ * 128) >> 4;
      const uint64_t qdesc = tc::make_sw128_desc(tc::smem_u32(smem + OFF_Q)) + (uint64_t)(t * TILE16);
      const uint64_t kd0 = tc::make_sw128_desc(tc::smem_u32(smem + OFF_K));
      const uint64_t vd0 = tc::make_sw128_desc(tc::smem_u32(smem + OFF_V));
      const uint32_t tb = tmem + t * 128;
      // S = Q_t K^T for the quarter-block whose K rows start at descriptor `kd`
      auto issue_s = [&](uint64_t kd) {
#pragma unroll
        for (int k = 0; k < ksteps; ++k) {
          if (TSQ) umma_f16_ts(tb + TM6_S, tb + TM6_Q + k * 8, kd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
          else tc::umma_f16_ss(tb + TM6_S, qdesc + 2 * k, kd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        }
        tc::umma_commit(&s_full[t]);
      };
      long long tr_p = 0;
      if (TSQ) mbar_wait_lean(&q_ready[t], 0); else mbar_wait_lean(q_full, 0);
      mbar_wait_lean(&kv_full[0], 0);
      tc::tc_fence_after();
      issue_s(kd0);
      if (nq > 1) {                     // S_1 as soon as the row of S_0 is in registers
        mbar_wait_lean(&s_free[t], 0);
        tc::tc_fence_after();
        issue_s(kd0 + (uint64_t)QUART16);
      }
      // Round q: S_q+2 first (s_free(q + 1) arrives a little before p_full(q), and S is what the softmax waits for
      // next), then P V_q (which only has to complete before P_q+1 is written, a whole round later)
      uint32_t st = 0, kvph = 0;        // stage of K/V tile `blk`, phase of kv_full for this pass over the stages
      uint64_t kd = kd0, vd = vd0;      // descriptors of K/V tile `blk`
      for (int blk = 0; blk < nblk; ++blk) {
        // stage / phase / descriptors of K/V tile blk + 1
        const uint32_t st1 = (st + 1 == KV_STAGES) ? 0u : st + 1;
        const uint32_t ph1 = (st + 1 == KV_STAGES) ? (kvph ^ 1) : kvph;
        const uint64_t kd1 = kd0 + (uint64_t)(st1 * TILE16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int q = blk * 4 + j;
          if (q >= nq) break;
          long long* tli = (TRACE && tl && q >= TL_Q0 && q < TL_Q0 + TL_N) ? tl + 16 * TL_N * 8 + ((t * TL_N) + q - TL_Q0) * 4 : nullptr;
          if (TRACE && tli) tli[0] = clock64();
          if (q + 2 < nq) {
            if (j == 2) mbar_wait_lean(&kv_full[st1], ph1);      // quarter-block q + 2 opens K/V tile blk + 1
            mbar_wait_lean(&s_free[t], (j + 1) & 1);
            tc::tc_fence_after();
            issue_s(j >= 2 ? kd1 + (uint64_t)((j - 2) * QUART16) : kd + (uint64_t)((j + 2) * QUART16));
          }
          if (TRACE && tli) tli[1] = clock64();
          { ATR_T0(); mbar_wait_lean(&p_full[t], j & 1); ATR_ACC(tr_p); }
          tc::tc_fence_after();
          if (TRACE && tli) tli[2] = clock64();
          // O_t (+)= P_q V_q; A: 16 keys = 8 packed columns of P; B: 16 rows of 128 B of V
          umma_f16_ts(tb + TM6_O, tb + TM6_P, vd + (uint64_t)(j * QUART16), idesc_o, q != 0 ? 1u : 0u);
          umma_f16_ts(tb + TM6_O, tb + TM6_P + 8, vd + (uint64_t)(j * QUART16 + (16 * 128 >> 4)), idesc_o, 1u);
          tc::umma_commit(&pv_done[t]);
          if (j == 3 || q == nq - 1) tc::umma_commit(&kv_empty[st]);
          if (TRACE && tli) tli[3] = clock64();
        }
        st = st1; kvph = ph1; kd = kd1; vd = vd0 + (uint64_t)(st1 * TILE16);
      }
      if (trc && warp == 0) { trc[6] = tr_p; trc[7] = (nq + 1) / 2; }
    }
    __syncwarp();
  } else if (warp < SM_WARP0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    // ---------------------------------------------------------------- TMA loader (warp 4)
    if (warp == 4) {
      if (tc::elect_one()) {
        tc::mbar_arrive_expect_tx(q_full, ntiles * TILE_BYTES);
        for (int t = 0; t < ntiles; ++t)
          tc::tma_load_3d(smem + OFF_Q + t * TILE_BYTES, &map_qkv, q_full, cq, q0 + t * QT, b);
      }
      __syncwarp();
      uint32_t s = 0, ph = 0;
      for (int j = 0; j < nblk; ++j) {
        tc::mbar_wait(&kv_empty[s], ph ^ 1);
        if (tc::elect_one()) {
          tc::mbar_arrive_expect_tx(&kv_full[s], 2 * TILE_BYTES);
          tc::tma_load_3d(smem + OFF_K + s * TILE_BYTES, &map_qkv, &kv_full[s], ck, j * KT, b);
          tc::tma_load_3d(smem + OFF_V + s * TILE_BYTES, &map_qkv, &kv_full[s], cv, j * KT, b);
        }
        __syncwarp();
        if (++s == KV_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ---------------------------------------------------------------- softmax / output: thread <-> query row
    const int t = (warp - SM_WARP0) >> 2;
    if (t < ntiles) {
      const int qd = warp & 3;
      const int row = qd * 32 + lane;
      const uint32_t tb = tmem + ((uint32_t)(qd * 32) << 16) + t * 128;
      const uint32_t sa = tb + TM6_S, pa = tb + TM6_P, oa = tb + TM6_O;
      float m = -INFINITY;                   // integer-valued reference maximum (log2 domain)
      uint64_t l2 = pack2(0.f, 0.f);         // row sum, two partial sums
      long long tr_s = 0;
      const uint64_t sc2 = pack2(scale_log2, scale_log2);

      if (TSQ) {
        // this thread's Q row, shared memory (128-byte rows, 16-byte chunk j of row r at position j ^ (r & 7)) ->
        // tensor memory, 8 packed columns per K step of 16 channels (q arrives multiplied by the logit scale, log2
        // domain: TcAttnParams::k_one); the extra K step starts out zero (reference maximum 0)
        tc::mbar_wait(q_full, 0);
        const uint8_t* qrow = smem + OFF_Q + t * TILE_BYTES + row * 128;
#pragma unroll
        for (int k = 0; k < ON / 16; ++k) {
          uint32_t w[8];
          const uint4 c0 = *reinterpret_cast<const uint4*>(qrow + (((2 * k) ^ (row & 7)) << 4));
          const uint4 c1 = *reinterpret_cast<const uint4*>(qrow + (((2 * k + 1) ^ (row & 7)) << 4));
          w[0] = c0.x; w[1] = c0.y; w[2] = c0.z; w[3] = c0.w; w[4] = c1.x; w[5] = c1.y; w[6] = c1.z; w[7] = c1.w;
          tmem_st_32x8(tb + TM6_Q + k * 8, w);
        }
        {
          uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
          tmem_st_32x8(tb + TM6_Q + (ON / 16) * 8, z);
        }
        tmem_st_wait();
        tc::tc_fence_before();
        if (tc::elect_one()) tc::mbar_arrive(&q_ready[t]);
      }

      // the four barriers of this tile through ONE opaque base register (s_full + 0, s_free + 32, p_full + 64,
      // pv_done + 96 bytes): the compiler would otherwise re-derive every address from the warp index at every use
      uint32_t bar_t = tc::smem_u32(&s_full[t]);
      asm volatile("mov.u32 %0, %0;" : "+r"(bar_t));
      constexpr int B_SFULL = 0, B_SFREE = NTILE * 8, B_PFULL = 2 * NTILE * 8, B_PVDONE = 3 * NTILE * 8;
      // one non-blocking test (its latency overlaps the arithmetic that follows), then a spin only if it failed
      auto bar_test = [&](int off, uint32_t parity) -> uint32_t {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar_t + off), "r"(parity) : "memory");
        return ok;
      };
      auto bar_spin = [&](int off, uint32_t parity) {
        asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}"
                     :: "r"(bar_t + off), "r"(parity) : "memory");
      };
      // every lane arrives (128 per tile): one instruction, where electing a lane costs five (measured: 3 % of the kernel)
      auto bar_arrive = [&](int off) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar_t + off) : "memory");
      };

      // scaled row maximum of a quarter-block (MASK: keys beyond T set to -inf in the registers)
      auto row_max = [&](uint32_t (&r)[32], int nvalid) -> float {
        if (nvalid < KQ) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i >= nvalid) r[i] = 0xff800000u;   // -inf
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 32; i += 2)
          mx[(i >> 1) & 3] = max3(mx[(i >> 1) & 3], __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
        const float mxa = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        return TSQ ? mxa : mxa * scale_log2;      // TSQ: the logits come scaled, relative to the baked maximum
      };
      // pairs k0 .. k1 - 1 of P = 2^(S * scale - m), rounded to bf16 pairs; their sum into l2
      auto exps = [&](const uint32_t (&r)[32], uint32_t (&pk)[16], auto k0c, auto k1c, uint64_t nm2) {
        constexpr int k0 = decltype(k0c)::value, k1 = decltype(k1c)::value;
#pragma unroll
        for (int k = k0; k < k1; ++k) {
          const uint64_t v2 = pack2(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1]));
          const uint64_t x = TSQ ? v2 : fma2(v2, sc2, nm2);       // TSQ: S * scale - m straight from the tensor core
          const bool on_fma = POLY_NUM > 0 && ((k * POLY_NUM) % POLY_DEN) < POLY_NUM;
          float p0, p1;
          if (on_fma) {
            ex2_fma2(x, p0, p1);
          } else {
            float x0, x1;
            unpack2(x, x0, x1);
            p0 = ex2(x0); p1 = ex2(x1);
          }
          l2 = add2(l2, pack2(p0, p1));
          __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
          pk[k] = *reinterpret_cast<uint32_t*>(&h2);
        }
      };
      float cm;                              // SS: scaled row maximum of the next quarter-block
      float m_pend = 0.f, corr = 0.f;        // TSQ: the raised maximum, and (baked maximum of the row in registers) - m_pend
      bool raise;
      // TSQ: the next quarter-block's logits (in `r`, relative to the baked maximum == m) ask for a higher reference:
      // choose it (bf16-representable, so the tensor core subtracts it exactly), put it into this row's Q column for
      // every S issued from now on, and remember the correction of the row already in registers.
      auto bake = [&](float cmr, float mref) {
        float f = fmaxf(m, ceilf(cmr + mref));
        const uint32_t tr = __float_as_uint(f) & 0xffff0000u;          // bf16, rounded towards zero
        float g = __uint_as_float(tr);
        if (f > 0.f && g < f) g = __uint_as_float(tr + 0x10000u);       // ... and up (exact for |f| <= 256)
        m_pend = g;
        corr = mref - g;
        tmem_st_32x1(tb + TM6_Q + (ON / 16) * 8, __float_as_uint(-g) >> 16);      // channels (ON, ON + 1) = (-m, 0)
      };
      // Quarter-block q: `v` = its logits, `raise` the vote on them (both taken during q - 1).
      // The four softmax warps of a scheduler run in lockstep (same code, same data rates), so nothing but this warp's
      // own arithmetic can cover the latency of a barrier test, tcgen05.ld or tcgen05.st: each is issued early and its
      // result used late.  LAST: no quarter-block q + 1; MASK: quarter-block q + 1 may hold keys beyond T.
      auto round = [&](int q, uint32_t (&v)[32], uint32_t (&vn)[32], auto last_c, auto mask_c) {
        constexpr bool LAST = decltype(last_c)::value, MASK = decltype(mask_c)::value;
        long long* tlq = (TRACE && tl && lane == 0 && q >= TL_Q0 && q < TL_Q0 + TL_N) ? tl + (((t * 4 + qd) * TL_N) + q - TL_Q0) * 8 : nullptr;
        if (TRACE && tlq) tlq[0] = clock64();
        bool pv_seen = q == 0;          // P V_q-1 known complete
        if (raise) {
          // raise the reference maximum: O and l scale by the exact power of two 2^(m_old - m_new)
          const float m_new = TSQ ? m_pend : fmaxf(m, ceilf(cm));
          const float alpha = (m_new == m) ? 1.0f : ex2(m - m_new);     // first quarter-block: ex2(-inf) = 0
          m = m_new;
          l2 = fma2(l2, pack2(alpha, alpha), pack2(0.f, 0.f));
          if (TSQ) {                    // this row was produced against the previous baked maximum
            const uint64_t c2 = pack2(corr, corr);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float x0, x1;
              unpack2(add2(pack2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), c2), x0, x1);
              v[i] = __float_as_uint(x0); v[i + 1] = __float_as_uint(x1);
            }
          }
          if (q > 0) {
            bar_spin(B_PVDONE, (q - 1) & 1);
            pv_seen = true;
            tc::tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < on; c += 16) {
              uint32_t o[16];
              tmem_ld_32x16(oa + c, o);
              tc::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_32x16(oa + c, o);
            }
          }
        }
        const uint64_t nm2 = pack2(-m, -m);
        uint32_t pk[16];
        using I0 = std::integral_constant<int, 0>; using I6 = std::integral_constant<int, 6>; using I16 = std::integral_constant<int, 16>;
        uint32_t ok_s = 1;
        if (!LAST) ok_s = bar_test(B_SFULL, (q + 1) & 1);
        exps(v, pk, I0{}, I6{}, nm2);
        if (TRACE && tlq) tlq[1] = clock64();
        if (!LAST) {
          if (!ok_s) { ATR_T0(); bar_spin(B_SFULL, (q + 1) & 1); ATR_ACC(tr_s); }
          tc::tc_fence_after();
          tc::tmem_ld_32x32(sa, vn);
        }
        if (TRACE && tlq) tlq[2] = clock64();
        uint32_t ok_pv = 1;
        if (!pv_seen) ok_pv = bar_test(B_PVDONE, (q - 1) & 1);
        exps(v, pk, I6{}, I16{}, nm2);
        if (TRACE && tlq) tlq[3] = clock64();
        if (!ok_pv) bar_spin(B_PVDONE, (q - 1) & 1);       // P V_q-1 has read P
        if (!pv_seen) tc::tc_fence_after();
        if (TRACE && tlq) tlq[4] = clock64();
        tmem_st_32x16(pa, pk);
        raise = false;
        if (TSQ) {
          // the maximum / vote on the next row covers the latency of the tcgen05.st; S_q+2 must not be issued before
          // a raised maximum sits in the Q column, so s_free waits for the vote
          if (!LAST) {
            tc::tmem_ld_wait();
            const float cmr = row_max(vn, MASK ? T - (q + 1) * KQ : KQ);
            raise = __any_sync(0xffffffffu, cmr > RESCALE_THRESHOLD);       // warp-uniform (the TMEM accesses are warp-collective)
            if (raise) bake(cmr, m);
          }
          if (TRACE && tlq) tlq[5] = clock64();
          tmem_st_wait();
          tc::tc_fence_before();
          if (!LAST) bar_arrive(B_SFREE);
          bar_arrive(B_PFULL);
          if (TRACE && tlq) tlq[6] = clock64();
        } else {
          if (!LAST) {
            tc::tmem_ld_wait();                       // the row of S_q+1 is in registers: S_q+2 may overwrite it
            tc::tc_fence_before();
            bar_arrive(B_SFREE);
          }
          if (TRACE && tlq) tlq[5] = clock64();
          if (!LAST) cm = row_max(vn, MASK ? T - (q + 1) * KQ : KQ);     // covers the latency of the tcgen05.st
          tmem_st_wait();
          tc::tc_fence_before();
          bar_arrive(B_PFULL);                        // the warp's tcgen05.st are complete (wait::st is warp-wide)
          if (TRACE && tlq) tlq[6] = clock64();
          if (!LAST) raise = __any_sync(0xffffffffu, cm > m + RESCALE_THRESHOLD);
        }
        if (TRACE && tlq) tlq[7] = clock64();
      };

      uint32_t va[32], vb[32];
      bar_spin(B_SFULL, 0);
      tc::tc_fence_after();
      tc::tmem_ld_32x32(sa, va);
      tc::tmem_ld_wait();
      raise = true;
      if (TSQ) {
        bake(row_max(va, T), 0.f);        // S_0 was issued against a zero Q column
        tmem_st_wait();
      } else {
        cm = row_max(va, T);
      }
      tc::tc_fence_before();
      bar_arrive(B_SFREE);
      {
        using F = std::false_type; using Tr = std::true_type;
        int q = 0;
        for (; q + 3 < nq; q += 2) {       // rounds whose successor is not the last quarter-block
          round(q, va, vb, F{}, F{});
          round(q + 1, vb, va, F{}, F{});
        }
        if (nq - q == 3) {
          round(q, va, vb, F{}, F{});
          round(q + 1, vb, va, F{}, Tr{});
          round(q + 2, va, vb, Tr{}, F{});
        } else if (nq - q == 2) {
          round(q, va, vb, F{}, Tr{});
          round(q + 1, vb, va, Tr{}, F{});
        } else {
          round(q, va, vb, Tr{}, F{});
        }
      }
      if (trc && warp == SM_WARP0 && lane == 0) { trc[1] = tr_s; trc[2] = 0; trc[3] = 0; trc[4] = 0; trc[5] = 0; }
      tc::mbar_wait(&pv_done[t], (nq - 1) & 1);
      tc::tc_fence_after();
      const int qi = q0 + t * QT + row;
      __nv_bfloat16* op = out + ((long long)b * T + qi) * (heads * ch) + h * ch;
      float l0, l1;
      unpack2(l2, l0, l1);
      const float inv = 1.0f / (l0 + l1);
#pragma unroll 1
      for (int c = 0; c < on; c += 16) {
        uint32_t o[16];
        tmem_ld_32x16(oa + c, o);
        tc::tmem_ld_wait();
        if (qi < T) {
#pragma unroll
          for (int d = 0; d < 16; d += 8) {
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
  }
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, TM_COLS);
  }
  if (trc && threadIdx.x == 0) trc[0] = clock64() - t_entry;
}

long long* g_attn_trace = nullptr;
int g_attn_trace_n = 0;

}  // namespace

void tc_attn_set_trace(long long* dev_buf, int n) { g_attn_trace = dev_buf; g_attn_trace_n = n; }

struct TcAttnPlan {
  CUtensorMap map;
  TcAttnParams p;
};

int tc_attn_plan_create(const TcAttnParams& p, TcAttnPlan** out) {
  EO_REQUIRE(p.ch <= HD && p.ch % 8 == 0, EO_ERR_ARG,
             "tc_attn: head dimension %d unsupported (must be a multiple of 8, <= 64)", p.ch);
  EO_REQUIRE((p.k_one != 0) == (p.ch <= 48), EO_ERR_ARG, "tc_attn: heads of <= 48 channels need the 1.0 channel in k (and only they)");
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

template <bool TRACE>
static int attn6_launch(const TcAttnPlan* pl, int B, cudaStream_t st, long long* trace, int trace_n) {
  const TcAttnParams& p = pl->p;
  // logits = (q . k) * ch^-1/2 ; softmax evaluated with exp2
  float scale_log2 = (1.0f / sqrtf((float)p.ch)) * 1.4426950408889634f;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
  dim3 grid((unsigned)ceil_div(p.T, NTILE * QT), (unsigned)p.heads, (unsigned)B);
  static bool attr_set = false;        // per instantiation (TRACE)
  if (!attr_set) {
    EO_CHECK_CUDA(cudaFuncSetAttribute(k_attn_tc6<TRACE, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_SMEM));
    EO_CHECK_CUDA(cudaFuncSetAttribute(k_attn_tc6<TRACE, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_SMEM));
    EO_CHECK_CUDA(cudaFuncSetAttribute(k_attn_tc6<TRACE, 48>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_SMEM));
    EO_CHECK_CUDA(cudaFuncSetAttribute(k_attn_tc6<TRACE, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_SMEM));
    attr_set = true;
  }
  auto go = [&](auto kern) -> int {
    EO_CHECK_CUDA(launch_chain(kern, grid, dim3(ATTN_THREADS), ATTN_SMEM, st, pl->map, out, p.T, p.heads, p.ch, scale_log2,
                               trace, trace_n));
    return EO_OK;
  };
  // O is round16(ch) columns wide; below 64 the Q rows live in tensor memory
  if (p.ch <= 16) return go(k_attn_tc6<TRACE, 16>);
  if (p.ch <= 32) return go(k_attn_tc6<TRACE, 32>);
  if (p.ch <= 48) return go(k_attn_tc6<TRACE, 48>);
  return go(k_attn_tc6<TRACE, 64>);
}

int tc_attn_launch(const TcAttnPlan* pl, int B, cudaStream_t st) {
#ifdef EO_DEVTOOLS
  if (g_attn_trace) return attn6_launch<true>(pl, B, st, g_attn_trace, g_attn_trace_n);
#endif
  return attn6_launch<false>(pl, B, st, nullptr, 0);
}

}  // namespace eo
