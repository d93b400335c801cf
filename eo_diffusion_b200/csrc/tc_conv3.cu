// Persistent implicit-GEMM convolution on tcgen05 (third generation of tc_conv.cu's kernel).
//
// Same GEMM view and fusions as tc_conv.cu (reference backbones/unet_openai.py: ResBlock :316,
// :342,:353,:382,:385; AttentionBlock :412,:422,:433; Downsample :262; Upsample :227; th.cat
// :773), restructured after the per-CTA timeline measured on B200
// (profiles/r01_conv_tc_pair_cta_timeline.txt: the main loop ran at tensor-pipe rate but 55 % of
// every CTA's life was set-up, first-load latency, a store-bound epilogue and tear-down):
//
//   * persistent: one CTA per SM, clusters of two (cta_group::2, 256 x BN tile per pair) walk the
//     tile list; barriers, TMEM and descriptors are set up once per SM, not once per tile;
//   * the accumulator is double-buffered in TMEM (2 x 256 columns), so the epilogue of tile i
//     overlaps the main loop of tile i+1 by construction;
//   * epilogue through shared memory: TMEM -> registers -> (+bias, +timestep row, +residual)
//     -> bf16 -> 128B-swizzled staging tile -> ONE TMA store per 64 channels (the old per-row
//     16-byte stores cost a wavefront each and bounded the epilogue); the residual tile arrives
//     by TMA as well, prefetched by its own producer warp;
//   * halo patches: for a 3x3 segment the producer loads ONE (16+2) x (8+2) pixel patch per 64
//     channels and the nine taps are nine shared-memory descriptors into it (start address shifted
//     by whole 128-byte pixel rows, stride between 8-pixel groups = the patch pitch; the 128B
//     swizzle is a function of the absolute address, tools/probe_shifted_desc.cu), cutting the
//     activation traffic L2 -> SM from 9 x 16 KB to 22.5 KB per 64 channels.
//
// Warp roles (352 threads): 0 = operand producer (TMA), 1 = TMEM owner + MMA issuer, 2..9 =
// epilogue (two sets of four, TMEM lane quadrant warp % 4), 10 = residual producer.
//
// Roofline: tensor pipe.  Algorithmic FLOPs per launch = 2 * M * Cout * Ktot.
#include "kernels.h"
#include "tc_common.cuh"
#include "tc_conv_plan.h"
#include <cstdlib>
#include <vector>

namespace eo {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int PATCH_W = 10, PATCH_H = 18;                  // halo patch of an 8 x 16 pixel tile
constexpr int PATCH_BYTES = PATCH_W * PATCH_H * 128;       // 23040
constexpr int PLAIN_BYTES = BM * 128;                      // 16384
constexpr int A_STAGE = 23552;                             // 23 KB: PATCH_BYTES rounded up to 1 KB
constexpr int SA = 3;                                      // operand-A stages
constexpr int STG_BYTES = BM * 128;                        // one 128-row x 64-channel bf16 tile
constexpr int MAX_ENT = 176;
constexpr int MAX_SB = 12;
constexpr int TMEM_COLS = 512, ACC_STRIDE = 256;
constexpr int SMEM_LIMIT = 232448;                         // 227 KB per CTA
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 32 * (3 + NUM_EPI_WARPS);   // producer, MMA, 8 epilogue, residual producer

// fixed part of the shared-memory layout (offsets from the 1 KB-aligned base); the B ring follows
struct Smem {
  static constexpr int A_OFF = 0;
  static constexpr int STG_OFF = A_OFF + SA * A_STAGE;                 // 2 staging tiles
  static constexpr int TAB_OFF = STG_OFF + 2 * STG_BYTES;
  static constexpr int STAT_OFF = TAB_OFF + MAX_ENT * (int)sizeof(KEnt3);   // [4 warps][256][2] floats
  static constexpr int BAR_OFF = STAT_OFF + 4 * 256 * 2 * 4;
  // a_full[SA] a_empty[SA] b_full[MAX_SB] b_empty[MAX_SB] tmem_full[2] tmem_empty[2] res_full[2] res_empty[2]
  static constexpr int NBAR = 2 * SA + 2 * MAX_SB + 8;
  static constexpr int VAR_OFF = (BAR_OFF + NBAR * 8 + 16 + 1023) & ~1023;  // residual tiles (optional), then B ring
};

struct Tile { int w0, h0, n0, nbase; };

__device__ __forceinline__ Tile decode_tile(int w, int n_ntiles, uint32_t rank, const Geom3& g, int BN) {
  const int mp = w / n_ntiles, nt = w - mp * n_ntiles;
  const int mt = 2 * mp + (int)rank;
  Tile t;
  const int tw = mt % g.tiles_w;
  const int th = (mt / g.tiles_w) % g.tiles_h;
  t.w0 = tw * g.bw; t.h0 = th * g.bh; t.n0 = (mt / (g.tiles_w * g.tiles_h)) * g.bn;
  t.nbase = nt * BN;
  return t;
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      :: "l"(reinterpret_cast<uint64_t>(m)), "r"(tc::smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

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

// development aid (eo_debug_conv_trace, TRACE instantiation only): per-CTA counters, slot = 0 lifetime,
// 1 MMA waits on operands, 2 MMA waits on a free accumulator, 3 epilogue waits on the accumulator,
// 4 epilogue busy, 5 producer waits on free stages, 6 tiles, 7 SM id (clock64 ticks)
__device__ __forceinline__ void trace_put(const Epi3& ep, int slot, long long v) {
  if (ep.trace && (int)blockIdx.x < ep.trace_n) ep.trace[(long long)blockIdx.x * 8 + slot] = v;
}
#define TRACE_T0() const long long _t0 = TRACE ? clock64() : 0
#define TRACE_ACC(var) do { if (TRACE) (var) += clock64() - _t0; } while (0)

// ring position: stage index + phase parity, advanced without a division
struct Ring {
  uint32_t i = 0, ph = 0;
  __device__ __forceinline__ void next(uint32_t n) { if (++i == n) { i = 0; ph ^= 1; } }
};

template <bool TRACE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_conv_tc3(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
           const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
           const __grid_constant__ CUtensorMap mapOut, const __grid_constant__ CUtensorMap mapRes,
           const KEnt3* __restrict__ ents, int nent, Geom3 g, int B, int BN, int SB, int n_work, int n_ntiles,
           Epi3 ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (tc::smem_u32(smem_raw) & 1023)) & 1023);
  KEnt3* tab = reinterpret_cast<KEnt3*>(smem + Smem::TAB_OFF);
  float* sstat = reinterpret_cast<float*>(smem + Smem::STAT_OFF);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + Smem::BAR_OFF);
  uint64_t* a_empty = a_full + SA;
  uint64_t* b_full = a_empty + SA;
  uint64_t* b_empty = b_full + MAX_SB;
  uint64_t* tmem_full = b_empty + MAX_SB;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_full = tmem_empty + 2;
  uint64_t* res_empty = res_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_empty + 2);
  const bool has_res = ep.has_res != 0;
  uint8_t* res_sm = smem + Smem::VAR_OFF;
  const int b_bytes = (BN / 2) * 128;
  uint8_t* b_sm = res_sm + (has_res ? 2 * STG_BYTES : 0);
  uint8_t* stg_sm = smem + Smem::STG_OFF;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const long long t_entry = TRACE ? clock64() : 0;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&mapA0);
    tc::tma_prefetch_desc(&mapB);
    tc::tma_prefetch_desc(&mapOut);
    for (int s = 0; s < SA; ++s) { tc::mbar_init(&a_full[s], 1); tc::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < MAX_SB; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&tmem_full[s], 1);
      tc::mbar_init(&tmem_empty[s], 2 * NUM_EPI_WARPS);   // every epilogue warp of both CTAs (leader's copy is used)
      tc::mbar_init(&res_full[s], 1);
      tc::mbar_init(&res_empty[s], 4);                    // the four warps of one epilogue set
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) { tc::tmem_alloc2(tmem_ptr, TMEM_COLS); tc::tmem_relinquish2(); }
  for (int i = threadIdx.x; i < nent; i += blockDim.x) tab[i] = ents[i];
  tc::tc_fence_before();
  tc::cluster_sync_all();                    // peer barriers are initialised past here
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // Producer and MMA warps: the WHOLE warp walks the loops (warp-uniform control flow keeps addresses,
  // coordinates and descriptors in uniform registers) and one elected lane issues.  Under
  // `if (lane == 0)` ptxas wraps every TMA / MMA instruction in a read-lane loop and the issuing
  // thread, not the tensor pipe, becomes the limiter (measured: 650 clk per 64-deep K block).
  if (warp == 0) {
    // ------------------------------------------------------------------ operand producer
    Ring ra, rb;
    long long tr_wait = 0;
    const uint32_t a_full_l = tc::mapa_u32(tc::smem_u32(&a_full[0]), 0);   // the leader's barriers
    const uint32_t b_full_l = tc::mapa_u32(tc::smem_u32(&b_full[0]), 0);
    const int b_row = (int)rank * (BN / 2);
    for (int w = cid; w < n_work; w += ncl) {
      const Tile t = decode_tile(w, n_ntiles, rank, g, BN);
      for (int e = 0; e < nent; ++e) {
        const KEnt3 en = tab[e];
        { TRACE_T0(); tc::mbar_wait(&a_empty[ra.i], ra.ph ^ 1); TRACE_ACC(tr_wait); }
        const CUtensorMap* ma = en.seg == 0 ? &mapA0 : (en.seg == 1 ? &mapA1 : &mapA2);
        if (tc::elect_one()) {
          // One arrival per phase: the leader's producer, which posts the byte count of BOTH CTAs'
          // loads; the peer's TMA only completes transactions on the leader's barrier.
          if (rank == 0) tc::mbar_arrive_expect_tx(&a_full[ra.i], 2 * (en.patch ? PATCH_BYTES : PLAIN_BYTES));
          tc::tma2_load_4d(smem + Smem::A_OFF + ra.i * A_STAGE, ma, a_full_l + ra.i * 8, en.c0,
                           t.w0 + (en.patch ? -1 : en.dw), t.h0 + (en.patch ? -1 : en.dh), t.n0 + en.dn);
        }
        __syncwarp();
        ra.next(SA);
        const int nb = en.patch ? 9 : 1;
        for (int j = 0; j < nb; ++j) {
          { TRACE_T0(); tc::mbar_wait(&b_empty[rb.i], rb.ph ^ 1); TRACE_ACC(tr_wait); }
          if (tc::elect_one()) {
            if (rank == 0) tc::mbar_arrive_expect_tx(&b_full[rb.i], 2 * b_bytes);
            tc::tma2_load_2d(b_sm + rb.i * b_bytes, &mapB, b_full_l + rb.i * 8, en.kofs + j * BK, t.nbase + b_row);
          }
          __syncwarp();
          rb.next((uint32_t)SB);
        }
      }
    }
    if (TRACE && lane == 0) trace_put(ep, 5, tr_wait);
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    if (rank == 0) {
      const uint32_t idesc = tc::make_idesc_bf16(2 * BM, BN, 0, 0);
      const uint64_t bdesc0 = tc::make_sw128_desc(tc::smem_u32(b_sm));
      const uint32_t b_step = (uint32_t)b_bytes >> 4;
      Ring ra, rb;
      uint32_t it = 0;
      long long tr_ops = 0, tr_acc = 0;
      for (int w = cid; w < n_work; w += ncl, ++it) {
        const uint32_t ab = it & 1;
        { TRACE_T0(); tc::mbar_wait_cluster(&tmem_empty[ab], ((it >> 1) & 1) ^ 1); TRACE_ACC(tr_acc); }   // both CTAs' epilogues drained this buffer
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * ACC_STRIDE;
        uint32_t acc = 0;
        for (int e = 0; e < nent; ++e) {
          const int patch = tab[e].patch;
          { TRACE_T0(); tc::mbar_wait(&a_full[ra.i], ra.ph); TRACE_ACC(tr_ops); }
          tc::tc_fence_after();
          const uint32_t a_base = tc::smem_u32(smem + Smem::A_OFF + ra.i * A_STAGE);
          const bool last_e = e == nent - 1;
          // one weight tile: wait for it, issue the 4 K=16 MMAs of this 64-deep K block, release it
          auto kblock = [&](uint64_t adesc, bool last_of_a) {
            { TRACE_T0(); tc::mbar_wait(&b_full[rb.i], rb.ph); TRACE_ACC(tr_ops); }
            tc::tc_fence_after();
            const uint64_t bdesc = bdesc0 + (uint64_t)(rb.i * b_step);
            if (tc::elect_one()) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                tc::umma2_f16_ss(d_tmem, tc::desc_advance(adesc, k * 32), tc::desc_advance(bdesc, k * 32), idesc,
                                 k ? 1u : acc);
              tc::umma2_commit_mc(&b_empty[rb.i], 3);                   // frees the weight stage in both CTAs
              if (last_of_a) tc::umma2_commit_mc(&a_empty[ra.i], 3);    // ... and the activation stage
              if (last_of_a && last_e) tc::umma2_commit_mc(&tmem_full[ab], 3);   // accumulator complete
            }
            __syncwarp();
            acc = 1;
            rb.next((uint32_t)SB);
          };
          if (patch) {
            // tap (kh, kw) of a patch starts kh patch rows + kw pixels into it
            const uint64_t ad0 = tc::make_sw128_desc_sbo(a_base, PATCH_W * 128);
#pragma unroll
            for (int j = 0; j < 9; ++j)
              kblock(ad0 + (uint64_t)(((j / 3) * PATCH_W + (j % 3)) * 8), j == 8);
          } else {
            kblock(tc::make_sw128_desc(a_base), true);
          }
          ra.next(SA);
        }
      }
      if (TRACE && lane == 0) { trace_put(ep, 1, tr_ops); trace_put(ep, 2, tr_acc); trace_put(ep, 6, it); }
    }
  } else if (warp == 2 + NUM_EPI_WARPS) {
    // ------------------------------------------------------------------ residual producer
    if (has_res) {
      if (lane == 0) tc::tma_prefetch_desc(&mapRes);
      uint32_t uses[2] = {0, 0};
      const int nchunks = BN / 64;
      for (int w = cid; w < n_work; w += ncl) {
        const Tile t = decode_tile(w, n_ntiles, rank, g, BN);
        for (int c = 0; c < nchunks; ++c) {
          const uint32_t rbuf = c & 1;                  // chunk c belongs to epilogue set c & 1
          tc::mbar_wait(&res_empty[rbuf], (uses[rbuf] & 1) ^ 1);
          ++uses[rbuf];
          if (tc::elect_one()) {
            tc::mbar_arrive_expect_tx(&res_full[rbuf], STG_BYTES);
            tc::tma_load_4d(res_sm + rbuf * STG_BYTES, &mapRes, &res_full[rbuf], t.nbase + c * 64, t.w0, t.h0, t.n0);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    // Two sets of four warps (one warp per TMEM lane quadrant in each); set s takes the 64-channel
    // chunks c = s, s+2, ... of every tile.  A warp owns 32 accumulator rows end to end: TMEM ->
    // registers -> its 4 KB slice of the staging tile -> its own TMA store (a 32-pixel sub-box), so
    // the only cross-warp synchronisation is the per-tile combine of the GroupNorm sums.
    const int ew = warp - 2;                       // 0..7
    const int set = ew >> 2;
    const int q = warp & 3;                        // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;
    const int et = ew * 32 + lane;                 // 0..255
    const int nn = row / (g.bw * g.bh);
    // the warp's 32 rows as a sub-box of the tile
    const int row0 = q * 32;
    const int sub_w = row0 % g.bw, sub_h = (row0 / g.bw) % g.bh, sub_n = row0 / (g.bw * g.bh);
    const int nchunks = BN / 64;
    const bool do_stats = ep.stats != nullptr;
    const bool single_image = g.bn == 1;
    const int Cout = ep.Cout;
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t row_off = (uint32_t)row * 128u;
    uint8_t* stg = stg_sm + set * STG_BYTES;       // this set's staging tile; the warp writes rows q*32..q*32+31
    const uint8_t* rs = res_sm + set * STG_BYTES;
    const uint32_t tmem_empty_leader = tc::mapa_u32(tc::smem_u32(&tmem_empty[0]), 0);
    uint32_t it = 0, res_uses = 0;
    long long tr_wait = 0, tr_busy = 0;
    for (int w = cid; w < n_work; w += ncl, ++it) {
      const Tile t = decode_tile(w, n_ntiles, rank, g, BN);
      const int n_img = t.n0 + nn;
      const bool valid = n_img < B;
      const bool warp_valid = __shfl_sync(0xffffffffu, valid ? 1 : 0, 0) != 0;
      const float* bnc = (ep.bias_nc && valid) ? ep.bias_nc + (long long)n_img * ep.ld_bias_nc : nullptr;
      const uint32_t ab = it & 1;
      { TRACE_T0(); tc::mbar_wait(&tmem_full[ab], (it >> 1) & 1); TRACE_ACC(tr_wait); }
      tc::tc_fence_after();
      const long long t_busy0 = TRACE ? clock64() : 0;
      if (set >= nchunks) {                        // nothing to read for this warp: release the accumulator at once
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive_cluster_relaxed(tmem_empty_leader + ab * 8);
      }
#pragma unroll 1
      for (int c = set; c < nchunks; c += 2) {
        const uint32_t taddr = tmem_base + ab * ACC_STRIDE + (uint32_t)(c * 64) + ((uint32_t)(q * 32) << 16);
        uint32_t v[2][32];
        tc::tmem_ld_32x32(taddr, v[0]);
        tc::tmem_ld_32x32(taddr + 32, v[1]);
        tc::tmem_ld_wait();
        if (c + 2 >= nchunks) {                    // this warp's share of the accumulator now lives in registers
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive_cluster_relaxed(tmem_empty_leader + ab * 8);
        }
        if (has_res) { tc::mbar_wait(&res_full[set], res_uses & 1); ++res_uses; }
        // the TMA store this warp issued from its staging rows one chunk ago has read them out
        if (lane == 0) bulk_wait_read0();      // (bulk groups belong to the issuing thread: always lane 0)
        __syncwarp();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int n = t.nbase + c * 64 + half * 32;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[half][j]);
          if (ep.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n + j));
              f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
            }
          }
          if (bnc) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bnc + n + j));
              f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
            }
          }
          if (has_res) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint4 r = *reinterpret_cast<const uint4*>(rs + row_off + ((((uint32_t)(half * 4 + k)) ^ swz) << 4));
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
              for (int e2 = 0; e2 < 4; ++e2) {
                const float2 tt = __bfloat1622float2(h2[e2]);
                f[k * 8 + e2 * 2] += tt.x; f[k * 8 + e2 * 2 + 1] += tt.y;
              }
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint4 o;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) h2[e2] = __floats2bfloat162_rn(f[k * 8 + e2 * 2], f[k * 8 + e2 * 2 + 1]);
            *reinterpret_cast<uint4*>(stg + row_off + ((((uint32_t)(half * 4 + k)) ^ swz) << 4)) = o;
          }
          if (do_stats && warp_valid && n < Cout) {
            // per-channel sum and sum of squares of this warp's 32 pixel rows (GroupNorm statistics of
            // the tensor being written, from the fp32 values; reduced per group by k_gn_finalize_ch).
            // Deterministic: fixed-order fp32 partial sums; only the cross-tile accumulation is atomic,
            // and that one is in double.
            float sq[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) sq[j] = f[j] * f[j];
            const float cs = warp_column_sums(f, lane);
            const float cq = warp_column_sums(sq, lane);
            const int col = c * 64 + half * 32 + lane;
            if (single_image) {
              sstat[(q * 256 + col) * 2] = cs;
              sstat[(q * 256 + col) * 2 + 1] = cq;
            } else {
              double* dst = ep.stats + ((long long)n_img * Cout + n + lane) * 2;
              atomicAdd(dst, (double)cs);
              atomicAdd(dst + 1, (double)cq);
            }
          }
        }
        if (has_res) {
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&res_empty[set]);
        }
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&mapOut, stg + row0 * 128, t.nbase + c * 64, t.w0 + sub_w, t.h0 + sub_h, t.n0 + sub_n);
          bulk_commit();
        }
        __syncwarp();
      }
      if (do_stats && single_image) {
        asm volatile("bar.sync 1, 256;" ::: "memory");      // all eight epilogue warps
        if (t.n0 < B) {
          for (int i = et; i < BN * 2; i += 32 * NUM_EPI_WARPS) {
            const int cch = t.nbase + (i >> 1);
            if (cch < Cout) {
              const float tsum = (sstat[i] + sstat[512 + i]) + (sstat[1024 + i] + sstat[1536 + i]);
              atomicAdd(ep.stats + ((long long)t.n0 * Cout + cch) * 2 + (i & 1), (double)tsum);
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");      // sstat is rewritten by the next tile
      }
      if (TRACE) tr_busy += clock64() - t_busy0;
    }
    if (lane == 0) bulk_wait0();
    __syncwarp();
    if (TRACE && et == 0) { trace_put(ep, 3, tr_wait); trace_put(ep, 4, tr_busy); }
    tc::tc_fence_before();
  }
  __syncwarp();
  tc::cluster_sync_all();     // nobody leaves while the peer still uses its smem/TMEM
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc2(tmem_base, TMEM_COLS);
  }
  if (TRACE && threadIdx.x == 0) {
    unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    trace_put(ep, 0, clock64() - t_entry); trace_put(ep, 7, sm);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static bool env_flag(const char* name, bool dflt) {
  const char* e = std::getenv(name);
  if (!e || !e[0]) return dflt;
  return e[0] != '0';
}

bool tc_conv3_enabled() {
  static int v = -1;
  if (v < 0) v = env_flag("EO_CONV_V2", false) ? 0 : 1;
  return v != 0;
}

bool tc_conv_patch_supported(int H, int W) {
  static int v = -1;
  if (v < 0) v = env_flag("EO_CONV_PATCH", true) ? 1 : 0;
  return tc_conv3_enabled() && v != 0 && H % 16 == 0 && W % 8 == 0;
}

static int floor_pow2_(int v) { int r = 1; while (r * 2 <= v) r *= 2; return r; }

int tc_conv3_plan_fill(const TcConvParams& p, TcConvPlan* pl) {
  Geom3 g;
  g.H = p.H; g.W = p.W;
  bool any_patch = false;
  for (int s = 0; s < p.nseg; ++s) any_patch |= p.seg[s].patch != 0;
  if (any_patch) {
    EO_REQUIRE(p.H % 16 == 0 && p.W % 8 == 0, EO_ERR_ARG, "tc_conv3: halo patches need H %% 16 == 0 and W %% 8 == 0 (%dx%d)", p.H, p.W);
    g.bw = 8; g.bh = 16; g.bn = 1;
  } else {
    g.bw = floor_pow2_(p.W < 16 ? p.W : 16);
    g.bh = floor_pow2_(p.H < BM / g.bw ? p.H : BM / g.bw);
    g.bn = BM / (g.bw * g.bh);
  }
  EO_REQUIRE(p.W % g.bw == 0 && p.H % g.bh == 0, EO_ERR_ARG, "tc_conv3: feature map %dx%d is not tileable by %dx%d boxes",
             p.H, p.W, g.bh, g.bw);
  EO_REQUIRE(!(p.stats && g.bw * g.bh < 32), EO_ERR_ARG,
             "tc_conv3: fused GroupNorm statistics need at least 32 pixels per image (%dx%d)", p.H, p.W);
  EO_REQUIRE(!p.out_f32 && !p.res_f32, EO_ERR_ARG, "tc_conv3: bf16 outputs and residuals only");
  g.tiles_w = p.W / g.bw; g.tiles_h = p.H / g.bh;
  pl->g3 = g;
  int BN = 64;
  for (int cand : {256, 192, 128, 64}) if (p.Cout % cand == 0) { BN = cand; break; }
  pl->bn_tile = BN;
  std::vector<KEnt3> tab;
  int kofs = 0;
  for (int s = 0; s < p.nseg; ++s) {
    const TcConvSeg& sg = p.seg[s];
    EO_REQUIRE(sg.C % BK == 0, EO_ERR_ARG, "tc_conv3: segment channels %d must be a multiple of 64", sg.C);
    uint64_t dims[4] = {(uint64_t)sg.C, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)sg.Bt};
    uint64_t str[3] = {(uint64_t)sg.C * 2, (uint64_t)p.W * sg.C * 2, (uint64_t)p.H * p.W * sg.C * 2};
    uint32_t box[4] = {(uint32_t)BK, (uint32_t)g.bw, (uint32_t)g.bh, (uint32_t)g.bn};
    if (sg.patch) {
      EO_REQUIRE(sg.ntaps == 9, EO_ERR_ARG, "tc_conv3: a patch segment has nine taps");
      for (int t = 0; t < 9; ++t)
        EO_REQUIRE(sg.dh[t] == t / 3 - 1 && sg.dw[t] == t % 3 - 1 && sg.dn[t] == 0, EO_ERR_ARG,
                   "tc_conv3: a patch segment is a plain 3x3 window");
      box[1] = PATCH_W; box[2] = PATCH_H; box[3] = 1;
    }
    int rc = encode_tmap_bf16(&pl->mapA[s], sg.ptr, 4, dims, str, box);
    if (rc != EO_OK) return rc;
    if (sg.patch) {
      // K order of a patch segment: (64-channel block, tap, channel)
      for (int c0 = 0; c0 < sg.C; c0 += BK) {
        KEnt3 e{}; e.seg = s; e.c0 = c0; e.patch = 1; e.kofs = kofs;
        tab.push_back(e);
        kofs += 9 * BK;
      }
    } else {
      for (int t = 0; t < sg.ntaps; ++t)
        for (int c0 = 0; c0 < sg.C; c0 += BK) {
          KEnt3 e{}; e.seg = s; e.c0 = c0; e.dh = sg.dh[t]; e.dw = sg.dw[t]; e.dn = sg.dn[t]; e.kofs = kofs;
          tab.push_back(e);
          kofs += BK;
        }
    }
  }
  for (int s = p.nseg; s < 3; ++s) pl->mapA[s] = pl->mapA[0];
  EO_REQUIRE(kofs == p.Ktot && (int)tab.size() <= MAX_ENT, EO_ERR_ARG,
             "tc_conv3: K extent %d inconsistent with Ktot %d (or more than %d operand loads: %d)", kofs, p.Ktot, MAX_ENT,
             (int)tab.size());
  pl->nkb = (int)tab.size();
  {
    uint64_t dims[2] = {(uint64_t)p.Ktot, (uint64_t)p.Cout};
    uint64_t str[1] = {(uint64_t)p.Ktot * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)(BN / 2)};
    int rc = encode_tmap_bf16(&pl->mapB, p.Wp, 2, dims, str, box);
    if (rc != EO_OK) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)p.Cout, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B};
    uint64_t str[3] = {(uint64_t)p.Cout * 2, (uint64_t)p.W * p.Cout * 2, (uint64_t)p.H * p.W * p.Cout * 2};
    // output: one store per epilogue warp = the 32 pixels of its TMEM lane quadrant
    const uint32_t sw = (uint32_t)(g.bw < 32 ? g.bw : 32);
    const uint32_t sh = (uint32_t)(g.bh < (int)(32 / sw) ? g.bh : (int)(32 / sw));
    uint32_t sbox[4] = {(uint32_t)BK, sw, sh, 32 / (sw * sh)};
    int rc = encode_tmap_bf16(&pl->mapOut, p.out, 4, dims, str, sbox);
    if (rc != EO_OK) return rc;
    pl->mapRes = pl->mapOut;
    if (p.residual) {
      uint32_t box[4] = {(uint32_t)BK, (uint32_t)g.bw, (uint32_t)g.bh, (uint32_t)g.bn};
      rc = encode_tmap_bf16(&pl->mapRes, p.residual, 4, dims, str, box);
      if (rc != EO_OK) return rc;
    }
  }
  cudaError_t e = cudaMalloc(&pl->d_kblks, tab.size() * sizeof(KEnt3));
  if (e == cudaSuccess) e = cudaMemcpy(pl->d_kblks, tab.data(), tab.size() * sizeof(KEnt3), cudaMemcpyHostToDevice);
  EO_REQUIRE(e == cudaSuccess, EO_ERR_CUDA, "tc_conv3: operand table upload failed: %s", cudaGetErrorString(e));
  pl->v3 = true;
  return EO_OK;
}

static long long* g_trace3 = nullptr;
static int g_trace3_n = 0;
void tc_conv3_set_trace(long long* dev_buf, int n) { g_trace3 = dev_buf; g_trace3_n = n; }

int tc_conv3_launch(const TcConvPlan* pl, int B, cudaStream_t st) {
  static bool attr_set = false;
  const TcConvParams& p = pl->p;
  const Geom3& g = pl->g3;
  const int BN = pl->bn_tile;
  const bool has_res = p.residual != nullptr;
  const int b_bytes = (BN / 2) * 128;
  const int fixed = Smem::VAR_OFF + (has_res ? 2 * STG_BYTES : 0);
  int SB = (SMEM_LIMIT - 1024 - fixed) / b_bytes;
  if (SB > MAX_SB) SB = MAX_SB;
  EO_REQUIRE(SB >= 2, EO_ERR_STATE, "tc_conv3: shared memory budget leaves %d weight stages", SB);
  const int dyn = fixed + SB * b_bytes + 1024;
  if (!attr_set) {
    EO_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tc3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    EO_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tc3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    attr_set = true;
  }
  const int tiles_n = (int)ceil_div(B, g.bn);
  const int mtiles = g.tiles_w * g.tiles_h * tiles_n;
  const int mpairs = (mtiles + 1) / 2;            // an odd last tile gets an all-masked partner
  const int n_ntiles = p.Cout / BN;
  const int n_work = mpairs * n_ntiles;
  int ncl = num_sms() / 2;
  if (ncl > n_work) ncl = n_work;
  Epi3 ep{};
  ep.bias = p.bias; ep.bias_nc = p.bias_nc; ep.ld_bias_nc = p.ld_bias_nc; ep.stats = p.stats; ep.Cout = p.Cout;
  ep.has_res = has_res ? 1 : 0;
  ep.trace = g_trace3; ep.trace_n = g_trace3_n;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * ncl));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = (size_t)dyn;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  auto kern = g_trace3 ? k_conv_tc3<true> : k_conv_tc3<false>;
  EO_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, pl->mapA[0], pl->mapA[1], pl->mapA[2], pl->mapB, pl->mapOut,
                                   pl->mapRes, (const KEnt3*)pl->d_kblks, pl->nkb, g, B, BN, SB, n_work, n_ntiles, ep));
  return EO_OK;
}

}  // namespace eo
