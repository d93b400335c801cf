// Persistent implicit-GEMM convolution on tcgen05 (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
// GEMM view and fusions as stated in tc_conv.cu (reference backbones/unet_openai.py: ResBlock :316,
// :342,:353,:382,:385; AttentionBlock :412,:422,:433; Downsample :262; Upsample :227; th.cat
// :773).  Structure, after the per-CTA timeline of a one-tile-per-CTA first version measured on B200
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
//   * launches made of plain 128-pixel tiles only (1x1 and stride-2 convolutions, 3x3 windows over maps that take no
//     halo patch) feed four MMAs per operand load, fewer clocks of tensor-pipe work than the MMA warp's wait / fence /
//     issue / commit chain: they pack 16 KB operand stages and hand the MMA warp groups of up to four K blocks per
//     barrier round trip (GRP; EO_CONV_GROUP_MAX below).
//
//   * GroupNorm + SiLU of the operand folded in: for such a segment the halo patch is not fetched by
//     TMA; four transform warps load it from global memory into registers, apply
//     act(x * scale[n,c] + shift[n,c]) and write the 128B-swizzled operand stage themselves (zeros for
//     halo pixels outside the image: the convolution pads the NORMALISED activation).  One pass through
//     the shared-memory port, like the TMA write it replaces -- a first version that let TMA land the
//     raw patch and rewrote it in place tripled the patch's shared-memory traffic against a port the
//     MMA operand reads already saturate, and lost more than the separate bandwidth pass
//     (k_gn_apply: 10.7 ms of an 85 ms step) it was meant to remove.  A patch is transformed once for
//     its nine taps.
//
// Warp roles (640 threads, registers redistributed with setmaxnreg): 0 = operand producer (TMA), 1 = TMEM owner + MMA issuer, 2 = residual
// producer, 4..11 = epilogue (two sets of four, TMEM lane quadrant warp % 4), 12..19 = transform.
//
// Roofline: tensor pipe.  Algorithmic FLOPs per launch = 2 * M * Cout * Ktot.
#include "kernels.h"
#include "tc_common.cuh"
#include "tc_conv_plan.h"
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <vector>

namespace eo {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int PATCH_W = 10, PATCH_H = 18;                  // halo patch of an 8 x 16 pixel tile
constexpr int PATCH_BYTES = PATCH_W * PATCH_H * 128;       // 23040
constexpr int PLAIN_BYTES = BM * 128;                      // 16384
constexpr int A_STAGE = 23552;                             // 23 KB: PATCH_BYTES rounded up to 1 KB
constexpr int SA_MAX = 8;                                  // operand-A stages: SAR filled by TMA + SAG by the transform warps
// A launch whose operand loads are all plain 128-pixel tiles (1x1 / stride-2 convolutions, 3x3 windows over maps that
// take no halo patch) feeds only four K = 16 MMAs per 16 KB operand load: its K loop runs at the latency of a TMA load
// divided by the loads in flight.  Such launches pack their stages at PLAIN_BYTES and split the shared memory evenly
// between operand and weight stages (up to SA_MAX deep) instead of three 23 KB stages next to twelve weight stages.
#ifndef EO_CONV_DEEP_RING
#define EO_CONV_DEEP_RING 1
#endif
// ... and, for tiles narrower than 256 channels, hand the MMA warp GROUPS of up to EO_CONV_GROUP_MAX such loads per
// barrier round trip (one operand stage = the group's tiles side by side, one weight stage = its weight tiles): the
// wait / fence / issue / commit chain around the four MMAs of a single 64-deep K block is longer than the MMAs themselves
#ifndef EO_CONV_GROUP_MAX
#define EO_CONV_GROUP_MAX 4
#endif
#ifndef EO_CONV_GROUP_BNMAX
#define EO_CONV_GROUP_BNMAX 192      // widest channel tile that is grouped (256: measured neutral to slightly slower)
#endif
#ifndef EO_CONV_GROUP_STAGES
#define EO_CONV_GROUP_STAGES 2       // stage pairs a group size must leave room for
#endif
constexpr int STG_BYTES = BM * 128;                        // one 128-row x 64-channel bf16 tile
constexpr int MAX_ENT = 176;
constexpr int MAX_XENT = 32;                               // GroupNorm-folded patch loads per tile (<= 2048 channels)
constexpr int MAX_SB = 12;
constexpr int TMEM_COLS = 512, ACC_STRIDE = 256;
constexpr int SMEM_LIMIT = 232448;                         // 227 KB per CTA
constexpr int NUM_EPI_WARPS = 8;
constexpr int EPI_WARP0 = 4, XF_WARP0 = 12, XF_WARPS = 8;   // first epilogue warp, first transform warp
constexpr int NUM_THREADS = 32 * (XF_WARP0 + XF_WARPS);     // producer, MMA, residual producer, (spare), 8 epilogue, 8 transform
constexpr int XF_PIX = (PATCH_W * PATCH_H + 4 * XF_WARPS - 1) / (4 * XF_WARPS);   // patch pixels per transform thread (6)

// shared-memory layout after the nA operand-A stages (offsets from nA * A_STAGE past the 1 KB-aligned
// base); the B ring follows
struct Smem {
  static constexpr int STG_OFF = 0;                                    // 2 staging tiles
  static constexpr int TAB_OFF = STG_OFF + 2 * STG_BYTES;
  static constexpr int XTAB_OFF = TAB_OFF + MAX_ENT * (int)sizeof(KEnt3);   // the GroupNorm-folded loads, resolved
  static constexpr int STAT_OFF = XTAB_OFF + MAX_XENT * (int)sizeof(XEnt3);  // [4 warps][256][2] floats
  static constexpr int WB_OFF = STAT_OFF + 4 * 256 * 2 * 4;            // [8 epilogue warps][2 chunks][64] floats: bias + timestep row
  static constexpr int BAR_OFF = WB_OFF + 8 * 128 * 4;
  // r_full r_empty g_ready g_empty [SA_MAX each] b_full[MAX_SB] b_empty[MAX_SB] tmem_full[2] tmem_empty[2]
  // res_full[2] res_empty[2]
  static constexpr int NBAR = 4 * SA_MAX + 2 * MAX_SB + 8;
  static constexpr int VAR_OFF = (BAR_OFF + NBAR * 8 + 16 + 1023) & ~1023;  // residual tiles (optional), then B ring
};

struct Tile { int w0, h0, n0, nbase; };

__device__ __forceinline__ uint32_t fast_div(uint32_t x, const FastDiv& f) { return f.d > 1 ? __umulhi(x, f.mul) >> f.shr : x; }

__device__ __forceinline__ Tile decode_tile(int w, int n_ntiles, uint32_t rank, const Geom3& g, int BN) {
  const uint32_t mp = fast_div((uint32_t)w, g.d_nt), nt = (uint32_t)w - mp * (uint32_t)n_ntiles;
  const uint32_t mt = 2 * mp + rank;
  Tile t;
  const uint32_t r1 = fast_div(mt, g.d_tw), tw = mt - r1 * (uint32_t)g.tiles_w;     // mt = (n * tiles_h + th) * tiles_w + tw
  const uint32_t r2 = fast_div(r1, g.d_th), th = r1 - r2 * (uint32_t)g.tiles_h;
  t.w0 = (int)tw * g.bw; t.h0 = (int)th * g.bh; t.n0 = (int)r2 * g.bn;
  t.nbase = (int)nt * BN;
  return t;
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      :: "l"(reinterpret_cast<uint64_t>(m)), "r"(tc::smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// pull one TMA box into L2 (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               :: "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// development aid (eo_debug_conv_trace, TRACE instantiation only): per-CTA counters, slot = 0 lifetime,
// 1 MMA waits on operands, 2 MMA waits on a free accumulator, 3 epilogue waits on the accumulator,
// 4 epilogue busy, 5 producer waits on free stages, 6 tiles, 7 transform warps busy (clock64 ticks)
// slots 8.. (EO_TRACE_EXT=1, buffer of 2 * n * 8 counters): 8 MMA waits on operand A only, 9 transform warps wait
// for a free stage
__device__ __forceinline__ void trace_put(const Epi3& ep, int slot, long long v) {
  if (!ep.trace || (int)blockIdx.x >= ep.trace_n) return;
  if (slot < 8) ep.trace[(long long)blockIdx.x * 8 + slot] = v;
  else if (ep.trace_ext) ep.trace[((long long)ep.trace_n * (slot >> 3) + blockIdx.x) * 8 + (slot & 7)] = v;   // trace_ext: 4 * n * 8 counters
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
           const KEnt3* __restrict__ ents, int nent, const XEnt3* __restrict__ xents, int n_xent, Geom3 g, int B, int BN,
           int SB, int TPB, int SAR, int SAG, int a_stage, int GRP, int n_work, int n_ntiles, Epi3 ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem_a = smem_raw + ((1024 - (tc::smem_u32(smem_raw) & 1023)) & 1023);   // operand-A stages
  uint8_t* smem = smem_a + (SAR + SAG) * a_stage;                                    // everything else (a_stage: A_STAGE, or PLAIN_BYTES for plain-tile launches)
  uint8_t* smem_g = smem_a + SAR * a_stage;                                          // the transform warps' stages
  KEnt3* tab = reinterpret_cast<KEnt3*>(smem + Smem::TAB_OFF);
  XEnt3* xtab = reinterpret_cast<XEnt3*>(smem + Smem::XTAB_OFF);
  float* sstat = reinterpret_cast<float*>(smem + Smem::STAT_OFF);
  // Operand A has two rings, one per filler, so that nobody has to track a stage it does not fill: stages
  // loaded by TMA as is (both CTAs' bytes complete on the leader's r_full), and stages the transform warps
  // write with GroupNorm folded in (both CTAs' warps arrive on the leader's g_ready).
  uint64_t* r_full = reinterpret_cast<uint64_t*>(smem + Smem::BAR_OFF);
  uint64_t* r_empty = r_full + SA_MAX;
  uint64_t* g_ready = r_empty + SA_MAX;
  uint64_t* g_empty = g_ready + SA_MAX;
  uint64_t* b_full = g_empty + SA_MAX;
  uint64_t* b_empty = b_full + MAX_SB;
  uint64_t* tmem_full = b_empty + MAX_SB;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_full = tmem_empty + 2;
  uint64_t* res_empty = res_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_empty + 2);
  const bool has_res = ep.has_res != 0;
  uint8_t* res_sm = smem + Smem::VAR_OFF;
  const int b_bytes = (BN / 2) * 128;          // one 64-deep weight tile (this CTA's half of the BN rows)
  const int b_stage = TPB * b_bytes;           // a weight stage holds TPB tiles: the taps of one kernel row of a patch
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
    for (int s = 0; s < SA_MAX; ++s) {
      tc::mbar_init(&r_full[s], 1); tc::mbar_init(&r_empty[s], 1);
      tc::mbar_init(&g_ready[s], 2 * XF_WARPS); tc::mbar_init(&g_empty[s], 1);
    }
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
  for (int i = threadIdx.x; i < n_xent * (int)(sizeof(XEnt3) / 16); i += blockDim.x)
    reinterpret_cast<uint4*>(xtab)[i] = reinterpret_cast<const uint4*>(xents)[i];
  tc::tc_fence_before();
  tc::cluster_sync_all();                    // peer barriers are initialised past here
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // everything above touched only this launch's own constants (operand tables, tensor maps) and on-chip state: the CTA
  // may have been scheduled while the previous kernel of the forward was still running (common.cuh)
  pdl_wait();
  pdl_trigger();

  // 640 threads leave 96 registers each (61440 for the CTA); the epilogue warps need more and the issue
  // warps far fewer: warpgroup 0 drops to 40, the epilogue warpgroups grow to 104, the transform warpgroups
  // (two patches of operand data in registers) to 112 (40 + 2*104 + 2*112 = 472 <= 480 = 5 * 96)

  // Producer and MMA warps: the WHOLE warp walks the loops (warp-uniform control flow keeps addresses,
  // coordinates and descriptors in uniform registers) and one elected lane issues.  Under
  // `if (lane == 0)` ptxas wraps every TMA / MMA instruction in a read-lane loop and the issuing
  // thread, not the tensor pipe, becomes the limiter (measured: 650 clk per 64-deep K block).
  if (warp == 0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    // ------------------------------------------------------------------ operand producer
    Ring ra, rb;
    long long tr_wait = 0;
    const uint32_t r_full_l = tc::mapa_u32(tc::smem_u32(&r_full[0]), 0);   // the leader's barriers
    const uint32_t b_full_l = tc::mapa_u32(tc::smem_u32(&b_full[0]), 0);
    const int b_row = (int)rank * (BN / 2);
    KEnt3 en_next = tab[0];           // the table is read one entry ahead: an ld.shared takes ~200 clk here
    for (int w = cid; w < n_work; w += ncl) {
      const Tile t = decode_tile(w, n_ntiles, rank, g, BN);
      if (GRP > 1) {
        // plain tiles only (no patch, no GroupNorm-folded load): GRP loads per operand stage, their weight tiles in one
        // weight stage (b_stage = GRP tiles)
        for (int e = 0; e < nent; e += GRP) {
          const int ng = nent - e < GRP ? nent - e : GRP;
          { TRACE_T0(); tc::mbar_wait(&r_empty[ra.i], ra.ph ^ 1); TRACE_ACC(tr_wait); }
          if (tc::elect_one()) {
            if (rank == 0) tc::mbar_arrive_expect_tx(&r_full[ra.i], 2u * (uint32_t)ng * PLAIN_BYTES);
            for (int u = 0; u < ng; ++u) {
              const KEnt3 en = tab[e + u];
              const CUtensorMap* ma = en.seg == 0 ? &mapA0 : (en.seg == 1 ? &mapA1 : &mapA2);
              const int dh = (int)(short)(en.dhw & 0xffff), dw = en.dhw >> 16;
              tc::tma2_load_4d(smem_a + ra.i * a_stage + u * PLAIN_BYTES, ma, r_full_l + ra.i * 8, en.c0,
                               t.w0 * en.sc + dw, t.h0 * en.sc + dh, t.n0 + en.dn);
            }
          }
          __syncwarp();
          ra.next((uint32_t)SAR);
          { TRACE_T0(); tc::mbar_wait(&b_empty[rb.i], rb.ph ^ 1); TRACE_ACC(tr_wait); }
          if (tc::elect_one()) {
            if (rank == 0) tc::mbar_arrive_expect_tx(&b_full[rb.i], 2u * (uint32_t)ng * (uint32_t)b_bytes);
            for (int u = 0; u < ng; ++u)
              tc::tma2_load_2d(b_sm + rb.i * b_stage + u * b_bytes, &mapB, b_full_l + rb.i * 8, tab[e + u].kofs, t.nbase + b_row);
          }
          __syncwarp();
          rb.next((uint32_t)SB);
        }
        continue;
      }
      for (int e = 0; e < nent; ++e) {
        const KEnt3 en = en_next;
        en_next = tab[e + 1 == nent ? 0 : e + 1];
        if (!en.gn) { TRACE_T0(); tc::mbar_wait(&r_empty[ra.i], ra.ph ^ 1); TRACE_ACC(tr_wait); }
        const CUtensorMap* ma = en.seg == 0 ? &mapA0 : (en.seg == 1 ? &mapA1 : &mapA2);
        if (!en.gn && tc::elect_one()) {
          const int dh = (int)(short)(en.dhw & 0xffff), dw = en.dhw >> 16;
          const int cw = t.w0 * en.sc + (en.patch ? -1 : dw), chh = t.h0 * en.sc + (en.patch ? -1 : dh);
          const uint32_t bytes = en.patch ? PATCH_BYTES : PLAIN_BYTES;
          uint8_t* dst = smem_a + ra.i * a_stage;
          // one arrival per phase: the leader's producer, which posts the byte count of BOTH CTAs' loads
          if (rank == 0) tc::mbar_arrive_expect_tx(&r_full[ra.i], 2 * bytes);
          tc::tma2_load_4d(dst, ma, r_full_l + ra.i * 8, en.c0, cw, chh, t.n0 + en.dn);
        }
        // a GroupNorm-folded patch is fetched by the transform warps with plain loads, one patch ahead of the
        // tensor pipe: too late to hide an HBM miss (measured: 2340 clk per patch, 1300 of them load latency).
        // Pull the SAME entry's patch of this CTA's NEXT tile into L2 now, a whole tile ahead.
        if (en.gn && w + ncl < n_work && tc::elect_one()) {
          const Tile tn = decode_tile(w + ncl, n_ntiles, rank, g, BN);
          if (tn.n0 < B) tma_prefetch_4d(ma, en.c0, tn.w0 - 1, tn.h0 - 1, tn.n0 + en.dn);
        }
        __syncwarp();
        if (!en.gn) ra.next((uint32_t)SAR);
        // weight tiles of this load: 9 for a patch (TPB per stage), 1 otherwise
        const int pnc = (en.patch >> 16) & 15;                         // columns of the patch's tap window
        const int nb = en.patch ? ((en.patch >> 8) & 15) * pnc : 1;
        const int per = en.patch ? (TPB > 1 ? pnc : 1) : 1;             // TPB > 1: one kernel row of taps per stage
        for (int j = 0; j < nb; j += per) {
          { TRACE_T0(); tc::mbar_wait(&b_empty[rb.i], rb.ph ^ 1); TRACE_ACC(tr_wait); }
          if (tc::elect_one()) {
            // one arrival per phase: the leader's producer, which posts the byte count of BOTH CTAs'
            // loads; the peer's TMA only completes transactions on the leader's barrier
            if (rank == 0) tc::mbar_arrive_expect_tx(&b_full[rb.i], 2 * per * b_bytes);
            for (int u = 0; u < per; ++u)
              tc::tma2_load_2d(b_sm + rb.i * b_stage + u * b_bytes, &mapB, b_full_l + rb.i * 8, en.kofs + (j + u) * BK,
                               t.nbase + b_row);
          }
          __syncwarp();
          rb.next((uint32_t)SB);
        }
      }
    }
    if (TRACE && lane == 0) trace_put(ep, 5, tr_wait);
  } else if (warp == 1) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    if (rank == 0) {
      const uint32_t idesc = tc::make_idesc_bf16(2 * BM, BN, 0, 0);
      const uint64_t bdesc0 = tc::make_sw128_desc(tc::smem_u32(b_sm));
      const uint32_t b_step = (uint32_t)b_stage >> 4, b_tile = (uint32_t)b_bytes >> 4;
      Ring ra, rg, rb;
      uint32_t it = 0;
      long long tr_ops = 0, tr_acc = 0, tr_a = 0;
      int pg_next = tab[0].gn | (tab[0].patch << 4);      // (patch, gn) of the next entry, read one entry ahead
      for (int w = cid; w < n_work; w += ncl, ++it) {
        const uint32_t ab = it & 1;
        { TRACE_T0(); tc::mbar_wait_cluster(&tmem_empty[ab], ((it >> 1) & 1) ^ 1); TRACE_ACC(tr_acc); }   // both CTAs' epilogues drained this buffer
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * ACC_STRIDE;
        uint32_t acc = 0;
        if (GRP > 1) {
          for (int e = 0; e < nent; e += GRP) {
            const int ng = nent - e < GRP ? nent - e : GRP;
            { TRACE_T0(); tc::mbar_wait(&r_full[ra.i], ra.ph); TRACE_ACC(tr_ops); TRACE_ACC(tr_a); }
            { TRACE_T0(); tc::mbar_wait(&b_full[rb.i], rb.ph); TRACE_ACC(tr_ops); }
            tc::tc_fence_after();
            const uint64_t adesc = tc::make_sw128_desc(tc::smem_u32(smem_a + ra.i * a_stage));
            const uint64_t bdesc = bdesc0 + (uint64_t)(rb.i * b_step);
            if (tc::elect_one()) {
              for (int u = 0; u < ng; ++u) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                  tc::umma2_f16_ss(d_tmem, tc::desc_advance(adesc + (uint64_t)(u * (PLAIN_BYTES >> 4)), k * 32),
                                   tc::desc_advance(bdesc + (uint64_t)(u * b_tile), k * 32), idesc, (k | u) ? 1u : acc);
              }
              tc::umma2_commit_mc(&b_empty[rb.i], 3);
              tc::umma2_commit_mc(&r_empty[ra.i], 3);
              if (e + ng >= nent) tc::umma2_commit_mc(&tmem_full[ab], 3);
            }
            __syncwarp();
            acc = 1;
            rb.next((uint32_t)SB);
            ra.next((uint32_t)SAR);
          }
          continue;
        }
        for (int e = 0; e < nent; ++e) {
          const int patch = pg_next >> 4;
          const bool gn = (pg_next & 15) != 0;
          { const int e1 = e + 1 == nent ? 0 : e + 1; pg_next = tab[e1].gn | (tab[e1].patch << 4); }
          {
            TRACE_T0();
            if (gn) tc::mbar_wait_cluster(&g_ready[rg.i], rg.ph);
            else tc::mbar_wait(&r_full[ra.i], ra.ph);
            TRACE_ACC(tr_ops); TRACE_ACC(tr_a);
          }
          tc::tc_fence_after();
          const uint32_t a_base = tc::smem_u32(gn ? smem_g + rg.i * a_stage : smem_a + ra.i * a_stage);
          uint64_t* a_release = gn ? &g_empty[rg.i] : &r_empty[ra.i];
          const bool last_e = e == nent - 1;
          // one weight stage: wait for it, issue the 4 K=16 MMAs of each of its NT 64-deep K blocks (tile u
          // against operand descriptor adesc + u * a_step), release it.  Issuing blocks the thread at the
          // tensor pipe's pace and the wait / fence / commit around it are not hidden behind the pipe's
          // short queue (measured: 415 clk per K block whatever the tile width below 256), so narrow
          // tiles take a whole kernel row of taps per stage.
          auto kgroup = [&](uint64_t adesc, uint32_t a_step, int nt, bool last_of_a) {
            { TRACE_T0(); tc::mbar_wait(&b_full[rb.i], rb.ph); TRACE_ACC(tr_ops); }
            tc::tc_fence_after();
            const uint64_t bdesc = bdesc0 + (uint64_t)(rb.i * b_step);
            if (tc::elect_one()) {
              for (int u = 0; u < nt; ++u) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                  tc::umma2_f16_ss(d_tmem, tc::desc_advance(adesc + (uint64_t)(u * a_step), k * 32),
                                   tc::desc_advance(bdesc + (uint64_t)(u * b_tile), k * 32), idesc, (k | u) ? 1u : acc);
              }
              tc::umma2_commit_mc(&b_empty[rb.i], 3);                   // frees the weight stage in both CTAs
              if (last_of_a) tc::umma2_commit_mc(a_release, 3);          // ... and the activation stage
              if (last_of_a && last_e) tc::umma2_commit_mc(&tmem_full[ab], 3);   // accumulator complete
            }
            __syncwarp();
            acc = 1;
            rb.next((uint32_t)SB);
          };
          if (patch) {
            // tap (kh, kw) of a patch starts kh patch rows + kw pixels into it
            const uint64_t ad0 = tc::make_sw128_desc_sbo(a_base, PATCH_W * 128);
            if (patch != TC_PATCH_3X3) {
              // a sub-window of the 3x3 neighbourhood (the 2x2 corner of a sub-pixel convolution): same patch, fewer taps
              const int r0 = (patch >> 4) & 15, nr = (patch >> 8) & 15, c0 = (patch >> 12) & 15, nc = (patch >> 16) & 15;
              if (TPB == 3) {
                for (int kh = 0; kh < nr; ++kh) kgroup(ad0 + (uint64_t)(((r0 + kh) * PATCH_W + c0) * 8), 8u, nc, kh == nr - 1);
              } else {
                for (int kh = 0; kh < nr; ++kh)
                  for (int kw = 0; kw < nc; ++kw)
                    kgroup(ad0 + (uint64_t)(((r0 + kh) * PATCH_W + c0 + kw) * 8), 0u, 1, kh == nr - 1 && kw == nc - 1);
              }
            } else if (TPB == 3) {
#pragma unroll
              for (int kh = 0; kh < 3; ++kh) kgroup(ad0 + (uint64_t)(kh * PATCH_W * 8), 8u, 3, kh == 2);
            } else {
#pragma unroll
              for (int j = 0; j < 9; ++j)
                kgroup(ad0 + (uint64_t)(((j / 3) * PATCH_W + (j % 3)) * 8), 0u, 1, j == 8);
            }
          } else {
            kgroup(tc::make_sw128_desc(a_base), 0u, 1, true);
          }
          if (gn) rg.next((uint32_t)SAG); else ra.next((uint32_t)SAR);
        }
      }
      if (TRACE && lane == 0) { trace_put(ep, 1, tr_ops); trace_put(ep, 2, tr_acc); trace_put(ep, 6, it); trace_put(ep, 8, tr_a); }
    }
  } else if (warp == 2 || warp == 3) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    // ------------------------------------------------------------------ residual producer (warp 3 idles)
    if (has_res && warp == 2) {
      if (lane == 0) tc::tma_prefetch_desc(&mapRes);
      uint32_t uses0 = 0, uses1 = 0;
      const int nchunks = BN / 64;
      for (int w = cid; w < n_work; w += ncl) {
        const Tile t = decode_tile(w, n_ntiles, rank, g, BN);
        for (int c = 0; c < nchunks; ++c) {
          const uint32_t rbuf = c & 1;                  // chunk c belongs to epilogue set c & 1
          tc::mbar_wait(&res_empty[rbuf], ((rbuf ? uses1 : uses0) & 1) ^ 1);
          if (rbuf) ++uses1; else ++uses0;
          if (tc::elect_one()) {
            tc::mbar_arrive_expect_tx(&res_full[rbuf], STG_BYTES);
            tc::tma_load_4d(res_sm + rbuf * STG_BYTES, &mapRes, &res_full[rbuf], t.nbase + c * 64, t.w0, t.h0, t.n0);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= XF_WARP0) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    // ------------------------------------------------------------------ operand transform
    // GroupNorm affine (+ SiLU) of the activation, applied in place to the stage TMA just filled
    // (reference: normalization / nn.SiLU in front of every conv, unet_openai.py:313-316, :337-342,
    // :411-412; same arithmetic as k_gn_apply: fp32 x*scale+shift, tanh-form SiLU, round to bf16).
    // Thread <-> one 16-byte chunk (8 channels) of every 16th pixel; the 128B swizzle TMA applied
    // (chunk ^ pixel & 7) is undone in the address.  Pixels of a halo patch that lie outside the image
    // stay zero: the convolution pads the NORMALISED activation with zeros.
    const int xt = threadIdx.x - XF_WARP0 * 32;        // 0..255
    const int ch8 = xt & 7, pl = xt >> 3;              // channel chunk, pixel lane (0..31)
    const uint32_t g_ready_l = tc::mapa_u32(tc::smem_u32(&g_ready[0]), 0);
    // the thread's 6 pixels q = pl + 32 i of a patch never change: byte offset of its 16-byte chunk in the
    // (swizzled) stage in the low 16 bits, the patch borders the pixel lies on in bits 16.. (top, bottom,
    // left, right; bit 20 = beyond the patch); plin = pixel offset inside the image relative to the
    // patch origin
    uint32_t ptab[XF_PIX];
    uint32_t plin[XF_PIX];
#pragma unroll
    for (int i = 0; i < XF_PIX; ++i) {
      const int q = pl + 4 * XF_WARPS * i;
      const int ph = q / PATCH_W, pw = q - ph * PATCH_W;
      uint32_t edge = (ph == 0 ? 1u : 0u) | (ph == PATCH_H - 1 ? 2u : 0u) | (pw == 0 ? 4u : 0u) | (pw == PATCH_W - 1 ? 8u : 0u);
      if (q >= PATCH_W * PATCH_H) edge = 16u;
      ptab[i] = (uint32_t)(q * 128 + ((ch8 ^ (q & 7)) << 4)) | (edge << 16);
      plin[i] = (uint32_t)(ph * g.W + pw);
    }
    // Cursor over the GroupNorm-folded operand loads of this CTA, in the order the MMA warp consumes them: load j of
    // the compact list `xtab` (one XEnt3 per such load, resolved pointers included) of tile w.  With the tensor core
    // streaming operands out of the same shared memory an ld.shared takes ~200 clk, and the walk used to make five
    // DEPENDENT reads of the operand table per patch (the gn flags while seeking, the entry for its affine rows, again
    // for the patch loads, again for the SiLU flag: 1280 clk of set-up per patch in the per-CTA trace, on the critical
    // path of every Cout = 128 layer).  Now every entry is read ONCE, a whole patch before it is needed.
    struct Cur { int w, j; Ring ring; Tile t; bool valid; };
    auto advance = [&](Cur& c) {
      c.ring.next((uint32_t)SAG);
      if (++c.j == n_xent) {
        c.j = 0; c.w += ncl;
        if (c.w >= n_work) { c.valid = false; return; }
        c.t = decode_tile(c.w, n_ntiles, rank, g, BN);
      }
    };
    auto read_xent = [&](int j) -> XEnt3 {
      XEnt3 x;
      const uint4* q = reinterpret_cast<const uint4*>(xtab + j);
      uint4* d = reinterpret_cast<uint4*>(&x);
      d[0] = q[0]; d[1] = q[1]; d[2] = q[2];
      return x;
    };
    // two patches of 6 x 16 bytes per thread in registers: the loads of patch i+1 are issued BEFORE patch i is
    // transformed (a first version issued them after the hand-over, because fence.proxy.async waits for every
    // load the thread has in flight: the patch then cost its full load latency, 2340 clk against 1000 of
    // arithmetic; issued a whole transform earlier they have landed by the time the fence is reached)
    uint4 bufa[XF_PIX], bufb[XF_PIX];
    float4 na0, na1, nb0, nb1;                 // scale / shift of the next patch
    na0 = na1 = nb0 = nb1 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto edges_of = [&](const Tile& t) -> uint32_t {   // image borders the tile touches (+ "beyond the patch")
      return ((t.h0 == 0 ? 1u : 0u) | (t.h0 + 16 == g.H ? 2u : 0u) | (t.w0 == 0 ? 4u : 0u) | (t.w0 + 8 == g.W ? 8u : 0u) | 16u) << 16;
    };
    auto load_affine = [&](const Cur& c, const XEnt3& xe) {
      if (c.t.n0 >= B) return;
      const long long o = (long long)c.t.n0 * xe.gld + ch8 * 8;
      const float* gsc = reinterpret_cast<const float*>(xe.gsc) + o;
      const float* gsh = reinterpret_cast<const float*>(xe.gsh) + o;
      na0 = __ldg(reinterpret_cast<const float4*>(gsc));
      na1 = __ldg(reinterpret_cast<const float4*>(gsc + 4));
      nb0 = __ldg(reinterpret_cast<const float4*>(gsh));
      nb1 = __ldg(reinterpret_cast<const float4*>(gsh + 4));
    };
    auto load_patch = [&](const Cur& c, const XEnt3& xe, uint4 (&buf)[XF_PIX]) {
      const long long pix0 = ((long long)c.t.n0 * g.H + (c.t.h0 - 1)) * g.W + (c.t.w0 - 1);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(xe.src) + pix0 * (long long)xe.cs2 + ch8 * 16;
      const uint32_t te = c.t.n0 < B ? edges_of(c.t) : 0xffffffffu;
#pragma unroll
      for (int i = 0; i < XF_PIX; ++i)      // plin * cs2 < 2^32 (one image plane of <= 2048 channels)
        if (!(ptab[i] & te)) buf[i] = __ldg(reinterpret_cast<const uint4*>(src + (unsigned long long)plin[i] * xe.cs2));
    };
    long long tr_xb = 0, tr_xw = 0;
    // transform the patch in `buf` (cursor `cur`, entry `xc`) while the loads of the next one (`nxt`, entry `xn`) are in
    // flight in `bufn`; `xn2` = the entry after that, read now and first used a whole patch from now
    long long tr_xp = 0;
    long long tr_f[4 + XF_PIX] = {};        // TRACE: advance + entry read, patch-load issue, each pixel chunk, fence, arrive
    XEnt3 xc, xn;
    auto step = [&](Cur& cur, uint4 (&buf)[XF_PIX], uint4 (&bufn)[XF_PIX]) {
      const long long t_p0 = TRACE ? clock64() : 0;
      const bool silu = xc.silu != 0;
      const uint32_t te = cur.t.n0 < B ? edges_of(cur.t) : 0xffffffffu;
      // scale / shift of this thread's 8 channels.  With SiLU they are halved: silu(y) = h + h tanh(h),
      // h = y/2 (common.cuh silu_f), and 0.5 * fma(x, s, b) == fma(x, 0.5 s, 0.5 b) exactly.
      const float hf = silu ? 0.5f : 1.0f;
      // (packed fp32 pairs: FFMA2 does the affine and the SiLU recombination of a bf16 pair in one issue slot each)
      const uint64_t a01 = tc::pack2(na0.x * hf, na0.y * hf), a23 = tc::pack2(na0.z * hf, na0.w * hf);
      const uint64_t a45 = tc::pack2(na1.x * hf, na1.y * hf), a67 = tc::pack2(na1.z * hf, na1.w * hf);
      const uint64_t b01 = tc::pack2(nb0.x * hf, nb0.y * hf), b23 = tc::pack2(nb0.z * hf, nb0.w * hf);
      const uint64_t b45 = tc::pack2(nb1.x * hf, nb1.y * hf), b67 = tc::pack2(nb1.z * hf, nb1.w * hf);
      Cur nxt = cur;
      advance(nxt);
      const XEnt3 xn2 = read_xent(nxt.j + 1 == n_xent ? 0 : nxt.j + 1);
      if (TRACE) { const long long c = clock64(); tr_f[0] += c - t_p0; }
      if (nxt.valid) {
        load_affine(nxt, xn);
        if (TRACE) tr_f[1] -= clock64();
        load_patch(nxt, xn, bufn);
        if (TRACE) tr_f[1] += clock64();
      }
      if (TRACE) tr_xp += clock64() - t_p0;
      { TRACE_T0(); tc::mbar_wait(&g_empty[cur.ring.i], cur.ring.ph ^ 1); TRACE_ACC(tr_xw); }   // the MMAs that read this stage have retired
      const long long t_b0 = TRACE ? clock64() : 0;
      long long t_it = t_b0;
      uint8_t* st = smem_g + cur.ring.i * a_stage;
      // one bf16 pair: affine (+ SiLU) in fp32, back to bf16
      auto xf2 = [&](uint32_t in, uint64_t sc, uint64_t sh, bool act) -> uint32_t {
        uint64_t x = tc::fma2(tc::pack2(__uint_as_float(in << 16), __uint_as_float(in & 0xffff0000u)), sc, sh);
        float x0, x1;
        tc::unpack2(x, x0, x1);
        if (act) {
          float t0, t1;
          asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(x0));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(x1));
          tc::unpack2(tc::fma2(x, tc::pack2(t0, t1), x), x0, x1);
        }
        __nv_bfloat162 o = __floats2bfloat162_rn(x0, x1);
        return *reinterpret_cast<uint32_t*>(&o);
      };
      // (one copy of this body: the kernel's warp roles already crowd the instruction cache -- with the
      // loop duplicated per activation flag and per register buffer the epilogue warps stalled on fetches)
#pragma unroll
      for (int i = 0; i < XF_PIX; ++i) {
        if (ptab[i] & (16u << 16)) continue;                    // beyond the patch (only the last i)
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        const bool in_img = !(ptab[i] & te);
        if (in_img) v = buf[i];
        if (in_img) {
          v.x = xf2(v.x, a01, b01, silu);
          v.y = xf2(v.y, a23, b23, silu);
          v.z = xf2(v.z, a45, b45, silu);
          v.w = xf2(v.w, a67, b67, silu);
        }
        *reinterpret_cast<uint4*>(st + (ptab[i] & 0xffffu)) = v;    // zeros outside the image
        if (TRACE) { const long long c = clock64(); tr_f[2 + i] += c - (i ? t_it : t_b0); t_it = c; }
      }
      tc::fence_proxy_async_smem();     // generic-proxy writes -> visible to the tensor core's operand reads
      if (TRACE) { const long long c = clock64(); tr_f[2 + XF_PIX] += c - t_it; t_it = c; }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_remote(g_ready_l + cur.ring.i * 8);
      if (TRACE) { tr_f[3 + XF_PIX] += clock64() - t_it; }
      if (TRACE) tr_xb += clock64() - t_b0;
      cur = nxt;
      xc = xn; xn = xn2;
    };
    Cur cur{cid, 0, Ring(), Tile(), n_xent > 0 && cid < n_work};
    if (cur.valid) {
      cur.t = decode_tile(cur.w, n_ntiles, rank, g, BN);
      xc = read_xent(0);
      xn = read_xent(n_xent > 1 ? 1 : 0);
      load_affine(cur, xc);
      load_patch(cur, xc, bufa);
    }
    long long tr_xl = 0;
    while (cur.valid) {
      step(cur, bufa, bufb);
      TRACE_T0();
#pragma unroll
      for (int i = 0; i < XF_PIX; ++i) bufa[i] = bufb[i];
      if (TRACE) { asm volatile("" :: "r"(bufa[0].x), "r"(bufa[XF_PIX - 1].w) : "memory"); }
      TRACE_ACC(tr_xl);
    }
    if (TRACE && xt == 0) {
      trace_put(ep, 7, tr_xb); trace_put(ep, 9, tr_xw); trace_put(ep, 10, tr_xl); trace_put(ep, 11, tr_xp);
#pragma unroll
      for (int i = 0; i < 4 + XF_PIX; ++i) trace_put(ep, 16 + i, tr_f[i]);
    }
  } else if (warp >= EPI_WARP0) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ------------------------------------------------------------------ epilogue
    // Two sets of four warps (one warp per TMEM lane quadrant in each); set s takes the 64-channel
    // chunks c = s, s+2, ... of every tile.  A warp owns 32 accumulator rows end to end: TMEM ->
    // registers -> its 4 KB slice of the staging tile -> its own TMA store (a 32-pixel sub-box), so
    // the only cross-warp synchronisation is the per-tile combine of the GroupNorm sums.
    const int ew = warp - EPI_WARP0;               // 0..7
    const int set = ew >> 2;
    const int q = warp & 3;                        // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;
    const int et = ew * 32 + lane;                 // 0..255
    const int nn = row / (g.bw * g.bh);
    const int pix_w = row % g.bw, pix_h = (row / g.bw) % g.bh;      // this thread's pixel inside the tile
    const bool to_nchw = ep.out_nchw != nullptr;
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
    long long tr_wait = 0, tr_busy = 0, tr_e2 = 0, tr_e3 = 0, tr_e5 = 0, tr_e6 = 0;
    // Tiles inside one image (every halo-patch grid): bias[n] + bias_nc[image][n] of this warp's channels sit in a
    // warp-private shared-memory row, rebuilt only when the image or the channel tile changes -- the 16 global
    // loads per 32 columns they replace cost an L2 round trip each time (the L1 is all shared memory here).
    float* wb = reinterpret_cast<float*>(smem + Smem::WB_OFF) + ew * 128;
    const bool cache_bias = single_image && (ep.bias || ep.bias_nc);
    int wb_img = -1, wb_nbase = -1;
    for (int w = cid; w < n_work; w += ncl, ++it) {
      const Tile t = decode_tile(w, n_ntiles, rank, g, BN);
      const int n_img = t.n0 + nn;
      if (cache_bias && t.n0 < B && (t.n0 != wb_img || t.nbase != wb_nbase)) {
        __syncwarp();
        for (int slot = 0, c = set; c < nchunks; c += 2, ++slot) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int n = t.nbase + c * 64 + hh * 32 + lane;
            float v = ep.bias ? __ldg(ep.bias + n) : 0.f;
            if (ep.bias_nc) v += __ldg(ep.bias_nc + (long long)t.n0 * ep.ld_bias_nc + n);
            wb[slot * 64 + hh * 32 + lane] = v;
          }
        }
        wb_img = t.n0; wb_nbase = t.nbase;
        __syncwarp();
      }
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
        if (has_res) { tc::mbar_wait(&res_full[set], res_uses & 1); ++res_uses; }
        // the TMA store this warp issued from its staging rows one chunk ago has read them out
        if (lane == 0) bulk_wait_read0();      // (bulk groups belong to the issuing thread: always lane 0)
        __syncwarp();
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {       // not unrolled: instruction-cache footprint
          const int n = t.nbase + c * 64 + half * 32;
          float f[32];
          {
            uint32_t v[32];                       // 32 columns at a time keeps the warp under 128 registers
            TRACE_T0();
            tc::tmem_ld_32x32(taddr + half * 32, v);
            tc::tmem_ld_wait();
            TRACE_ACC(tr_e2);
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          }
          const long long t_e3 = TRACE ? clock64() : 0;
          if (half == 1 && c + 2 >= nchunks) {     // this warp's share of the accumulator now lives in registers
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_cluster_relaxed(tmem_empty_leader + ab * 8);
          }
          if (to_nchw && half * 32 >= ep.out_nchw_C) continue;      // padded channels of the head: nothing to write
          if (cache_bias) {
            if (valid) {
              const float* wr = wb + ((c - set) >> 1) * 64 + half * 32;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(wr + j);
                f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
              }
            }
          } else if (ep.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n + j));
              f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
            }
          }
          if (bnc && !cache_bias) {
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
          if (to_nchw) {
            // network output: fp32 planes, 8 consecutive pixels of a row per 8 lanes (32-byte segments)
            if (valid) {
              const long long plane = (long long)g.H * g.W;
              float* op = ep.out_nchw + (((long long)n_img * ep.out_nchw_C + half * 32) * g.H + (t.h0 + pix_h)) * g.W + t.w0 + pix_w;
              const int nj = ep.out_nchw_C - half * 32;          // warp-uniform: the loop leaves after the real channels
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (j >= nj) break;
                *op = f[j];
                op += plane;
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              uint4 o;
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int e2 = 0; e2 < 4; ++e2) h2[e2] = __floats2bfloat162_rn(f[k * 8 + e2 * 2], f[k * 8 + e2 * 2 + 1]);
              *reinterpret_cast<uint4*>(stg + row_off + ((((uint32_t)(half * 4 + k)) ^ swz) << 4)) = o;
            }
          }
          if (TRACE) tr_e3 += clock64() - t_e3;
        }
        if (has_res) {
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&res_empty[set]);
        }
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && !to_nchw) {
          tma_store_4d(&mapOut, stg + row0 * 128, t.nbase + c * 64, t.w0 + sub_w, t.h0 + sub_h, t.n0 + sub_n);
          bulk_commit();
        }
        __syncwarp();
        const long long t_e5 = TRACE ? clock64() : 0;
        if (do_stats && warp_valid) {
          // GroupNorm statistics of the tensor being written: per-channel sum and sum of squares of this warp's 32
          // pixel rows, read back from the bf16 staging rows the TMA store is draining (the values the consumer will
          // normalise).  Lane l owns channels 2l, 2l+1 of the chunk: one conflict-free 4-byte column read per row,
          // 32 rows in fixed order -- 7 instructions per row in a rolled loop, against 62 shuffles + 124 selects per
          // 32 columns for the register-resident column sums it replaces (k_gn_finalize_ch reduces per group).
          // Deterministic: fixed-order fp32 partial sums; only the cross-tile accumulation is atomic, and in double.
          const uint32_t cchunk = (uint32_t)lane >> 2, csub = ((uint32_t)lane & 3u) * 4u;
          const uint8_t* srow = stg + row0 * 128 + csub;
          // all 32 reads are issued before the first is consumed: with the tensor core streaming operands out of the
          // same shared memory a read takes ~200 clk (traced: 54 clk per row when issued four at a time)
          uint32_t wv[32];
#pragma unroll
          for (uint32_t r = 0; r < 32; ++r)
            wv[r] = *reinterpret_cast<const uint32_t*>(srow + r * 128u + ((cchunk ^ (r & 7u)) << 4));
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
          for (uint32_t r = 0; r < 32; ++r) {
            const float x0 = __uint_as_float(wv[r] << 16), x1 = __uint_as_float(wv[r] & 0xffff0000u);
            s0 += x0; s1 += x1;
            q0 = fmaf(x0, x0, q0); q1 = fmaf(x1, x1, q1);
          }
          const int col = c * 64 + 2 * lane;
          if (single_image) {
            *reinterpret_cast<float4*>(sstat + (q * 256 + col) * 2) = make_float4(s0, q0, s1, q1);
          } else if (t.nbase + col < Cout) {
            double* dst = ep.stats + ((long long)n_img * Cout + t.nbase + col) * 2;
            atomicAdd(dst, (double)s0); atomicAdd(dst + 1, (double)q0);
            atomicAdd(dst + 2, (double)s1); atomicAdd(dst + 3, (double)q1);
          }
        }
        if (TRACE) tr_e5 += clock64() - t_e5;
      }
      const long long t_e6 = TRACE ? clock64() : 0;
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
      if (TRACE) { tr_e6 += clock64() - t_e6; tr_busy += clock64() - t_busy0; }
    }
    if (lane == 0) bulk_wait0();
    __syncwarp();
    if (TRACE && et == 0) {
      trace_put(ep, 3, tr_wait); trace_put(ep, 4, tr_busy);
      trace_put(ep, 12, tr_e2); trace_put(ep, 13, tr_e3); trace_put(ep, 14, tr_e5); trace_put(ep, 15, tr_e6);
    }
    tc::tc_fence_before();
  }
  __syncwarp();
  tc::cluster_sync_all();     // nobody leaves while the peer still uses its smem/TMEM
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc2(tmem_base, TMEM_COLS);
  }
  if (TRACE && threadIdx.x == 0) {
    trace_put(ep, 0, clock64() - t_entry);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
#ifdef EO_DEVTOOLS
static bool env_flag(const char* name, bool dflt) {
  const char* e = std::getenv(name);
  if (!e || !e[0]) return dflt;
  return e[0] != '0';
}
#endif

int tc_conv3_plan_fill(const TcConvParams& p, TcConvPlan* pl) {
  Geom3 g;
  g.H = p.H; g.W = p.W;
  bool any_patch = false;
  for (int s = 0; s < p.nseg; ++s) any_patch |= p.seg[s].patch != 0;
  if (any_patch) {
    EO_REQUIRE(p.H % 16 == 0 && p.W % 8 == 0, EO_ERR_ARG, "tc_conv3: halo patches need H %% 16 == 0 and W %% 8 == 0 (%dx%d)", p.H, p.W);
    g.bw = 8; g.bh = 16; g.bn = 1;
  } else {
    tc_conv_tile_geom(p.H, p.W, &g.bw, &g.bh, &g.bn);
  }
  EO_REQUIRE(p.W % g.bw == 0 && p.H % g.bh == 0, EO_ERR_ARG, "tc_conv3: feature map %dx%d is not tileable by %dx%d boxes",
             p.H, p.W, g.bh, g.bw);
  EO_REQUIRE(!(p.stats && g.bw * g.bh < 32), EO_ERR_ARG,
             "tc_conv3: fused GroupNorm statistics need at least 32 pixels per image (%dx%d)", p.H, p.W);
  EO_REQUIRE(!p.out_f32 && !p.res_f32, EO_ERR_ARG, "tc_conv3: bf16 outputs and residuals only");
  EO_REQUIRE(!p.out_nchw_C || (p.Cout == 64 && p.out_nchw_C >= 1 && p.out_nchw_C <= 64 && !p.stats && !p.out_sw), EO_ERR_ARG,
             "tc_conv3: the NCHW fp32 output takes one 64-wide channel tile, no statistics, no strided view");
  g.tiles_w = p.W / g.bw; g.tiles_h = p.H / g.bh;
  g.d_tw.set((unsigned)g.tiles_w); g.d_th.set((unsigned)g.tiles_h);
  pl->g3 = g;
  // Channel-tile width: the widest that divides Cout (fewest re-reads / re-transforms of the activation operand) unless
  // the problem is so small that it leaves SMs idle (the reference's 64 x 64 batch-1 case, the deep levels of
  // 128 x 128 batch 16): then the width that minimises waves x cost per 16-deep K step.  Costs are the MEASURED clk per
  // K step of full launches (per-CTA traces): the tensor pipe's BN / 2 for wide tiles, the
  // shared-memory port for narrow ones
  int BN = 64;
  {
    const long long mpairs = ((long long)g.tiles_w * g.tiles_h * ceil_div(p.B, g.bn) + 1) / 2;
    const long long ncl = std::max(1, num_sms() / 2);
    long long best = -1;
    for (int cand : {256, 192, 128, 64}) {
      if (p.Cout % cand != 0) continue;
      const long long waves = ceil_div(mpairs * (p.Cout / cand), ncl);
      const long long cost = waves * (cand == 256 ? 132 : cand == 192 ? 106 : cand == 128 ? 96 : 80);
      if (best < 0 || cost < best) { best = cost; BN = cand; }      // ties keep the wider tile
    }
  }
  pl->bn_tile = BN;
  std::vector<KEnt3> tab;
  int kofs = 0;
  for (int s = 0; s < p.nseg; ++s) {
    const TcConvSeg& sg = p.seg[s];
    EO_REQUIRE(sg.C % BK == 0, EO_ERR_ARG, "tc_conv3: segment channels %d must be a multiple of 64", sg.C);
    EO_REQUIRE(sg.stride == 1 || (sg.stride == 2 && !sg.patch), EO_ERR_ARG, "tc_conv3: segment stride %d", sg.stride);
    const uint64_t sW = (uint64_t)p.W * sg.stride, sH = (uint64_t)p.H * sg.stride;      // the segment's own grid
    uint64_t dims[4] = {(uint64_t)sg.C, sW, sH, (uint64_t)sg.Bt};
    uint64_t str[3] = {(uint64_t)sg.C * 2, sW * sg.C * 2, sH * sW * sg.C * 2};
    uint32_t box[4] = {(uint32_t)BK, (uint32_t)(g.bw * sg.stride), (uint32_t)(g.bh * sg.stride), (uint32_t)g.bn};
    uint32_t est[4] = {1u, (uint32_t)sg.stride, (uint32_t)sg.stride, 1u};
    EO_REQUIRE(!sg.gn_scale || sg.patch == TC_PATCH_3X3, EO_ERR_ARG, "tc_conv3: GroupNorm can only be folded into a 3x3 halo-patch segment");
    EO_REQUIRE(!sg.gn_scale || (sg.gn_shift && sg.gn_ld % 4 == 0 && sg.gn_coff % 8 == 0), EO_ERR_ARG,
               "tc_conv3: GroupNorm rows must be 16-byte aligned");
    const int pr0 = (sg.patch >> 4) & 15, pnr = (sg.patch >> 8) & 15, pc0 = (sg.patch >> 12) & 15, pnc = (sg.patch >> 16) & 15;
    if (sg.patch) {
      EO_REQUIRE(pnr >= 1 && pnc >= 1 && pr0 + pnr <= 3 && pc0 + pnc <= 3 && sg.ntaps == pnr * pnc, EO_ERR_ARG,
                 "tc_conv3: a patch segment takes a window of the 3x3 neighbourhood (code %#x, %d taps)", sg.patch, sg.ntaps);
      for (int t = 0; t < sg.ntaps; ++t)
        EO_REQUIRE(sg.dh[t] == pr0 + t / pnc - 1 && sg.dw[t] == pc0 + t % pnc - 1 && sg.dn[t] == 0, EO_ERR_ARG,
                   "tc_conv3: the taps of a patch segment walk its window row by row");
      box[1] = PATCH_W; box[2] = PATCH_H; box[3] = 1;
    }
    int rc = encode_tmap_bf16(&pl->mapA[s], sg.ptr, 4, dims, str, box, est);
    if (rc != EO_OK) return rc;
    if (sg.patch) {
      // K order of a patch segment: (64-channel block, tap, channel)
      for (int c0 = 0; c0 < sg.C; c0 += BK) {
        KEnt3 e{}; e.seg = s; e.c0 = c0; e.patch = sg.patch; e.kofs = kofs; e.sc = 1;
        e.gn = sg.gn_scale ? (sg.silu ? 2 : 1) : 0; e.gnc = sg.gn_coff + c0;
        tab.push_back(e);
        kofs += pnr * pnc * BK;
      }
    } else {
      for (int t = 0; t < sg.ntaps; ++t)
        for (int c0 = 0; c0 < sg.C; c0 += BK) {
          KEnt3 e{}; e.seg = s; e.c0 = c0; e.dhw = ((int)sg.dh[t] & 0xffff) | ((int)sg.dw[t] << 16); e.dn = sg.dn[t]; e.kofs = kofs; e.sc = sg.stride;
          e.gn = sg.gn_scale ? (sg.silu ? 2 : 1) : 0; e.gnc = sg.gn_coff + c0;
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
    if (p.out_sw) {
      EO_REQUIRE(!p.residual, EO_ERR_ARG, "tc_conv3: a strided output view takes no residual");
      str[0] = (uint64_t)p.out_sw * 2; str[1] = (uint64_t)p.out_sh * 2; str[2] = (uint64_t)p.out_sn * 2;
    }
    // output: one store per epilogue warp = the 32 pixels of its TMEM lane quadrant
    const uint32_t sw = (uint32_t)(g.bw < 32 ? g.bw : 32);
    const uint32_t sh = (uint32_t)(g.bh < (int)(32 / sw) ? g.bh : (int)(32 / sw));
    uint32_t sbox[4] = {(uint32_t)BK, sw, sh, 32 / (sw * sh)};
    // (an NCHW-output conv never stores through this map; it still needs a valid base address)
    int rc = encode_tmap_bf16(&pl->mapOut, p.out_nchw_C ? p.Wp : p.out, 4, dims, str, sbox);
    if (rc != EO_OK) return rc;
    pl->mapRes = pl->mapOut;
    if (p.residual) {
      uint32_t box[4] = {(uint32_t)BK, (uint32_t)g.bw, (uint32_t)g.bh, (uint32_t)g.bn};
      rc = encode_tmap_bf16(&pl->mapRes, p.residual, 4, dims, str, box);
      if (rc != EO_OK) return rc;
    }
  }
  // the GroupNorm-folded loads of one tile, in K order, with everything the transform warps need resolved
  std::vector<XEnt3> xt;
  for (const KEnt3& e : tab) {
    if (!e.gn) continue;
    const TcConvSeg& sg = p.seg[e.seg];
    XEnt3 x{};
    x.src = reinterpret_cast<unsigned long long>(sg.ptr) + (unsigned long long)e.c0 * 2;
    x.gsc = reinterpret_cast<unsigned long long>(sg.gn_scale + e.gnc);
    x.gsh = reinterpret_cast<unsigned long long>(sg.gn_shift + e.gnc);
    x.cs2 = (uint32_t)sg.C * 2u; x.gld = (uint32_t)sg.gn_ld; x.silu = e.gn == 2 ? 1u : 0u;
    xt.push_back(x);
  }
  EO_REQUIRE((int)xt.size() <= MAX_XENT, EO_ERR_ARG, "tc_conv3: %d GroupNorm-folded loads per tile (max %d)", (int)xt.size(), MAX_XENT);
  pl->n_xent = (int)xt.size();
  const size_t tab_bytes = tab.size() * sizeof(KEnt3), xt_off = (tab_bytes + 15) & ~(size_t)15;
  cudaError_t e = cudaMalloc(&pl->d_kblks, xt_off + std::max<size_t>(xt.size(), 1) * sizeof(XEnt3));
  if (e == cudaSuccess) e = cudaMemcpy(pl->d_kblks, tab.data(), tab_bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && !xt.empty())
    e = cudaMemcpy(reinterpret_cast<uint8_t*>(pl->d_kblks) + xt_off, xt.data(), xt.size() * sizeof(XEnt3), cudaMemcpyHostToDevice);
  pl->xent_off = xt_off;
  EO_REQUIRE(e == cudaSuccess, EO_ERR_CUDA, "tc_conv3: operand table upload failed: %s", cudaGetErrorString(e));
  return EO_OK;
}

static long long* g_trace3 = nullptr;
static int g_trace3_n = 0;
void tc_conv3_set_trace(long long* dev_buf, int n) { g_trace3 = dev_buf; g_trace3_n = n; }

int tc_conv3_launch(const TcConvPlan* pl, int B, cudaStream_t st, float* out_nchw) {
  static bool attr_set = false;
  const TcConvParams& p = pl->p;
  Geom3 g = pl->g3;
  const int BN = pl->bn_tile;
  g.d_nt.set((unsigned)(p.Cout / BN));
  const bool has_res = p.residual != nullptr;
  const int b_bytes = (BN / 2) * 128;
  // operand-A stages: all three to whoever fills them, two each when a conv has both kinds of load
  bool any_gn = false, any_raw = false;
  for (int s = 0; s < p.nseg; ++s) { if (p.seg[s].gn_scale) any_gn = true; else any_raw = true; }
  int SAR = any_raw ? (any_gn ? 2 : 3) : 0, SAG = any_gn ? (any_raw ? 2 : 3) : 0;
  // narrow tiles: three weight tiles (one kernel row of a patch) per stage, see the MMA warp
  bool any_patch = false;
  for (int s = 0; s < p.nseg; ++s) any_patch |= p.seg[s].patch != 0;
  int TPB = (any_patch && BN <= 192) ? 3 : 1;
  int b_stage = TPB * b_bytes;
  int a_stage = A_STAGE, GRP = 1;
  if (EO_CONV_DEEP_RING && !any_patch && !any_gn) {
    // plain tiles only: one weight tile per operand load, so as many (operand, weight) stage pairs as fit
    a_stage = PLAIN_BYTES;
    const int avail = SMEM_LIMIT - 1024 - Smem::VAR_OFF - (has_res ? 2 * STG_BYTES : 0);
    if (BN <= EO_CONV_GROUP_BNMAX) {        // a 256-wide tile's four MMAs (512 clk) already cover the issue chain
      for (int cand = std::min(EO_CONV_GROUP_MAX, pl->nkb); cand > 1; --cand)
        if (EO_CONV_GROUP_STAGES * cand * (PLAIN_BYTES + b_bytes) <= avail) { GRP = cand; break; }
    }
    if (GRP > 1) { TPB = GRP; a_stage = GRP * PLAIN_BYTES; b_stage = GRP * b_bytes; }
    SAR = std::max(GRP > 1 ? 2 : 3, std::min(SA_MAX, avail / (a_stage + b_stage)));
  }
  const int fixed = (SAR + SAG) * a_stage + Smem::VAR_OFF + (has_res ? 2 * STG_BYTES : 0);
  int SB = (SMEM_LIMIT - 1024 - fixed) / b_stage;
  if (SB > MAX_SB) SB = MAX_SB;
  EO_REQUIRE(SB >= 2, EO_ERR_STATE, "tc_conv3: shared memory budget leaves %d weight stages", SB);
  const int dyn = fixed + SB * b_stage + 1024;
  if (!attr_set) {
    EO_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tc3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
#ifdef EO_DEVTOOLS
    EO_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tc3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
#endif
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
  EO_REQUIRE((p.out_nchw_C > 0) == (out_nchw != nullptr), EO_ERR_ARG, "tc_conv3: NCHW output pointer / plan mismatch");
  ep.out_nchw = out_nchw; ep.out_nchw_C = p.out_nchw_C;
  for (int s = 0; s < 3; ++s) {
    const bool on = s < p.nseg && p.seg[s].gn_scale != nullptr;
    ep.gn_scale[s] = on ? p.seg[s].gn_scale : nullptr;
    ep.gn_shift[s] = on ? p.seg[s].gn_shift : nullptr;
    ep.gn_ld[s] = on ? p.seg[s].gn_ld : 0;
    ep.gn_src[s] = on ? p.seg[s].ptr : nullptr;
    ep.gn_C[s] = on ? p.seg[s].C : 0;
    ep.any_gn |= on ? 1 : 0;
  }
#ifdef EO_DEVTOOLS
  ep.trace = g_trace3; ep.trace_n = g_trace3_n;
  ep.trace_ext = env_flag("EO_TRACE_EXT", false) ? 1 : 0;
#endif
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * ncl));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = (size_t)dyn;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_allowed() ? 2 : 1;
#ifdef EO_DEVTOOLS
  auto kern = g_trace3 ? k_conv_tc3<true> : k_conv_tc3<false>;
#else
  auto kern = k_conv_tc3<false>;
#endif
  EO_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, pl->mapA[0], pl->mapA[1], pl->mapA[2], pl->mapB, pl->mapOut,
                                   pl->mapRes, (const KEnt3*)pl->d_kblks, pl->nkb,
                                   (const XEnt3*)(reinterpret_cast<const uint8_t*>(pl->d_kblks) + pl->xent_off), pl->n_xent, g, B,
                                   BN, SB, TPB, SAR, SAG, a_stage, GRP, n_work, n_ntiles, ep));
  return EO_OK;
}

}  // namespace eo
