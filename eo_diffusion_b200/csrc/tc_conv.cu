// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05.mma, accumulators
// in TMEM, operands staged by TMA).  bf16 operands, fp32 accumulate.
//
// Replaces, for the bf16 mode, every nn.Conv2d 3x3 / 1x1 and nn.Conv1d k=1 on the UNet
// path (reference backbones/unet_openai.py: ResBlock :316,:342,:353; AttentionBlock qkv/
// proj_out :412,:422; Downsample :262; Upsample :227), with the adds that follow them
// (`h + emb_out` :382, `skip_connection(x) + h` :385, `x + h` :433) in the epilogue, the
// 1x1 skip convolution as extra K blocks of the same accumulator, and th.cat (:773) as a
// second K segment.
//
// GEMM view:  D[M = pixels, N = Cout] = sum_k A[M, k] * Wp[N, k],  k = (segment, tap, channel)
//   A tile  : 128 output pixels = a (bn x bh x bw) box of the NHWC activation; for tap
//             (dh, dw) the SAME box shifted by (dh, dw) is fetched by one 4-D tiled TMA
//             load; out-of-bounds rows/columns are zero-filled by the TMA unit, which is
//             exactly the conv's zero padding.  64 channels (128 B) per K block, 128B swizzle.
//   B tile  : BN x 64 slice of the packed weights Wp[Cout][Ktot] (K contiguous), 2-D TMA.
//   D       : 128 lanes x BN fp32 columns of TMEM.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one
// elected lane), warps 2..5 = epilogue (each owns the TMEM lane quadrant warp_id % 4).
// Two CTAs are co-resident per SM (<= 96 KB smem and <= 256 TMEM columns each) so that one
// CTA's epilogue overlaps the other's main loop.
//
// Roofline: tensor pipe.  Algorithmic FLOPs per launch = 2 * M * Cout * Ktot.
#include "kernels.h"
#include "tc_common.cuh"
#include "tc_conv_epi.cuh"
#include "tc_conv_plan.h"
#include <cstdlib>
#include <vector>

namespace eo {

namespace {

using tc::Epi;

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;   // 16 KB

// BN = N extent of the accumulator tile.  PAIR: two CTAs (a cluster of 2 on one TPC) compute a
// 256 x BN tile with cta_group::2 MMAs; each CTA stages its own 128 pixel rows of A and HALF of
// the BN weight rows, so the shared-memory traffic per MMA drops by a third to a half.
template <int BN, int STAGES, bool PAIR>
struct SmemLayout {
  static constexpr int B_ROWS = PAIR ? BN / 2 : BN;
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = STAGES * A_BYTES;
  static constexpr int TAB_OFF = B_OFF + STAGES * B_BYTES;
  static constexpr int MAX_KB = 192;
  static constexpr int STAT_OFF = TAB_OFF + MAX_KB * 16;          // [4 warps][BN][2] floats
  static constexpr int BAR_OFF = STAT_OFF + 4 * BN * 2 * 4;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;   // slack for manual 1024 B alignment
};

template <int BN, int STAGES, bool PAIR>
__global__ void __launch_bounds__(192, 2)
k_conv_tc(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
          const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB,
          const KBlk* __restrict__ kblks, int nkb, TileGeom g, int B, Epi ep) {
  using L = SmemLayout<BN, STAGES, PAIR>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (tc::smem_u32(smem_raw) & 1023)) & 1023);
  KBlk* tab = reinterpret_cast<KBlk*>(smem + L::TAB_OFF);
  float* sstat = reinterpret_cast<float*>(smem + L::STAT_OFF);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? tc::cluster_ctarank() : 0u;
  if (threadIdx.x == 0) { tc::trace_stamp(ep, 0); tc::trace_stamp(ep, 7); tc::trace_stamp(ep, 1); }

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&mapA0);
    tc::tma_prefetch_desc(&mapB);
    // full: one arrival (the leader's producer, which also posts the byte count of BOTH CTAs' loads;
    // the peer's TMA only completes transactions on it) -- only the leader's copy is used
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(tmem_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) { tc::tmem_alloc2(tmem_ptr, BN); tc::tmem_relinquish2(); }
    else { tc::tmem_alloc(tmem_ptr, BN); tc::tmem_relinquish(); }
  }
  for (int i = threadIdx.x; i < nkb; i += blockDim.x) tab[i] = kblks[i];
  tc::tc_fence_before();
  if (PAIR) tc::cluster_sync_all(); else __syncthreads();   // peer barriers are initialised past here
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) tc::trace_stamp(ep, 2);

  // tile coordinates (the two CTAs of a pair take consecutive pixel tiles)
  const int mt = blockIdx.x;
  const int tw = mt % g.tiles_w;
  const int th = (mt / g.tiles_w) % g.tiles_h;
  const int nt = mt / (g.tiles_w * g.tiles_h);
  const int w0 = tw * g.bw, h0 = th * g.bh, n0 = nt * g.bn;
  const int nbase = blockIdx.y * BN;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t full0 = tc::smem_u32(&full[0]);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        tc::mbar_wait(&empty[s], ph ^ 1);
        const KBlk e = tab[kb];
        const int dh = (int)(short)(e.dh_dw & 0xffff), dw = (int)(short)(e.dh_dw >> 16);
        const CUtensorMap* ma = e.seg == 0 ? &mapA0 : (e.seg == 1 ? &mapA1 : &mapA2);
        if (PAIR) {
          const uint32_t lbar = tc::mapa_u32(full0 + s * 8, 0);       // the leader's full[s]
          // The peer's bytes for this phase can only be issued after the leader's MMA released the
          // slot (multicast commit), i.e. after the previous phase completed; if they land before the
          // leader arms the phase the transaction count just goes negative until expect_tx.
          if (rank == 0) tc::mbar_arrive_expect_tx(&full[s], 2 * (A_BYTES + L::B_BYTES));
          tc::tma2_load_4d(smem + L::A_OFF + s * A_BYTES, ma, lbar, e.c0, w0 + dw, h0 + dh, n0 + e.dn);
          tc::tma2_load_2d(smem + L::B_OFF + s * L::B_BYTES, &mapB, lbar, kb * BK, nbase + (int)rank * L::B_ROWS);
        } else {
          tc::mbar_arrive_expect_tx(&full[s], A_BYTES + L::B_BYTES);
          tc::tma_load_4d(smem + L::A_OFF + s * A_BYTES, ma, &full[s], e.c0, w0 + dw, h0 + dh, n0 + e.dn);
          tc::tma_load_2d(smem + L::B_OFF + s * L::B_BYTES, &mapB, &full[s], kb * BK, nbase);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = tc::make_idesc_bf16(PAIR ? 2 * BM : BM, BN, 0, 0);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        tc::mbar_wait(&full[s], ph);
        tc::tc_fence_after();
        if (kb == 0) tc::trace_stamp(ep, 3);
        const uint64_t adesc = tc::make_sw128_desc(tc::smem_u32(smem + L::A_OFF + s * A_BYTES));
        const uint64_t bdesc = tc::make_sw128_desc(tc::smem_u32(smem + L::B_OFF + s * L::B_BYTES));
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          if (PAIR) tc::umma2_f16_ss(tmem_base, tc::desc_advance(adesc, k * 32), tc::desc_advance(bdesc, k * 32),
                                     idesc, (kb | k) != 0 ? 1u : 0u);
          else tc::umma_f16_ss(tmem_base, tc::desc_advance(adesc, k * 32), tc::desc_advance(bdesc, k * 32),
                               idesc, (kb | k) != 0 ? 1u : 0u);
        }
        // frees the smem stage (in both CTAs) when these MMAs retire
        if (PAIR) tc::umma2_commit_mc(&empty[s], 3); else tc::umma_commit(&empty[s]);
      }
      if (PAIR) tc::umma2_commit_mc(tmem_full, 3); else tc::umma_commit(tmem_full);   // accumulator complete
    }
  } else {
    // ---- epilogue: TMEM -> registers -> (+bias, +per-sample bias, +residual) -> HBM [+ GN partial sums]
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ww = row % g.bw;
    const int hh = (row / g.bw) % g.bh;
    const int nn = row / (g.bw * g.bh);
    const int n_img = n0 + nn;
    const bool valid = n_img < B;
    const long long pix = ((long long)n_img * g.H + (h0 + hh)) * g.W + (w0 + ww);
    tc::mbar_wait(tmem_full, 0);
    tc::tc_fence_after();
    if (threadIdx.x == 64) tc::trace_stamp(ep, 4);
    // statistics need a warp's 32 rows inside one image (host guarantees bw*bh >= 32 when stats != null)
    tc::conv_epilogue<BN>(tmem_base, q, lane, threadIdx.x - 64, nbase, valid, n_img, pix, ep, g.bn == 1, sstat);
    if (threadIdx.x == 64) tc::trace_stamp(ep, 5);
    tc::tc_fence_before();
  }
  if (PAIR) tc::cluster_sync_all(); else __syncthreads();   // nobody leaves while the peer still uses its smem/TMEM
  if (warp == 1) {
    tc::tc_fence_after();
    if (PAIR) tc::tmem_dealloc2(tmem_base, BN); else tc::tmem_dealloc(tmem_base, BN);
  }
  if (threadIdx.x == 0) tc::trace_stamp(ep, 6);
}

}  // namespace

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (PFN_encodeTiled)p;
  }
  return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
  PFN_encodeTiled fn = get_encode_fn();
  EO_REQUIRE(fn != nullptr, EO_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t d[5]; cuuint64_t s[4]; cuuint32_t b[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                  d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EO_REQUIRE(r == CUDA_SUCCESS, EO_ERR_CUDA,
             "cuTensorMapEncodeTiled failed (CUresult %d; rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)",
             (int)r, rank, (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0),
             (unsigned long long)(rank > 2 ? d[2] : 0), (unsigned long long)(rank > 3 ? d[3] : 0),
             b[0], rank > 1 ? b[1] : 0, rank > 2 ? b[2] : 0, rank > 3 ? b[3] : 0);
  return EO_OK;
}

static long long* g_trace = nullptr;
static int g_trace_n = 0;
void tc_conv_set_trace(long long* dev_buf, int n_ctas) { g_trace = dev_buf; g_trace_n = n_ctas; tc_conv3_set_trace(dev_buf, n_ctas); }

static int floor_pow2(int v) { int r = 1; while (r * 2 <= v) r *= 2; return r; }

static bool use_pairs() {
  static int v = -1;
  if (v < 0) { const char* e = std::getenv("EO_CONV_1CTA"); v = (e && e[0] == '1') ? 0 : 1; }
  return v != 0;
}

// 128-pixel tile of a plain (non-patch) operand: bw x bh pixels of bn images, every extent a power of two that
// divides the feature map (24 x 24 -> 8 x 8 x 2, 12 x 12 -> 4 x 4 x 8, 48 x 32 -> 16 x 8 x 1)
void tc_conv_tile_geom(int H, int W, int* bw, int* bh, int* bn) {
  int w = W & -W;                       // largest power of two dividing W
  if (w > 16) w = 16;
  int h = H & -H;
  if (h > BM / w) h = BM / w;
  *bw = w; *bh = h; *bn = BM / (w * h);
}

bool tc_conv_stats_supported(int H, int W) {
  if (tc_conv3_enabled()) {
    int bw, bh, bn;
    tc_conv_tile_geom(H, W, &bw, &bh, &bn);
    return bw * bh >= 32;
  }
  int bw = floor_pow2(W < 16 ? W : 16);
  int bh = floor_pow2(H < BM / bw ? H : BM / bw);
  return bw * bh >= 32;
}

int tc_conv_plan_create(const TcConvParams& p, TcConvPlan** out) {
  EO_REQUIRE(p.nseg >= 1 && p.nseg <= 3, EO_ERR_ARG, "tc_conv: nseg");
  EO_REQUIRE(p.Cout % 64 == 0, EO_ERR_ARG, "tc_conv: Cout %d must be a multiple of 64", p.Cout);
  TcConvPlan* pl = new TcConvPlan();
  pl->p = p;
  if (tc_conv3_enabled()) {
    int rc3 = tc_conv3_plan_fill(p, pl);
    if (rc3 != EO_OK) { tc_conv_plan_destroy(pl); return rc3; }
    *out = pl;
    return EO_OK;
  }
  for (int s = 0; s < p.nseg; ++s)
    if (p.seg[s].patch || p.seg[s].gn_scale || p.out_sw) {
      delete pl;
      set_error("tc_conv: halo patches and folded GroupNorm need the persistent kernel");
      return EO_ERR_ARG;
    }
  pl->pair = use_pairs();
  // ---- tile geometry: 128 pixels = bn x bh x bw
  TileGeom g;
  g.H = p.H; g.W = p.W;
  g.bw = floor_pow2(p.W < 16 ? p.W : 16);
  g.bh = floor_pow2(p.H < BM / g.bw ? p.H : BM / g.bw);
  g.bn = BM / (g.bw * g.bh);
  if (p.W % g.bw != 0 || p.H % g.bh != 0) {
    delete pl;
    set_error("tc_conv: feature map %dx%d is not tileable by %dx%d boxes", p.H, p.W, g.bh, g.bw);
    return EO_ERR_ARG;
  }
  if (p.stats && g.bw * g.bh < 32) {
    delete pl;
    set_error("tc_conv: fused GroupNorm statistics need at least 32 pixels per image (%dx%d)", p.H, p.W);
    return EO_ERR_ARG;
  }
  g.tiles_w = p.W / g.bw; g.tiles_h = p.H / g.bh;
  pl->g = g;
  pl->bn_tile = (p.Cout % 256 == 0) ? 256 : 128;
  // ---- K-block table and activation maps
  std::vector<KBlk> tab;
  for (int s = 0; s < p.nseg; ++s) {
    const TcConvSeg& sg = p.seg[s];
    if (sg.C % BK != 0) {
      delete pl;
      set_error("tc_conv: segment channels %d must be a multiple of 64", sg.C);
      return EO_ERR_ARG;
    }
    uint64_t dims[4] = {(uint64_t)sg.C, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)sg.Bt};
    uint64_t str[3] = {(uint64_t)sg.C * 2, (uint64_t)p.W * sg.C * 2, (uint64_t)p.H * p.W * sg.C * 2};
    uint32_t box[4] = {(uint32_t)BK, (uint32_t)g.bw, (uint32_t)g.bh, (uint32_t)g.bn};
    int rc = encode_tmap_bf16(&pl->mapA[s], sg.ptr, 4, dims, str, box);
    if (rc != EO_OK) { delete pl; return rc; }
    for (int t = 0; t < sg.ntaps; ++t)
      for (int c0 = 0; c0 < sg.C; c0 += BK) {
        KBlk e;
        e.seg = s; e.c0 = c0;
        e.dh_dw = ((int)sg.dh[t] & 0xffff) | ((int)sg.dw[t] << 16);
        e.dn = sg.dn[t];
        tab.push_back(e);
      }
  }
  for (int s = p.nseg; s < 3; ++s) pl->mapA[s] = pl->mapA[0];
  pl->nkb = (int)tab.size();
  if (pl->nkb * BK != p.Ktot || pl->nkb > 192) {
    delete pl;
    set_error("tc_conv: K blocks %d inconsistent with Ktot %d (max 192 blocks)", pl->nkb, p.Ktot);
    return EO_ERR_ARG;
  }
  {
    uint64_t dims[2] = {(uint64_t)p.Ktot, (uint64_t)p.Cout};
    uint64_t str[1] = {(uint64_t)p.Ktot * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)(pl->pair ? pl->bn_tile / 2 : pl->bn_tile)};
    int rc = encode_tmap_bf16(&pl->mapB, p.Wp, 2, dims, str, box);
    if (rc != EO_OK) { delete pl; return rc; }
  }
  cudaError_t e = cudaMalloc(&pl->d_kblks, tab.size() * sizeof(KBlk));
  if (e == cudaSuccess)
    e = cudaMemcpy(pl->d_kblks, tab.data(), tab.size() * sizeof(KBlk), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("tc_conv: K-block table upload failed: %s", cudaGetErrorString(e));
    tc_conv_plan_destroy(pl);
    return EO_ERR_CUDA;
  }
  *out = pl;
  return EO_OK;
}

void tc_conv_plan_destroy(TcConvPlan* p) {
  if (!p) return;
  if (p->d_kblks) cudaFree(p->d_kblks);
  delete p;
}

template <int BN, int STAGES, bool PAIR>
static int launch_tc(const TcConvPlan* pl, int B, cudaStream_t st) {
  using L = SmemLayout<BN, STAGES, PAIR>;
  static bool attr_set = false;
  auto kern = k_conv_tc<BN, STAGES, PAIR>;
  if (!attr_set) {
    EO_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES));
    attr_set = true;
  }
  const TileGeom& g = pl->g;
  const TcConvParams& p = pl->p;
  int tiles_n = (int)ceil_div(B, g.bn);
  int mtiles = g.tiles_w * g.tiles_h * tiles_n;
  if (PAIR) mtiles = (mtiles + 1) & ~1;      // an odd last tile gets an all-masked partner
  Epi ep;
  ep.bias = p.bias; ep.bias_nc = p.bias_nc; ep.ld_bias_nc = p.ld_bias_nc; ep.residual = p.residual;
  ep.out = p.out; ep.stats = p.stats; ep.Cout = p.Cout; ep.res_f32 = p.res_f32; ep.out_f32 = p.out_f32;
  ep.trace = g_trace; ep.trace_n = g_trace_n;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)mtiles, (unsigned)ceil_div(p.Cout, BN));
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = L::DYN_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  EO_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, pl->mapA[0], pl->mapA[1], pl->mapA[2], pl->mapB,
                                   (const KBlk*)pl->d_kblks, pl->nkb, g, B, ep));
  return EO_OK;
}

int tc_conv_launch(const TcConvPlan* pl, int B, cudaStream_t st) {
  if (pl->v3) return tc_conv3_launch(pl, B, st);
  if (pl->pair) {
    // per CTA and stage: 16 KB of A + BN/2 weight rows (16 or 8 KB); two CTAs stay co-resident per SM
    if (pl->bn_tile == 256) return launch_tc<256, 3, true>(pl, B, st);
    return launch_tc<128, 4, true>(pl, B, st);
  }
  if (pl->bn_tile == 256) return launch_tc<256, 2, false>(pl, B, st);
  return launch_tc<128, 3, false>(pl, B, st);
}

}  // namespace eo
