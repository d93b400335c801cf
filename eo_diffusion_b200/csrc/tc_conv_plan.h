// Plan object of the tensor-core convolution (tc_conv3.cu; host helpers in tc_conv.cu).
#pragma once
#include <cuda.h>
#include "kernels.h"

namespace eo {

// One operand-A load of the K loop: a plain 128-pixel tile shifted by (dh, dw) in image plane dn
// (one weight tile follows), or a halo patch (nine weight tiles follow, one per tap).
//   gn: 0 = operand loaded by TMA as is, 1 = GroupNorm affine folded in (x*scale + shift), 2 = affine + SiLU:
//   the transform warps load the patch from global memory, apply it and write the operand stage;
//   gnc = first channel of this load in the scale/shift rows;
//   sc = 1, or 2 for a stride-2 source (the tile origin is doubled: the segment's tensor map walks every second pixel)
struct KEnt3 { int seg, c0, dhw, dn, kofs, patch, gn, gnc, sc; };   // dhw: dh in the low 16 bits, dw in the high
// One GroupNorm-folded patch load with its pointers resolved (what the transform warps need, 48 bytes = three 16-byte
// shared-memory reads made a whole patch ahead): src = the segment's activation at channel c0 (bytes), gsc / gsh = its
// scale / shift rows at that channel (row of image n: + n * gld floats), cs2 = pixel pitch in bytes
struct XEnt3 { unsigned long long src, gsc, gsh; unsigned cs2, gld, silu, pad0; unsigned long long pad1; };
static_assert(sizeof(XEnt3) == 48, "XEnt3 is read as three uint4");
// division by a launch-time constant as multiply-high + shift (exact for dividends below 2^31): the generic
// integer division sequence is ~35 dependent instructions, and every warp role decodes a tile index per tile
struct FastDiv {
  unsigned d = 1, mul = 0, shr = 0;
  void set(unsigned div) {
    d = div; mul = 0; shr = 0;
    if (div > 1) {
      unsigned lg = 0;
      while ((1u << lg) < div) ++lg;
      const unsigned p = 31 + lg;
      mul = (unsigned)(((1ull << p) + div - 1) / div);
      shr = p - 32;
    }
  }
};
struct Geom3 {
  int bw, bh, bn;
  int tiles_w, tiles_h;
  int H, W;
  FastDiv d_nt, d_tw, d_th;   // by n_ntiles, tiles_w, tiles_h
};
struct Epi3 {
  const float* bias;        // [Cout] or null
  const float* bias_nc;     // [B, ld_bias_nc] or null
  int ld_bias_nc;
  double* stats;            // [B, Cout, 2] (sum, sum of squares) accumulated with atomics, or null
  int Cout, has_res;
  float* out_nchw;          // head convolution: the first out_nchw_C channels go straight to this NCHW fp32 tensor
  int out_nchw_C;           // (the network output, unet_openai.py:780) instead of the bf16 NHWC staging / TMA store
  const float* gn_scale[3];   // per segment: [B, gn_ld] rows of the GroupNorm it reads through, or null
  const float* gn_shift[3];
  int gn_ld[3];
  const void* gn_src[3];      // the segment's bf16 NHWC activation (the transform warps load it themselves)
  int gn_C[3];
  int any_gn;
  long long* trace;         // development aid: [n][8] per-CTA counters, or null
  int trace_n;
  int trace_ext;           // development aid: the trace buffer holds 2 * trace_n * 8 counters (slots 8..15 in use)
};

struct TcConvPlan {
  CUtensorMap mapA[3];
  CUtensorMap mapB;
  CUtensorMap mapOut, mapRes;   // v3: TMA store of the output, TMA load of the residual
  void* d_kblks = nullptr;      // KEnt3[nkb], then (at xent_off bytes) XEnt3[n_xent]
  int nkb = 0;
  int n_xent = 0;
  size_t xent_off = 0;
  Geom3 g3{};
  int bn_tile = 128;
  TcConvParams p;
};

void tc_conv_tile_geom(int H, int W, int* bw, int* bh, int* bn);
int tc_conv3_plan_fill(const TcConvParams& p, TcConvPlan* pl);
int tc_conv3_launch(const TcConvPlan* pl, int B, cudaStream_t st, float* out_nchw);
void tc_conv3_set_trace(long long* dev_buf, int n_ctas);

}  // namespace eo
