// Plan object shared by the two tensor-core convolution kernels (tc_conv.cu: one tile per CTA,
// kept for A/B comparison behind EO_CONV_V2=1; tc_conv3.cu: persistent, the default).
#pragma once
#include <cuda.h>
#include "kernels.h"

namespace eo {

// ---- tc_conv.cu
struct KBlk { int seg; int c0; int dh_dw; int dn; };   // dh: low 16 bits, dw: high 16 bits
struct TileGeom {
  int bw, bh, bn;          // box extents, bw*bh*bn == 128
  int tiles_w, tiles_h;
  int H, W;
};

// ---- tc_conv3.cu
// One operand-A load of the K loop: a plain 128-pixel tile shifted by (dh, dw) in image plane dn
// (one weight tile follows), or a halo patch (nine weight tiles follow, one per tap).
struct KEnt3 { int seg, c0, dh, dw, dn, kofs, patch, pad; };
struct Geom3 {
  int bw, bh, bn;
  int tiles_w, tiles_h;
  int H, W;
};
struct Epi3 {
  const float* bias;        // [Cout] or null
  const float* bias_nc;     // [B, ld_bias_nc] or null
  int ld_bias_nc;
  double* stats;            // [B, Cout, 2] (sum, sum of squares) accumulated with atomics, or null
  int Cout, has_res;
  long long* trace;         // development aid: [n][8] per-CTA counters, or null
  int trace_n;
};

struct TcConvPlan {
  CUtensorMap mapA[3];
  CUtensorMap mapB;
  CUtensorMap mapOut, mapRes;   // v3: TMA store of the output, TMA load of the residual
  void* d_kblks = nullptr;
  int nkb = 0;
  TileGeom g{};
  Geom3 g3{};
  int bn_tile = 128;
  bool pair = true;
  bool v3 = false;
  TcConvParams p;
};

bool tc_conv3_enabled();
int tc_conv3_plan_fill(const TcConvParams& p, TcConvPlan* pl);
int tc_conv3_launch(const TcConvPlan* pl, int B, cudaStream_t st);
void tc_conv3_set_trace(long long* dev_buf, int n_ctas);

}  // namespace eo
