// Host side shared by the tensor-core kernels: the TMA tensor-map encoder, tile geometry and the
// plan object of the implicit-GEMM convolution (the kernel itself is tc_conv3.cu).
//
// The convolution replaces, for the bf16 mode, every nn.Conv2d 3x3 / 1x1 and nn.Conv1d k=1 on the UNet
// path (reference backbones/unet_openai.py: ResBlock :316,:342,:353; AttentionBlock qkv / proj_out
// :412,:422; Downsample :262; Upsample :227), with the adds that follow them (`h + emb_out` :382,
// `skip_connection(x) + h` :385, `x + h` :433) in the epilogue, the 1x1 skip convolution as extra K
// blocks of the same accumulator, and th.cat (:773) as a second K segment.
//
// GEMM view:  D[M = pixels, N = Cout] = sum_k A[M, k] * Wp[N, k],  k = (segment, tap, channel)
#include "kernels.h"
#include "tc_common.cuh"
#include "tc_conv_plan.h"

namespace eo {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda needed)
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
                     const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  PFN_encodeTiled fn = get_encode_fn();
  EO_REQUIRE(fn != nullptr, EO_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t d[5]; cuuint64_t s[4]; cuuint32_t b[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
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

void tc_conv_set_trace(long long* dev_buf, int n_ctas) { tc_conv3_set_trace(dev_buf, n_ctas); }

// 128-pixel tile of a plain (non-patch) operand: bw x bh pixels of bn images, every extent a power of two that
// divides the feature map (24 x 24 -> 8 x 8 x 2, 12 x 12 -> 4 x 4 x 8, 48 x 32 -> 16 x 8 x 1)
void tc_conv_tile_geom(int H, int W, int* bw, int* bh, int* bn) {
  int w = W & -W;                       // largest power of two dividing W
  if (w > 16) w = 16;
  int h = H & -H;
  if (h > 128 / w) h = 128 / w;
  *bw = w; *bh = h; *bn = 128 / (w * h);
}

bool tc_conv_stats_supported(int H, int W) {
  int bw, bh, bn;
  tc_conv_tile_geom(H, W, &bw, &bh, &bn);
  return bw * bh >= 32;
}

bool tc_conv_patch_supported(int H, int W) { return H % 16 == 0 && W % 8 == 0; }

int tc_conv_plan_create(const TcConvParams& p, TcConvPlan** out) {
  EO_REQUIRE(p.nseg >= 1 && p.nseg <= 3, EO_ERR_ARG, "tc_conv: nseg");
  EO_REQUIRE(p.Cout % 64 == 0, EO_ERR_ARG, "tc_conv: Cout %d must be a multiple of 64", p.Cout);
  TcConvPlan* pl = new TcConvPlan();
  pl->p = p;
  int rc = tc_conv3_plan_fill(p, pl);
  if (rc != EO_OK) { tc_conv_plan_destroy(pl); return rc; }
  *out = pl;
  return EO_OK;
}

void tc_conv_plan_destroy(TcConvPlan* p) {
  if (!p) return;
  if (p->d_kblks) cudaFree(p->d_kblks);
  delete p;
}

int tc_conv_launch(const TcConvPlan* pl, int B, cudaStream_t st, float* out_nchw) {
  return tc_conv3_launch(pl, B, st, out_nchw);
}

}  // namespace eo
