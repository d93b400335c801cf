// Host-side launch interface of every kernel in libeo_b200 (one translation unit per
// kernel family).  All launches are asynchronous on the given stream and return EO_OK or
// a negative error code with eo::set_error() filled in.
#pragma once
#include "common.cuh"

namespace eo {

enum DType { DT_F32 = 0, DT_BF16 = 1 };
inline size_t dtype_size(int dt) { return dt == DT_F32 ? 4 : 2; }

// ---------------------------------------------------------------------------------------
// generic SIMT implicit-GEMM convolution (fp32 FFMA).  Used for every conv of the fp32
// parity mode, and for the stem conv of the bf16 mode.
// ---------------------------------------------------------------------------------------
struct ConvSrc {
  const void* ptr = nullptr;   // NHWC activation of dtype `dt`, or NCHW fp32 if nchw
  int C = 0;                   // channels of this source
  int ksize = 3;               // 1 or 3 (pad = ksize/2)
  int dt = DT_F32;
  int nchw = 0;
  const float* gn_scale = nullptr;  // [B, C] (GroupNorm folded to x*scale+shift) or null
  const float* gn_shift = nullptr;
  int gn_ld = 0;               // row pitch of gn_scale / gn_shift (total channels of the GroupNorm)
  int silu = 0;                // SiLU after the affine
  int w_off = 0;               // first row (k index) of this source in the packed weights
};

struct ConvSimtParams {
  ConvSrc src[3];
  int nsrc = 0;
  int B = 0, Hin = 0, Win = 0;   // source spatial size (before `up`)
  int Hout = 0, Wout = 0;
  int stride = 1;                // 1 or 2 (3x3 sources only)
  int up = 0;                    // nearest x2 upsample of the sources before the conv
  const float* W = nullptr;      // packed [Ktot][Cout] fp32
  int Ktot = 0, Cout = 0;
  const float* bias = nullptr;     // [Cout] or null
  const float* bias_nc = nullptr;  // [B, ld_bias_nc] per-sample vector or null
  int ld_bias_nc = 0;
  const void* residual = nullptr;  // NHWC [B,Hout,Wout,Cout] dtype out_dt, or null
  void* out = nullptr;
  int out_dt = DT_F32;
  int out_nchw = 0;              // write NCHW fp32 (final eps)
};
int launch_conv_simt(const ConvSimtParams& p, cudaStream_t st);

// stem conv: conv3x3 over cat(x, cond), both NCHW fp32 with few channels; Wp packed [K][Cout] fp32
// with K ordered (x: tap, c), (cond: tap, c); NHWC output of dtype out_dt
int launch_conv_stem(const float* x, int Cx, const float* cond, int Cc, const float* Wp, const float* bias,
                     void* out, int out_dt, int B, int H, int W, int Cout, cudaStream_t st);

// out conv of the bf16 mode: GN+SiLU folded on load, tiny Cout (<=16), NHWC in (any dtype),
// NCHW fp32 out.  Bandwidth kernel (arithmetic intensity ~26 flop/B).
struct ConvSmallNParams {
  const void* x = nullptr; int dt = DT_BF16;
  const float* gn_scale = nullptr; const float* gn_shift = nullptr;  // [B, C]
  int B = 0, H = 0, W = 0, C = 0, Cout = 0;
  const float* Wp = nullptr;   // packed [9*C][Cout] fp32
  const float* bias = nullptr; // [Cout]
  float* out_nchw = nullptr;
};
int launch_conv_small_n(const ConvSmallNParams& p, cudaStream_t st);

// ---------------------------------------------------------------------------------------
// GroupNorm(32 groups, eps 1e-5) over one or two NHWC sources (channel concat)
// ---------------------------------------------------------------------------------------
struct GnSrc { const void* ptr = nullptr; int C = 0; };
// sums[B][32][2] (double; sum, sum of squares) must be zero on entry
int launch_gn_stats(const GnSrc* src, int nsrc, int dt, int B, int HW, double* sums,
                    cudaStream_t st);
// scale[b,c] = rstd*gamma[c]; shift[b,c] = beta[c] - mean*rstd*gamma[c]   (C = total channels)
int launch_gn_finalize(const double* sums, const float* gamma, const float* beta, int B, int C,
                       int HW, float* scale, float* shift, cudaStream_t st);
// per-channel (sum, sum of squares) of a bf16 NHWC tensor, added into stats[b][c][2]
int launch_chan_stats(const void* x_bf16, int B, int HW, int C, double* stats, cudaStream_t st);
// scale/shift from the per-channel sums of one or two channel-concatenated tensors
int launch_gn_finalize_ch(const double* sa, int Ca, const double* sb, int Cb, const float* gamma, const float* beta,
                          int B, int HW, float* scale, float* shift, cudaStream_t st);
// dst[b,p,off_s + c] = act(src_s[b,p,c]*scale[b,off_s+c] + shift[b,off_s+c])   (bf16 -> bf16)
int launch_gn_apply(const GnSrc* src, int nsrc, int B, int HW, const float* scale,
                    const float* shift, int silu, void* dst, cudaStream_t st);

// ---------------------------------------------------------------------------------------
// timestep embedding + small linears (always fp32)
// ---------------------------------------------------------------------------------------
// out[b, j] = cos(t_b * freqs[j]) (j < half), sin(...) (half <= j < 2*half)
int launch_sinusoid(const int64_t* t, const float* freqs, int B, int half, float* out,
                    cudaStream_t st);
// out[b,n] = bias[n] + (bias2 ? bias2[n] : 0) + (emb_rows ? emb_rows[idx[b]*N + n] : 0)
//            + sum_k W[n,k] * act(in[b,k])          (W row-major [N][K], as nn.Linear)
int launch_linear(const float* in, const float* W, const float* bias, const float* bias2,
                  const float* emb_rows, const int64_t* idx, int silu_in, int B, int K, int N,
                  float* out, cudaStream_t st);

// out[b, :] = table[t[b], :] (rows of ld floats, ld % 4 == 0); NaN rows for t[b] outside [0, n)
int launch_gather_rows(const float* table, int n, int ld, const int64_t* t, int B, float* out, cudaStream_t st);
// out[i] = i
int launch_iota64(int64_t* out, int n, cudaStream_t st);

// ---------------------------------------------------------------------------------------
// attention, fp32 SIMT (parity mode).  qkv is [B,T,ld] with q/k/v of head h at channel
// offsets h*head_stride + {0,1,2}*part_stride.  out is [B,T,heads*ch].
// ---------------------------------------------------------------------------------------
int launch_attention_simt(const float* qkv, float* out, int B, int T, int heads, int ch,
                          int ld, int head_stride, int part_stride, cudaStream_t st);
// head dimension 65 .. 1024 (the reference's own scripts: num_heads = 1), qkv / out of dtype dt (same layout
// arguments): one warp per query row, fp32 arithmetic.  launch_attention_simt forwards to it for ch > 64.
int launch_attention_wide(const void* qkv, void* out, int dt, int B, int T, int heads, int ch, int ld,
                          int head_stride, int part_stride, cudaStream_t st);

// ---------------------------------------------------------------------------------------
// layout pre-passes of the bf16 mode (bandwidth kernels, 16-byte vectors)
// ---------------------------------------------------------------------------------------
// dst[b,2h+i,2w+j,c] = src[b,h,w,c]
int launch_upsample2x(const void* src, void* dst, int B, int H, int W, int C, cudaStream_t st);
// dst[b, ho, wo, coff + c] = resample(act(src * scale[b, gc + c] + shift[b, gc + c])): mode 0 same grid, 1 nearest x2,
// 2 average of 2x2; scale/shift may be null; src/dst NHWC of dtype dt (up/down ResBlocks, unet_openai.py:366-371)
int launch_resample(const void* src, int Cs, void* dst, int Cd, int coff, int dt, int B, int Hi, int Wi, int mode,
                    const float* scale, const float* shift, int gld, int gc, int silu, cudaStream_t st);
// FiLM on a folded GroupNorm: scale *= 1 + s, shift = shift * (1 + s) + t, (s | t) = tb[b, off .. off + 2C)  (:377-381)
int launch_gn_modulate(float* scale, float* shift, const float* tb, int ld, int off, int B, int C, cudaStream_t st);
// tensor-core stem (<= 3 input channels): im2col of cat(x, cond) (NCHW fp32) into 64-channel bf16 NHWC pixels,
// channels (tap, c) rounded to bf16 then their rounding residuals; and the matching [Cout][64] fp32 weights
int launch_stem_im2col(const float* x, int Cx, const float* cond, int Cc, void* dst, int B, int H, int W, cudaStream_t st);
int launch_stem_weight(const float* w, int Cout, int C, float* w2, cudaStream_t st);
// 4..32 input channels: cat(x, cond) -> 64-channel bf16 NHWC pixels (values, then rounding residuals) and the
// matching [Cout][64][3][3] fp32 weights; the stem is then a plain 3x3 tensor-core conv
int launch_stem_nhwc(const float* x, int Cx, const float* cond, int Cc, void* dst, int B, int H, int W, cudaStream_t st);
int launch_stem_weight3(const float* w, int Cout, int C, float* w2, cudaStream_t st);
// generic NHWC (dt) -> NCHW fp32 copy, for eo_unet_read_activation
int launch_nhwc_to_nchw_f32(const void* src, int dt, float* dst, int B, int HW, int C,
                            cudaStream_t st);

// ---------------------------------------------------------------------------------------
// weight packing (run once at finalize)
// ---------------------------------------------------------------------------------------
// Packs torch conv weights w[Cout][Cin_total][k][k] (fp32) into a GEMM operand whose K axis
// is ordered (source segment, tap, channel).  One call per segment:
//   dst[n*stride_n + (k_off + tap*C + c)*stride_k] = row_scale[n] * w[row(n)][cin_off + c][tap]
// row(n) = row_map ? row_map[n] : n  (row_map[n] < 0 -> zero row); n in [0, Nout); row_scale (device, optional) = 1.
// If tap_fold (upsample sub-pixel folding) is non-null it is unused here (see engine).
int launch_pack_conv_weight(const float* w, int Cin_total, int ksize, int cin_off, int C,
                            void* dst, int dst_dt, long long stride_n, long long stride_k,
                            int k_off, int Nout, const int* row_map, cudaStream_t st, const float* row_scale = nullptr);
// wf[Cout][Cin][2][2] = the 3x3 weights w[Cout][Cin][3][3] folded for output parity (a, b) of a nearest x2
// upsample + conv3x3 (taps at low-resolution rows {a-1, a}, columns {b-1, b})
int launch_fold_upsample_weight(const float* w, int Cout, int Cin, int a, int b, float* wf, cudaStream_t st);
// dst[n] = row_scale[n] * ((a ? a[row(n)] : 0) + (b ? b[row(n)] : 0))
int launch_pack_bias(const float* a, const float* b, float* dst, int Nout, const int* row_map,
                     cudaStream_t st, const float* row_scale = nullptr);
// x[r][col0 .. col0 + ncols) *= scale, re-rounded to bf16 (rows of ld elements)
int launch_scale_cols_bf16(void* x, long long rows, int ld, int col0, int ncols, float scale, cudaStream_t st);

// ---------------------------------------------------------------------------------------
// tcgen05 tensor-core kernels (tc_conv.cu / tc_attn.cu)
// ---------------------------------------------------------------------------------------
constexpr int tc_patch_code(int r0, int nr, int c0, int nc) { return 1 | (r0 << 4) | (nr << 8) | (c0 << 12) | (nc << 16); }
constexpr int TC_PATCH_3X3 = tc_patch_code(0, 3, 0, 3);
struct TcConvSeg {
  const void* ptr = nullptr;  // bf16 NHWC [Bt, H, W, C]  (Bt may be 4*B for space-to-depth planes)
  int C = 0;
  int Bt = 0;
  int ntaps = 0;              // number of taps taken from this segment
  int8_t dh[9], dw[9];        // spatial shift of each tap (in this segment's grid)
  int dn[9];                  // batch-coordinate shift of each tap (plane select)
  // GroupNorm (+ SiLU) of this segment folded into the operand load (persistent kernel only; a 3x3
  // segment must then be a patch segment): act(x * gn_scale[b, gn_coff + c] + gn_shift[b, gn_coff + c])
  const float* gn_scale = nullptr;   // [B, gn_ld]
  const float* gn_shift = nullptr;
  int gn_ld = 0, gn_coff = 0, silu = 0;
  int patch = 0;              // taps served from ONE halo patch per 64 channels: tc_patch_code(r0, nr, c0, nc) = the
                              // rows r0 .. r0+nr-1 and columns c0 .. c0+nc-1 of the 3x3 neighbourhood, tap t at
                              // (dh, dw) = (r0 + t / nc - 1, c0 + t % nc - 1) -- TC_PATCH_3X3 for a plain 3x3 conv, a
                              // 2x2 corner for a sub-pixel convolution of Upsample; the segment's weights are then
                              // K-ordered (64-channel block, tap, channel).  0 = plain 128-pixel tiles, one per tap
  int stride = 1;             // 2: the segment's grid is [Bt, 2H, 2W, C] and output pixel (h, w) reads source pixel
                              // (2h + dh, 2w + dw): a stride-2 convolution straight from the full-resolution tensor
                              // (tensor map with traversal stride 2), no space-to-depth copy
};
struct TcConvParams {
  TcConvSeg seg[3];
  int nseg = 0;
  int B = 0, H = 0, W = 0;      // OUTPUT grid (== segment grid; stride/upsample are pre-passes)
  const void* Wp = nullptr;     // bf16 [Cout_pad][Ktot], K contiguous, K ordered (seg, tap, c)
  int Ktot = 0, Cout = 0;       // Cout: multiple of 64
  const float* bias = nullptr;      // [Cout] or null
  const float* bias_nc = nullptr;   // [B, ld_bias_nc] or null
  int ld_bias_nc = 0;
  const void* residual = nullptr;   // NHWC [B,H,W,Cout] (fp32 if res_f32, else bf16) or null
  int res_f32 = 0;
  void* out = nullptr;              // NHWC [B,H,W,Cout] (fp32 if out_f32, else bf16)
  int out_f32 = 0;
  // persistent kernel only: `out` as a strided view (elements) -- pixel (n, h, w) at out + out_sn*n + out_sh*h +
  // out_sw*w; 0 = dense NHWC.  Used by the sub-pixel convolutions of Upsample.
  long long out_sw = 0, out_sh = 0, out_sn = 0;
  // head convolution (out_nchw_C > 0): channels [0, out_nchw_C) of the accumulator (+ bias) go as fp32 NCHW
  // [B, out_nchw_C, H, W] to the pointer given at launch (tc_conv_launch's out_nchw) and nothing to `out`, which
  // may be null: no bf16 rounding of the network output, no layout pass.  Cout must be one 64-wide tile.
  int out_nchw_C = 0;
  double* stats = nullptr;          // [B, Cout, 2] per-channel (sum, sum of squares) of `out`,
                                    // accumulated with atomics (zeroed by the caller), or null
};
// fused statistics need a warp's 32 output rows inside one image
bool tc_conv_stats_supported(int H, int W);
// halo patches are available for this output grid (H % 16 == 0, W % 8 == 0)
bool tc_conv_patch_supported(int H, int W);
// A prepared launch (tensor maps encoded once at plan time)
struct TcConvPlan;
int tc_conv_plan_create(const TcConvParams& p, TcConvPlan** out);
void tc_conv_plan_destroy(TcConvPlan* p);
int tc_conv_launch(const TcConvPlan* plan, int B, cudaStream_t st, float* out_nchw = nullptr);
// development aid: per-CTA phase stamps of subsequent launches ([n_ctas][8] int64, device), null = off
void tc_conv_set_trace(long long* dev_buf, int n_ctas);

struct TcAttnParams {
  const void* qkv = nullptr;  // bf16 [B, T, heads*3*64]: per head q|k|v each padded to 64 channels
  void* out = nullptr;        // bf16 [B, T, heads*ch]
  int B = 0, T = 0, heads = 0, ch = 0;  // ch <= 64
  int k_one = 0;              // ch <= 48 (required there): q arrives multiplied by ch^-1/2 * log2(e) (folded into the qkv
                              // convolution) and channel round16(ch) of every head's k is 1.0 -- the kernel keeps -m in that channel of q
};
struct TcAttnPlan;
int tc_attn_plan_create(const TcAttnParams& p, TcAttnPlan** out);
void tc_attn_plan_destroy(TcAttnPlan* p);
int tc_attn_launch(const TcAttnPlan* plan, int B, cudaStream_t st);
void tc_attn_set_trace(long long* dev_buf, int n_ctas);   // development aid, see eo_debug_conv_trace

}  // namespace eo
