// On-device post-processing of finished samples and the image-quality metrics that follow the
// sampling path: reference inference.py:128-150 --
//   samples.clip(0, 1) | (samples + 1) / 2                                   :128, :136
//   cond = image * ((mask + 0.7).clip(0, 1))                                 :135
//   torchvision adjust_brightness(x, 3) = (3 * x).clamp(0, 1) for floats     :141-142, :149
//   image.min(), x.mean() (the host branches on them)                        :128, :141, :149
//   torchmetrics peak_signal_noise_ratio / structural_similarity_index_measure(data_range=1.0)   :138
//   torchvision save_image(x, path, nrow=...) = make_grid + `mul(255).add_(0.5).clamp_(0, 255).to(uint8)`   :143-150,
//     and inside the sampling loop itself, diffusion/model.py:62-66 (`save_image((x_t + 1.) / 2., ...)`)
// The elementwise passes are bit-exact with the fp32 torch ops (round-to-nearest intrinsics, no FMA
// contraction).  PSNR and SSIM follow the published torchmetrics algorithms (torchmetrics is not part of
// the reference checkout: its arithmetic is restated in oracle/postprocess.py and pinned to four digits on the
// library's published docstring examples):
//   PSNR = 10 log10(data_range^2 / mean((a - b)^2)) over the whole batch;
//   SSIM: 11 x 11 Gaussian window (sigma 1.5) over reflect-padded images, the padded border cropped again,
//         i.e. exactly the windows that lie inside the image; c1 = (0.01 R)^2, c2 = (0.03 R)^2; mean per image,
//         then over the batch.
// All kernels are HBM-bound: one read of each operand (SSIM tiles re-read a 5-pixel halo).
#include "common.cuh"

namespace eo {
namespace {

inline unsigned grid_for(long long n, int per_block) {
  long long need = ceil_div(n, per_block), cap = (long long)num_sms() * 16;
  return (unsigned)(need < 1 ? 1 : (need < cap ? need : cap));
}

// mode 0: clamp(x, 0, 1); 1: (x + 1) / 2; 2: clamp(x * factor, 0, 1)
template <int MODE>
__global__ void __launch_bounds__(256) k_post_map(const float* __restrict__ x, float* __restrict__ out, long long n, float factor) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    float r;
    if (MODE == 0) r = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);                 // clamp: NaN passes through
    else if (MODE == 1) r = __fmul_rn(__fadd_rn(v, 1.0f), 0.5f);            // division by 2 is exact
    else { const float s = __fmul_rn(factor, v); r = s < 0.f ? 0.f : (s > 1.f ? 1.f : s); }
    out[i] = r;
  }
}

// out[b,c,p] = image[b,c,p] * clamp(mask[b,0,p] + 0.7, 0, 1)
__global__ void __launch_bounds__(256)
k_post_dim(const float* __restrict__ image, const float* __restrict__ mask, float* __restrict__ out, int C, long long HW, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / (C * HW);
    const long long p = i % HW;
    float m = __fadd_rn(mask[b * HW + p], 0.7f);
    m = m < 0.f ? 0.f : (m > 1.f ? 1.f : m);
    out[i] = __fmul_rn(image[i], m);
  }
}

// torchvision.utils.make_grid (padding, pad_value; a single-channel batch is repeated to three channels; ONE image is
// returned as it is, without a border) followed by save_image's quantisation, written as the HWC uint8 array
// PIL.Image.fromarray takes.  PRE 1 applies (x + 1) / 2 first (model.py:63).  One thread per grid pixel.
template <int PRE>
__global__ void __launch_bounds__(256)
k_post_grid_u8(const float* __restrict__ x, unsigned char* __restrict__ out, int B, int C, int H, int W, int Cg,
               int xmaps, int pad, int GH, int GW, float pad_value) {
  const long long npix = (long long)GH * GW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    const int gy = (int)(i / GW), gx = (int)(i % GW);
    const int cell_h = H + pad, cell_w = W + pad;
    const int ry = gy - pad, rx = gx - pad;                 // position relative to the first image's corner
    int k = -1, iy = 0, ix = 0;
    if (ry >= 0 && rx >= 0) {
      const int my = ry / cell_h, mx = rx / cell_w;
      iy = ry - my * cell_h; ix = rx - mx * cell_w;
      if (iy < H && ix < W && mx < xmaps && my * xmaps + mx < B) k = my * xmaps + mx;
    }
    for (int c = 0; c < Cg; ++c) {
      float v = pad_value;
      if (k >= 0) {
        v = x[(((long long)k * C + (C == 1 ? 0 : c)) * H + iy) * W + ix];
        if (PRE == 1) v = __fmul_rn(__fadd_rn(v, 1.0f), 0.5f);
      }
      v = __fadd_rn(__fmul_rn(v, 255.0f), 0.5f);
      v = v < 0.f ? 0.f : (v > 255.f ? 255.f : v);
      out[i * Cg + c] = (unsigned char)__float2int_rz(v);
    }
  }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// acc[0] += sum(x), acc[1] = min, acc[2] = max (as doubles; acc pre-set to {0, +inf, -inf})
__global__ void __launch_bounds__(256) k_post_stats(const float* __restrict__ x, long long n, double* __restrict__ acc) {
  double s = 0.0;
  float mn = INFINITY, mx = -INFINITY;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    s += (double)v; mn = fminf(mn, v); mx = fmaxf(mx, v);
  }
  __shared__ double sh[8]; __shared__ float shmn[8], shmx[8];
  s = warp_sum_d(s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { sh[w] = s; shmn[w] = mn; shmx[w] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { s += sh[i]; mn = fminf(mn, shmn[i]); mx = fmaxf(mx, shmx[i]); }
    atomicAdd(&acc[0], s);
    // min / max of doubles holding floats: compare-and-swap loops on the bit patterns
    unsigned long long* pm = reinterpret_cast<unsigned long long*>(&acc[1]);
    unsigned long long old = *pm;
    while ((double)mn < __longlong_as_double((long long)old)) {
      const unsigned long long prev = atomicCAS(pm, old, (unsigned long long)__double_as_longlong((double)mn));
      if (prev == old) break;
      old = prev;
    }
    unsigned long long* px = reinterpret_cast<unsigned long long*>(&acc[2]);
    old = *px;
    while ((double)mx > __longlong_as_double((long long)old)) {
      const unsigned long long prev = atomicCAS(px, old, (unsigned long long)__double_as_longlong((double)mx));
      if (prev == old) break;
      old = prev;
    }
  }
}
__global__ void k_post_stats_init(double* acc) { acc[0] = 0.0; acc[1] = INFINITY; acc[2] = -INFINITY; acc[3] = 0.0; }
// out[0] = mean, out[1] = min, out[2] = max
__global__ void k_post_stats_fin(const double* acc, long long n, float* out) {
  out[0] = (float)(acc[0] / (double)n); out[1] = (float)acc[1]; out[2] = (float)acc[2];
}

// acc[0] += sum((a - b)^2), squared differences taken in fp32 like torch.pow(preds - target, 2)
__global__ void __launch_bounds__(256) k_sse(const float* __restrict__ a, const float* __restrict__ b, long long n, double* __restrict__ acc) {
  double s = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = __fsub_rn(a[i], b[i]);
    s += (double)__fmul_rn(d, d);
  }
  __shared__ double sh[8];
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { for (int i = 1; i < 8; ++i) s += sh[i]; atomicAdd(&acc[0], s); }
}
__global__ void k_psnr_fin(const double* acc, long long n, float data_range, float* out) {
  const double mse = acc[0] / (double)n;
  out[0] = (float)(10.0 * log10((double)data_range * (double)data_range / mse));
}

// ---- SSIM: one CTA per 16 x 16 tile of valid window centres of one (image, channel) plane
constexpr int ST = 16, SK = 11, SR = 5, SP = ST + 2 * SR;   // tile, kernel, radius, patch (26)
struct Gauss { float w[SK]; };

__global__ void __launch_bounds__(256)
k_ssim(const float* __restrict__ pred, const float* __restrict__ target, int C, int H, int W, Gauss g, float c1, float c2,
       double* __restrict__ per_image) {
  __shared__ float sx[SP][SP + 1], sy[SP][SP + 1];
  __shared__ float hx[5][SP][ST + 1];            // horizontally filtered x, y, xx, yy, xy
  __shared__ double red[8];
  const int plane = blockIdx.z;                  // b * C + c
  const int oh0 = blockIdx.y * ST, ow0 = blockIdx.x * ST;   // first window centre of the tile, in crop coordinates
  const int Hc = H - 2 * SR, Wc = W - 2 * SR;
  const float* px = pred + (long long)plane * H * W;
  const float* py = target + (long long)plane * H * W;
  // window centre (i, j) in crop coordinates covers image rows i .. i + 10
  for (int idx = threadIdx.x; idx < SP * SP; idx += blockDim.x) {
    const int r = idx / SP, c = idx - r * SP;
    const int ih = oh0 + r, iw = ow0 + c;
    const bool in = ih < H && iw < W;
    sx[r][c] = in ? px[(long long)ih * W + iw] : 0.f;
    sy[r][c] = in ? py[(long long)ih * W + iw] : 0.f;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < SP * ST; idx += blockDim.x) {
    const int r = idx / ST, c = idx - r * ST;
    float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
    for (int k = 0; k < SK; ++k) {
      const float x = sx[r][c + k], y = sy[r][c + k], w = g.w[k];
      a = fmaf(w, x, a); b = fmaf(w, y, b);
      aa = fmaf(w, x * x, aa); bb = fmaf(w, y * y, bb); ab = fmaf(w, x * y, ab);
    }
    hx[0][r][c] = a; hx[1][r][c] = b; hx[2][r][c] = aa; hx[3][r][c] = bb; hx[4][r][c] = ab;
  }
  __syncthreads();
  const int r = threadIdx.x / ST, c = threadIdx.x - r * ST;
  double v = 0.0;
  if (oh0 + r < Hc && ow0 + c < Wc) {
    float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < SK; ++k) {
      const float w = g.w[k];
#pragma unroll
      for (int q = 0; q < 5; ++q) m[q] = fmaf(w, hx[q][r + k][c], m[q]);
    }
    const float mu_x2 = m[0] * m[0], mu_y2 = m[1] * m[1], mu_xy = m[0] * m[1];
    const float s_x = m[2] - mu_x2, s_y = m[3] - mu_y2, s_xy = m[4] - mu_xy;
    const float upper = 2.f * s_xy + c2, lower = s_x + s_y + c2;
    v = (double)(((2.f * mu_xy + c1) * upper) / ((mu_x2 + mu_y2 + c1) * lower));
  }
  v = warp_sum_d(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) v += red[i];
    atomicAdd(&per_image[plane / C], v);
  }
}
__global__ void k_zero_d(double* p, int n) { for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = 0.0; }
// per_image_out[b] = acc[b] / (C * Hc * Wc); mean_out = mean over b
__global__ void k_ssim_fin(const double* acc, int B, double count, float* per_image_out, float* mean_out) {
  double tot = 0.0;
  for (int b = 0; b < B; ++b) {
    const float s = (float)(acc[b] / count);
    if (per_image_out) per_image_out[b] = s;
    tot += (double)s;
  }
  mean_out[0] = (float)(tot / (double)B);
}

}  // namespace
}  // namespace eo

using namespace eo;

extern "C" {

int eo_post_map(const float* x, float* out, int64_t n, int mode, float factor, void* stream) {
  EO_REQUIRE(x && out && n >= 0, EO_ERR_ARG, "eo_post_map: null pointer");
  EO_REQUIRE(mode >= 0 && mode <= 2, EO_ERR_ARG, "eo_post_map: mode %d (0 clip, 1 (x+1)/2, 2 brightness)", mode);
  if (n == 0) return EO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = grid_for(n, 256);
  if (mode == 0) k_post_map<0><<<g, 256, 0, st>>>(x, out, n, factor);
  else if (mode == 1) k_post_map<1><<<g, 256, 0, st>>>(x, out, n, factor);
  else k_post_map<2><<<g, 256, 0, st>>>(x, out, n, factor);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

int eo_post_dim_masked(const float* image, const float* mask, float* out, int B, int C, int HW, void* stream) {
  EO_REQUIRE(image && mask && out && B > 0 && C > 0 && HW > 0, EO_ERR_ARG, "eo_post_dim_masked: bad argument");
  const long long n = (long long)B * C * HW;
  k_post_dim<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(image, mask, out, C, HW, n);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

int eo_post_grid_u8(const float* x, unsigned char* out, int B, int C, int H, int W, int nrow, int padding, float pad_value,
                    int pre, int* grid_hw_or_null, void* stream) {
  EO_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && nrow > 0 && padding >= 0, EO_ERR_ARG, "eo_post_grid_u8: bad geometry");
  EO_REQUIRE(pre == 0 || pre == 1, EO_ERR_ARG, "eo_post_grid_u8: pre %d (0 none, 1 (x+1)/2)", pre);
  // make_grid: one image comes back unpadded; otherwise xmaps = min(nrow, B) columns, ceil(B / xmaps) rows of
  // (H + padding) x (W + padding) cells plus one more border on the far sides
  const int pad = B == 1 ? 0 : padding;
  const int xmaps = nrow < B ? nrow : B;
  const int ymaps = (B + xmaps - 1) / xmaps;
  const long long GH = B == 1 ? H : (long long)(H + pad) * ymaps + pad, GW = B == 1 ? W : (long long)(W + pad) * xmaps + pad;
  EO_REQUIRE(GH <= INT32_MAX && GW <= INT32_MAX, EO_ERR_ARG, "eo_post_grid_u8: grid too large");
  if (grid_hw_or_null) { grid_hw_or_null[0] = (int)GH; grid_hw_or_null[1] = (int)GW; grid_hw_or_null[2] = C == 1 ? 3 : C; }
  if (!x && !out) return EO_OK;                         // geometry query
  EO_REQUIRE(x && out, EO_ERR_ARG, "eo_post_grid_u8: null pointer");
  const int Cg = C == 1 ? 3 : C;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = grid_for(GH * GW, 256);
  if (pre == 0) k_post_grid_u8<0><<<g, 256, 0, st>>>(x, out, B, C, H, W, Cg, xmaps, pad, (int)GH, (int)GW, pad_value);
  else k_post_grid_u8<1><<<g, 256, 0, st>>>(x, out, B, C, H, W, Cg, xmaps, pad, (int)GH, (int)GW, pad_value);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

int eo_post_stats(const float* x, int64_t n, double* workspace4, float* out3, void* stream) {
  EO_REQUIRE(x && workspace4 && out3 && n > 0, EO_ERR_ARG, "eo_post_stats: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  k_post_stats_init<<<1, 1, 0, st>>>(workspace4);
  k_post_stats<<<grid_for(n, 256 * 8), 256, 0, st>>>(x, n, workspace4);
  k_post_stats_fin<<<1, 1, 0, st>>>(workspace4, n, out3);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

int eo_psnr(const float* preds, const float* target, int64_t n, float data_range, double* workspace4, float* out1, void* stream) {
  EO_REQUIRE(preds && target && workspace4 && out1 && n > 0 && data_range > 0.f, EO_ERR_ARG, "eo_psnr: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  k_post_stats_init<<<1, 1, 0, st>>>(workspace4);
  k_sse<<<grid_for(n, 256 * 8), 256, 0, st>>>(preds, target, n, workspace4);
  k_psnr_fin<<<1, 1, 0, st>>>(workspace4, n, data_range, out1);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

int eo_ssim(const float* preds, const float* target, int B, int C, int H, int W, float data_range, const float* gauss11,
            double* workspaceB, float* per_image_or_null, float* mean_out, void* stream) {
  EO_REQUIRE(preds && target && gauss11 && workspaceB && mean_out, EO_ERR_ARG, "eo_ssim: null pointer");
  EO_REQUIRE(B > 0 && C > 0 && H > 2 * SR && W > 2 * SR, EO_ERR_ARG,
             "eo_ssim: %dx%d images are smaller than the 11x11 window", H, W);
  EO_REQUIRE((long long)B * C <= 65535, EO_ERR_ARG, "eo_ssim: more than 65535 image planes");
  cudaStream_t st = (cudaStream_t)stream;
  Gauss g;
  for (int i = 0; i < SK; ++i) g.w[i] = gauss11[i];       // host array (11 floats)
  const float c1 = (0.01f * data_range) * (0.01f * data_range), c2 = (0.03f * data_range) * (0.03f * data_range);
  const int Hc = H - 2 * SR, Wc = W - 2 * SR;
  k_zero_d<<<1, 256, 0, st>>>(workspaceB, B);
  dim3 grid((unsigned)ceil_div(Wc, ST), (unsigned)ceil_div(Hc, ST), (unsigned)(B * C));
  k_ssim<<<grid, 256, 0, st>>>(preds, target, C, H, W, g, c1, c2, workspaceB);
  k_ssim_fin<<<1, 1, 0, st>>>(workspaceB, B, (double)C * Hc * Wc, per_image_or_null, mean_out);
  EO_CHECK_LAUNCH();
  return EO_OK;
}

}  // extern "C"
