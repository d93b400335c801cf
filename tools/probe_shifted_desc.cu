// Hardware probe (development tool, not part of the library): can a tcgen05 shared-memory
// descriptor with 128B swizzle address a SHIFTED window of a halo patch?
//   patch  : pixels (ph, pw) at linear index q = ph * pitch + pw, 128 B (64 bf16) per pixel, written
//            the way TMA SWIZZLE_128B writes a box: 16-byte chunk c of pixel q lands at
//            q*128 + ((c ^ (q & 7)) << 4) from a 1024-aligned base.
//   window : 128 rows r = h*8 + w  ->  patch pixel (h + 1 + dh, w + 1 + dw); start address
//            = base + ((1+dh)*pitch + 1+dw) * 128, SBO = pitch * 128.
// Variants: pitch 10 or 16, descriptor base_offset field 0 or (start >> 7) & 7.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I eo_diffusion_b200/csrc -I include
//        tools/probe_shifted_desc.cu -o gpurun_out/probe_shifted_desc
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_bf16.h>
#include "tc_common.cuh"
using namespace eo;

__global__ void __launch_bounds__(128, 1)
k_probe(const __nv_bfloat16* __restrict__ patch, const __nv_bfloat16* __restrict__ Bm, float* __restrict__ D,
        int npix, int pitch, int dh, int dw, int use_base_offset) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (tc::smem_u32(smem_raw) & 1023)) & 1023);
  uint8_t* sA = smem;                       // npix * 128 bytes
  uint8_t* sB = smem + 40 * 1024;           // 64 rows * 128 bytes
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 48 * 1024);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < npix * 8; i += 128) {
    const int q = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sA + q * 128 + ((c ^ (q & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(patch + (size_t)q * 64 + c * 8);
  }
  for (int i = tid; i < 64 * 8; i += 128) {
    const int q = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sB + q * 128 + ((c ^ (q & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(Bm + (size_t)q * 64 + c * 8);
  }
  if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) { tc::tmem_alloc(tptr, 64); tc::tmem_relinquish(); }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tptr;
  if (tid == 0) {
    const uint32_t start = tc::smem_u32(sA) + (uint32_t)(((1 + dh) * pitch + 1 + dw) * 128);
    uint64_t ad = 0;
    ad |= (uint64_t)((start & 0x3FFFF) >> 4);
    ad |= (uint64_t)1 << 16;
    ad |= (uint64_t)((pitch * 128) >> 4) << 32;
    ad |= (uint64_t)1 << 46;
    if (use_base_offset) ad |= (uint64_t)((start >> 7) & 7) << 49;
    ad |= (uint64_t)2 << 61;
    const uint64_t bd = tc::make_sw128_desc(tc::smem_u32(sB));
    const uint32_t idesc = tc::make_idesc_bf16(128, 64, 0, 0);
    for (int k = 0; k < 4; ++k)
      tc::umma_f16_ss(tmem, tc::desc_advance(ad, k * 32), tc::desc_advance(bd, k * 32), idesc, k ? 1u : 0u);
    tc::umma_commit(bar);
  }
  tc::mbar_wait(bar, 0);
  tc::tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t v[32];
    tc::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 64); }
}

int main() {
  const int MAXPIX = 18 * 16;
  std::vector<__nv_bfloat16> hp(MAXPIX * 64), hb(64 * 64);
  std::vector<float> fp(MAXPIX * 64), fb(64 * 64);
  srand(1);
  for (size_t i = 0; i < hp.size(); ++i) { float v = (rand() % 255 - 127) / 64.0f; hp[i] = __float2bfloat16(v); fp[i] = __bfloat162float(hp[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { float v = (rand() % 255 - 127) / 64.0f; hb[i] = __float2bfloat16(v); fb[i] = __bfloat162float(hb[i]); }
  __nv_bfloat16 *dp, *db; float* dD;
  cudaMalloc(&dp, hp.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dp, hp.data(), hp.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024 + 1024);
  std::vector<float> hD(128 * 64);
  int pitches[2] = {10, 16};
  for (int pi = 0; pi < 2; ++pi)
    for (int bo = 0; bo < 2; ++bo) {
      const int pitch = pitches[pi];
      double worst = 0; int bad_taps = 0;
      for (int dh = -1; dh <= 1; ++dh)
        for (int dw = -1; dw <= 1; ++dw) {
          cudaMemset(dD, 0, 128 * 64 * 4);
          k_probe<<<1, 128, 50 * 1024 + 1024>>>(dp, db, dD, 18 * pitch, pitch, dh, dw, bo);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("pitch %d bo %d tap (%d,%d): CUDA error %s\n", pitch, bo, dh, dw, cudaGetErrorString(e)); return 1; }
          cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
          double err = 0;
          for (int r = 0; r < 128; ++r) {
            const int h = r >> 3, w = r & 7;
            const int q = (h + 1 + dh) * pitch + (w + 1 + dw);
            for (int n = 0; n < 64; ++n) {
              float ref = 0;
              for (int k = 0; k < 64; ++k) ref += fp[q * 64 + k] * fb[n * 64 + k];
              err = fmax(err, fabs(ref - hD[r * 64 + n]));
            }
          }
          printf("pitch %2d base_offset %d tap (%2d,%2d): max abs err %.4g\n", pitch, bo, dh, dw, err);
          if (err > 1e-2) ++bad_taps;
          worst = fmax(worst, err);
        }
      printf("== pitch %2d base_offset_field %s: %s (worst %.4g, %d/9 taps wrong)\n", pitch, bo ? "(start>>7)&7" : "0",
             bad_taps ? "FAIL" : "OK", worst, bad_taps);
    }
  return 0;
}
