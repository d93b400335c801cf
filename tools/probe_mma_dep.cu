// Hardware probe (development tool, not part of the library): issue interval of tcgen05.mma
// instructions that accumulate into the SAME TMEM accumulator versus round-robin over several
// independent accumulators, for M=128 (cta_group::1) and N in {64, 128, 256}, kind::f16, K=16.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I eo_diffusion_b200/csrc -I include
//        tools/probe_mma_dep.cu -o gpurun_out/probe_mma_dep
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "tc_common.cuh"
using namespace eo;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// TS: A from TMEM (columns 384..), B from smem, optionally MN-major (the attention P V product)
template <int N, int BMN>
__global__ void __launch_bounds__(128, 1) k_probe_ts(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (tc::smem_u32(smem_raw) & 1023)) & 1023);
  uint8_t* sB = smem;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32768 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) { tc::tmem_alloc(tptr, 512); tc::tmem_relinquish(); }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tptr;
  if (warp == 0) {
    const uint64_t bd = tc::make_sw128_desc(tc::smem_u32(sB));
    constexpr uint32_t idesc = tc::make_idesc_bf16(128, N, 0, BMN);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      if (tc::elect_one()) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ts(tmem, tmem + 384 + k * 8, tc::desc_advance(bd, BMN ? k * 16 * 128 : k * 32), idesc, 1u);
        }
        tc::umma_commit(bar);
      }
      __syncwarp();
      tc::mbar_wait(bar, rep & 1);
      t1 = clock64();
    }
    if (tid == 0) out[0] = t1 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 512); }
}
template <int N, int BMN>
void run_ts(long long* d, int iters) {
  cudaFuncSetAttribute(k_probe_ts<N, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  k_probe_ts<N, BMN><<<1, 128, 40 * 1024>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("TS N=%d: %s\n", N, cudaGetErrorString(e)); exit(1); }
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("TS (A in TMEM) M=128 N=%3d B %s-major: %7.1f clk per MMA (ideal %3d)\n", N, BMN ? "MN" : "K", h / ((double)iters * 4),
         128 * N * 16 * 2 / 8192);
}

template <int N, int NACC>
__global__ void __launch_bounds__(128, 1) k_probe(long long* out, int iters, int a_shift_rows, int a_sbo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (tc::smem_u32(smem_raw) & 1023)) & 1023);
  uint8_t* sA = smem;                  // 128 rows x 128 B
  uint8_t* sB = smem + 24576;          // 256 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 24576 + 32768);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (24576 + 32768) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) { tc::tmem_alloc(tptr, 512); tc::tmem_relinquish(); }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tptr;
  if (warp == 0) {
    const uint64_t ad = tc::make_sw128_desc_sbo(tc::smem_u32(sA) + a_shift_rows * 128, a_sbo);
    const uint64_t bd = tc::make_sw128_desc(tc::smem_u32(sB));
    constexpr uint32_t idesc = tc::make_idesc_bf16(128, N, 0, 0);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {       // rep 0 warms up
      t0 = clock64();
      if (tc::elect_one()) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int a = 0; a < NACC; ++a)
              tc::umma_f16_ss(tmem + a * N, tc::desc_advance(ad, k * 32), tc::desc_advance(bd, k * 32), idesc, 1u);
          }
        }
        tc::umma_commit(bar);
      }
      __syncwarp();
      const long long ti = clock64();
      tc::mbar_wait(bar, rep & 1);
      t1 = clock64();
      if (tid == 0) out[1] = ti - t0;
    }
    if (tid == 0) out[0] = t1 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 512); }
}

template <int N, int NACC>
void run(long long* d, int iters, int shift = 0, int sbo = 1024) {
  cudaFuncSetAttribute(k_probe<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
  k_probe<N, NACC><<<1, 128, 60 * 1024>>>(d, iters, shift, sbo);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d NACC=%d: %s\n", N, NACC, cudaGetErrorString(e)); exit(1); }
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double n_mma = (double)iters * 4 * NACC;
  printf("M=128 N=%3d accumulators=%d A start +%2d rows, SBO %4d, %4d MMAs: %7.1f clk per MMA to complete, %7.1f to issue (ideal %3d)\n", N, NACC,
         shift, sbo, (int)n_mma, h[0] / n_mma, h[1] / n_mma, 128 * N * 16 * 2 / 8192);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  const int iters = 512;
  run<32, 1>(d, iters); run<32, 2>(d, iters); run<32, 4>(d, iters);
  run<64, 1>(d, iters); run<64, 2>(d, iters); run<64, 4>(d, iters);
  run<128, 1>(d, iters); run<128, 2>(d, iters); run<128, 4>(d, iters);
  run<256, 1>(d, iters); run<256, 2>(d, iters);
  // A operand as a shifted window of a halo patch (tc_conv3.cu): start +k pixel rows, 8-row groups 1280 B apart
  for (int sh : {11}) { run<64, 1>(d, iters, sh, 1280); run<128, 1>(d, iters, sh, 1280); run<256, 1>(d, iters, sh, 1280); }
  run<64, 1>(d, iters, 1, 1024); run<64, 1>(d, iters, 0, 2048);
  for (int it : {1, 2, 4, 8, 16, 64}) run<128, 1>(d, it);
  run_ts<32, 1>(d, iters); run_ts<64, 0>(d, iters); run_ts<64, 1>(d, iters); run_ts<128, 0>(d, iters); run_ts<128, 1>(d, iters); run_ts<256, 1>(d, iters);
  return 0;
}
