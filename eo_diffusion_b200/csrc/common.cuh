// Shared helpers for libeo_b200: error plumbing, small device utilities.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>

#include "../../include/eo_b200.h"

namespace eo {

// thread-local last-error message (eo_last_error)
void set_error(const char* fmt, ...);
const char* get_error();

#define EO_CHECK_CUDA(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      eo::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                    __LINE__);                                                           \
      return EO_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

#define EO_REQUIRE(cond, code, ...)                                                      \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      eo::set_error(__VA_ARGS__);                                                        \
      return (code);                                                                     \
    }                                                                                    \
  } while (0)

// checks the kernel launch that just happened
#define EO_CHECK_LAUNCH()                                                                \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      eo::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),          \
                    __FILE__, __LINE__);                                                 \
      return EO_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------
// Programmatic dependent launch (griddepcontrol): inside one UNet forward every kernel depends on the one before it, so
// the chain pays a launch + scheduling gap and a cold prologue (barrier init, TMEM allocation, descriptor prefetch)
// ~170 times per step.  A kernel launched with the programmatic-serialization attribute may become resident while its
// predecessor is still running (as SM resources free up), runs its prologue, and blocks in pdl_wait() until the
// predecessor has COMPLETED and its writes are visible -- nothing a kernel reads or writes is touched before pdl_wait().
// pdl_trigger() (right after the wait, so at most one generation of waiting kernels exists) lets the NEXT kernel do the
// same.  Both are no-ops in a kernel launched the ordinary way.
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
// set by the engine around the ops of a forward: true when the previous operation enqueued on the stream was one of
// this library's kernels (a programmatic edge needs a kernel on both ends); thread-local like the handles' use
bool pdl_allowed();
void pdl_set_allowed(bool on);

#ifdef __CUDACC__
// <<<grid, block, smem, st>>> with the programmatic-serialization attribute when pdl_allowed()
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_allowed() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

int num_sms();  // SM count of the current device (cached)

// SiLU for bf16-bound outputs: x*sigmoid(x) = h + h*tanh(h), h = x/2 -- one MUFU (tanh.approx,
// max relative error 2^-11, a quarter of a bf16 ulp) instead of ex2 + rcp
__device__ __forceinline__ float silu_f(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
// accurate variant for the fp32 parity mode
__device__ __forceinline__ float silu_acc(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace eo
