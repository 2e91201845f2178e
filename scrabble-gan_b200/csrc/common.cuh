// Shared device/host helpers for libsgan (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/sgan.h"

struct sg_ctx {
  int device;
  cudaStream_t stream;
  int num_sms;
  void* encode_tiled;        // PFN cuTensorMapEncodeTiled (resolved lazily)
  long long launches;        // number of kernels launched through this context
  int speed_mode;            // 1: bf16 speed mode -- small fp32 GEMMs of the non-local block use bf16 warp-level tensor ops
};

void sg_set_error(const char* fmt, ...);

#define SG_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      sg_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__,    \
                   cudaGetErrorString(_e));                                                   \
      return SG_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define SG_REQUIRE(cond, ...)                                                                 \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      sg_set_error(__VA_ARGS__);                                                              \
      return SG_ERR_ARG;                                                                      \
    }                                                                                         \
  } while (0)

#define SG_POST_LAUNCH(ctx)                                                                   \
  do {                                                                                        \
    (ctx)->launches++;                                                                        \
    SG_CHECK_CUDA(cudaGetLastError());                                                        \
  } while (0)

static inline int sg_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------------
// typed load / store helpers (operand tensors are fp32 or bf16; arithmetic is always fp32)
// ---------------------------------------------------------------------------------------------------
template <typename T> struct sg_dt;
template <> struct sg_dt<float> { static constexpr int id = SG_F32; };
template <> struct sg_dt<__nv_bfloat16> { static constexpr int id = SG_BF16; };

__device__ __forceinline__ float sg_ld(const float* p) { return *p; }
__device__ __forceinline__ float sg_ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void sg_st(float* p, float v) { *p = v; }
__device__ __forceinline__ void sg_st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float4 sg_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 sg_ld4(const __nv_bfloat16* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void sg_st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void sg_st4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

__device__ __forceinline__ float sg_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float sg_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum (blockDim.x multiple of 32, <= 1024); result valid in all threads
__device__ __forceinline__ float sg_block_sum(float v, float* smem32) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = sg_warp_sum(v);
  __syncthreads();
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? smem32[threadIdx.x] : 0.f;
  if (w == 0) {
    r = sg_warp_sum(r);
    if (lane == 0) smem32[0] = r;
  }
  __syncthreads();
  return smem32[0];
}

#define SG_DISPATCH_DT(dt, T, ...)                              \
  do {                                                          \
    if ((dt) == SG_F32) {                                       \
      using T = float;                                          \
      __VA_ARGS__;                                              \
    } else if ((dt) == SG_BF16) {                               \
      using T = __nv_bfloat16;                                  \
      __VA_ARGS__;                                              \
    } else {                                                    \
      sg_set_error("bad dtype %d", (int)(dt));                  \
      return SG_ERR_ARG;                                        \
    }                                                           \
  } while (0)
