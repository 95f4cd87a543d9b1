// common.cuh -- shared helpers for the libvmgb kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/vmgb.h"

namespace vmgb {

extern int g_last_cuda_error;
extern long long g_launch_count;
extern int g_pack_pending;      // a tcgen05 weight pack was launched since the last programmatic-dependent-launch kernel

inline int check_launch() {
  ++g_launch_count;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_cuda_error = (int)e; return VMGB_ECUDA; }
  return VMGB_OK;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline bool dt_ok(int dt) { return dt == VMGB_F32 || dt == VMGB_BF16; }
inline size_t dt_size(int dt) { return dt == VMGB_BF16 ? 2 : 4; }

// ---- dtype-generic element access (runtime dtype; branches are warp-uniform) -----------------
__device__ __forceinline__ float ldf(const void* p, int dt, long long i) {
  if (dt == VMGB_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void stf(void* p, int dt, long long i, float v) {
  if (dt == VMGB_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(p)[i] = v;
}
// 4 consecutive elements; i must be a multiple of 4 and the base 16 B (f32) / 8 B (bf16) aligned
__device__ __forceinline__ float4 ld4(const void* p, int dt, long long i) {
  if (dt == VMGB_BF16) {
    uint2 r = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p) + i);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
}
__device__ __forceinline__ void st4(void* p, int dt, long long i, float4 v) {
  if (dt == VMGB_BF16) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p) + i) = r;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = v;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
  switch (act) {
    case VMGB_ACT_RELU: return fmaxf(v, 0.f);
    case VMGB_ACT_LRELU: return v > 0.f ? v : v * alpha;
    case VMGB_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    case VMGB_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

// Coordinate pipeline of flow_warp + ATen grid_sampler_2d (align_corners=True, padding 'border'),
// kept free of FMA contraction so that nearest-mode rounding follows the reference op order:
//   g = 2*(x+flow)/(size-1) - 1          (trajectory.py:101-106)
//   u = ((g+1)/2)*(size-1)               (ATen GridSampler.cuh grid_sampler_unnormalize)
//   u = min(size-1, max(u, 0))           (clip_coordinates)
// `sem` selects whose rounding the division by the Python scalar (size-1) follows:
//   VMGB_COORD_ATEN_CPU  : true division (ATen CPU div kernel)
//   VMGB_COORD_ATEN_CUDA : multiplication by the fp32 reciprocal (ATen CUDA div kernel with a CPU-scalar divisor computes
//                          a * (1/b), "may lose one bit of precision") -- what the reference's own GPU forward does
__device__ __forceinline__ float warp_coord(float base, float flow, int size, int sem) {
  float denom = (float)max(size - 1, 1);
  float gf = __fadd_rn(base, flow);
  float q = sem == VMGB_COORD_ATEN_CUDA ? __fmul_rn(__fmul_rn(2.0f, gf), __fdiv_rn(1.0f, denom)) : __fdiv_rn(__fmul_rn(2.0f, gf), denom);
  float g = __fsub_rn(q, 1.0f);
  float u = __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.0f), 2.0f), (float)(size - 1));
  return fminf((float)(size - 1), fmaxf(u, 0.f));
}

}  // namespace vmgb
