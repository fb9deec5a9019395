// jt_dev_output.cuh -- the reference's save path per pixel, shared by the single-GPU download
// (jt_api.cu) and the fused multi-GPU reduce + finalize kernel (jt_group.cu).
#pragma once
#include <cuda_runtime.h>

// rgb_to_srgb (src/color.jl:25-29) + clamp01nan + 8-bit quantisation of save_image (src/sceneio.jl:97-113) for one
// RGBA pixel already scaled to its mean. Julia's `^(rgb, 1/2.4f0)` is Float32(exp2(log2(Float64(x)) * Float64(y))).
__device__ __forceinline__ uchar4 jt_srgb8_pixel(float4 p) {
  float c[4] = {p.x, p.y, p.z, p.w};
  unsigned char b[4];
  const double expo = (double)(1.0f / 2.4f);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    float v = c[k];
    if (k < 3) v = (v <= 0.0031308f) ? 12.92f * v : 1.055f * (float)exp2(log2((double)v) * expo) - 0.055f;
    if (!(v == v)) v = 0.0f;  // clamp01nan: NaN -> 0
    v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
    b[k] = (unsigned char)__float2int_rn(v * 255.0f);
  }
  return make_uchar4(b[0], b[1], b[2], b[3]);
}
