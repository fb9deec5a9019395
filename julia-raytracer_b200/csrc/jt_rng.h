// jt_rng.h -- per-pixel counter-based RNG shared by the CUDA path and the CPU oracle.
//
// The reference draws from Julia's unseeded task-local generator (`rand(Float32)`,
// src/sampling.jl:18-22), so its images are not reproducible and "a fixed sample set" is only
// definable against a shared counter-based stream:
//   key     = (seed, pixel index, sample index)
//   counter = index of the draw inside one trace_sample call, in the reference's draw order
//             (SURVEY.md §8a "RNG draw order")
//   value   = (u32 >> 8) * 2^-24  -- a multiple of 2^-24 in [0,1), like Julia's rand(Float32)
// The mixing function is the splitmix64 finaliser; there is no reference counterpart.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define JT_RNG_HD __host__ __device__ __forceinline__
#else
#define JT_RNG_HD inline
#endif

JT_RNG_HD uint64_t jt_mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// One key per (seed, pixel, sample); pixel = W*j + i (0-based), sample = global sample index.
JT_RNG_HD uint64_t jt_rng_key(uint64_t seed, uint32_t pixel, uint32_t sample) {
  uint64_t k = jt_mix64(seed + 0x9E3779B97F4A7C15ull);
  k = jt_mix64(k ^ (((uint64_t)pixel << 32) | (uint64_t)sample));
  return k;
}

// The draw-th uniform of the stream identified by key.
JT_RNG_HD float jt_rng_float(uint64_t key, uint32_t draw) {
  uint64_t z = jt_mix64(key + (uint64_t)(draw + 1u) * 0x9E3779B97F4A7C15ull);
  uint32_t u = (uint32_t)(z >> 32);
  return (float)(u >> 8) * 5.9604644775390625e-8f;  // 2^-24
}
