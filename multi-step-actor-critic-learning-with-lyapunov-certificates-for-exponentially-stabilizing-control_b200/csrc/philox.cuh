// Philox4x32-10 counter RNG and the float transforms used for action noise and resets.
// Replaces the reference's non-reproducible host streams (torch global generator in
// RL/utils/act_distribution_cls.py:45-47; per-env PCG64 reseeded from Python `random` in
// e.g. RL/env/VanderPol.py:72-81).  Counter layout is documented in oracle/philox.py, which
// restates this file bit-for-bit on the integer side.
#pragma once
#include <cstdint>

namespace msacl {

constexpr uint32_t kStreamNoise = 0;
constexpr uint32_t kStreamReset = 1;

struct U4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                     uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return U4{c0, c1, c2, c3};
}

__device__ __forceinline__ U4 philox_env(uint64_t seed, uint64_t env, uint32_t index, uint32_t stream) {
  return philox4x32_10((uint32_t)env, (uint32_t)(env >> 32), index, stream, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// top 24 bits -> [0,1), exact in float32
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }

// Box-Muller: r = sqrt(-2 ln u1), u1 in (0,1]; angle 2*pi*u2 via sincospif(2*u2)
__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float& z0, float& z1) {
  const float u1 = ((float)(xa >> 8) + 1.0f) * 5.9604644775390625e-08f;
  const float u2 = u01(xb);
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// up to 4 N(0,1) draws for (env, step)
__device__ __forceinline__ void action_noise4(uint64_t seed, uint64_t env, uint32_t step, float (&z)[4]) {
  const U4 r = philox_env(seed, env, step, kStreamNoise);
  box_muller(r.x, r.y, z[0], z[1]);
  box_muller(r.z, r.w, z[2], z[3]);
}

}  // namespace msacl
