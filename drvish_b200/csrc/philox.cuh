// Philox4x32-10 counter-based generator (Salmon et al. 2011), shared by the noise kernels
// (k_sample.cu) and the molecule-splitting loader kernel (k_split.cu).
#pragma once
#include <stdint.h>

namespace drv {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  k[0] += 0x9E3779B9u;
  k[1] += 0xBB67AE85u;
}

// counter = (ctr, stream word c2, sub-counter c3), key = seed
__device__ __forceinline__ void philox4(uint64_t seed, uint64_t ctr, uint32_t (&r)[4], uint32_t c2 = 0x6472766Bu,
                                        uint32_t c3 = 0u) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), c2, c3};
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
  for (int i = 0; i < 10; ++i) philox_round(c, k);
  r[0] = c[0]; r[1] = c[1]; r[2] = c[2]; r[3] = c[3];
}

__device__ __forceinline__ float u01(uint32_t v) { return ((float)(v >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0,1)

}  // namespace drv
