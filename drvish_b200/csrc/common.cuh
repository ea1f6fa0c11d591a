// Shared device helpers for the drvish_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/drvish_b200.h"

#define DRV_HALF_LOG_2PI 0.91893853320467274178f

namespace drv {

// launch bookkeeping (host side, per thread) -------------------------------------------------
extern thread_local int g_launches;
extern thread_local cudaError_t g_err;

inline void note_launch() {
  ++g_launches;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess && g_err == cudaSuccess) g_err = e;
}

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// BatchNorm -> ReLU applied to one element (modules.py:19-20).  Explicit fma so that the
// forward value and the recomputed ReLU gate of the backward are bit-identical.
struct BnParams {
  const float* mean;   // batch mean per column
  const float* rstd;   // 1/sqrt(biased var + eps)
  const float* gamma;  // BatchNorm weight
  const float* beta;   // BatchNorm bias
};

__device__ __forceinline__ float bn_xhat(float u, float mean, float rstd) { return (u - mean) * rstd; }
__device__ __forceinline__ float bn_y(float xhat, float gamma, float beta) { return __fmaf_rn(xhat, gamma, beta); }
__device__ __forceinline__ float bn_relu(float u, float mean, float rstd, float gamma, float beta) {
  return fmaxf(bn_y(bn_xhat(u, mean, rstd), gamma, beta), 0.f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// special functions
// ---------------------------------------------------------------------------------------------
// digamma for y > 0: recurrence up to y >= 6, then the asymptotic series (next term 1/(240 y^8)).
__device__ __forceinline__ float digamma_pos(float y) {
  float acc = 0.f;
  while (y < 6.f) {
    acc -= 1.f / y;
    y += 1.f;
  }
  float inv = 1.f / y, inv2 = inv * inv;
  float series = inv2 * (0.083333333333f - inv2 * (0.0083333333333f - inv2 * 0.003968253968f));
  return acc + logf(y) - 0.5f * inv - series;
}

// lgamma(1 + x) for a count x >= 0.  Integer counts below 32 come from a table.
static __constant__ float c_lgamma1p[32] = {
    0.f, 0.f, 0.69314718055994531f, 1.79175946922805500f, 3.17805383034794562f, 4.78749174278204599f,
    6.57925121201010100f, 8.52516136106541430f, 10.6046029027452502f, 12.8018274800814696f,
    15.1044125730755153f, 17.5023078458738858f, 19.9872144956618861f, 22.5521638531234229f,
    25.1912211827386815f, 27.8992713838408916f, 30.6718601060806728f, 33.5050734501368889f,
    36.3954452080330536f, 39.3398841871994940f, 42.3356164607534850f, 45.3801388984769080f,
    48.4711813518352239f, 51.6066755677643736f, 54.7847293981123192f, 58.0036052229805199f,
    61.2617017610020020f, 64.5575386270063311f, 67.8897431371815350f, 71.2570389671680090f,
    74.6582363488301644f, 78.0922235533153106f};

__device__ __forceinline__ float lgamma1p_count(float x) {
  int xi = (int)x;
  if (x < 32.f && (float)xi == x) return c_lgamma1p[xi];
  return lgammaf(1.f + x);
}

// Result of one (cell, gene) element of the NB block.
struct NbElem {
  float ll;    // log NB(x | r, logits=ell)
  float a;     // dLL/d ell = x - (x + r) sigmoid(ell)
  float drho;  // dLL/d rho = e^rho (log sigmoid(-ell) + psi(r+x) - psi(r)) - a
};

// ls = s - logsumexp(s) (log softmax of the scale half), rho = log_r, l = library, x = count.
// Follows torch.distributions.NegativeBinomial.log_prob (negative_binomial.py:130-150):
//   r logsig(-ell) + x logsig(ell) + lgamma(r+x) - lgamma(1+x) - lgamma(r)
// with the lgamma / digamma differences formed directly (exact finite products for small
// integer counts), which is at least as accurate as torch's fp32 evaluation.
template <bool BWD>
__device__ __forceinline__ NbElem nb_elem(float ls, float rho, float l, float x, float eps) {
  NbElem o;
  float rexp = __expf(rho);
  float r = rexp + eps;
  float ell = l + ls - rho;
  float e = __expf(-fabsf(ell));
  float lp = log1pf(e);
  float sp = fmaxf(ell, 0.f) + lp;  // softplus(ell) = -logsigmoid(-ell)
  float lgd, psd = 0.f;
  int xi = (int)x;
  if (x == 0.f) {
    lgd = 0.f;
  } else if (x <= 8.f && (float)xi == x && r < 1.0e4f) {
    // lgamma(r+x) - lgamma(r) = log prod_{k<x} (r+k);  psi(r+x) - psi(r) = P'/P
    float P = 1.f, Pd = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < xi) {
        float t = r + (float)k;
        if (BWD) Pd = __fmaf_rn(Pd, t, P);
        P *= t;
      }
    }
    lgd = __logf(P);
    if (BWD) psd = __fdividef(Pd, P);
  } else {
    lgd = lgammaf(r + x) - lgammaf(r);
    if (BWD) psd = digamma_pos(r + x) - digamma_pos(r);
  }
  o.ll = x * ell - (x + r) * sp + lgd - lgamma1p_count(x);
  if (BWD) {
    float inv = __fdividef(1.f, 1.f + e);
    float sig = ell >= 0.f ? inv : e * inv;
    o.a = x - (x + r) * sig;
    o.drho = rexp * (psd - sp) - o.a;
  } else {
    o.a = 0.f;
    o.drho = 0.f;
  }
  return o;
}

// softplus / its derivative exactly as torch.nn.functional.softplus (beta=1, threshold=20)
__device__ __forceinline__ float softplus_t(float v) { return v > 20.f ? v : log1pf(expf(v)); }
__device__ __forceinline__ float softplus_grad_t(float v) {
  if (v > 20.f) return 1.f;
  float z = expf(v);
  return z / (z + 1.f);
}

}  // namespace drv
