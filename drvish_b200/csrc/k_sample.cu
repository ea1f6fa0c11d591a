// Reparameterised sampling of the "latent" / "library" sites, their guide and prior log-probs
// (the KL-like part of the ELBO) and the matching backward.  Reference: Encoder.forward
// modules.py:56-60 (column order [z_loc(L), l_loc, z_raw_scale(L), l_raw_scale], softplus),
// NBVAE.guide nbvae.py:89-90 (rsample), NBVAE.model nbvae.py:62-68 (priors), SURVEY Appendix A.2.
// Also the counter-based noise generators used in production mode.
#include "kernels.cuh"
#include "philox.cuh"

namespace drv {

namespace {

constexpr int ST = 128;

__global__ void __launch_bounds__(ST) k_sample_fwd(const float* __restrict__ O, const float* __restrict__ eps_z,
                                                   const float* __restrict__ eps_l, int B, int L, float c,
                                                   float lib_loc, float lib_scale, float* __restrict__ z,
                                                   float* __restrict__ l, drv_step_out* out) {
  pdl_prologue();
  int b = blockIdx.x * ST + threadIdx.x;
  const int P = 2 * L + 2;
  float t = 0.f;  // log p - log q of this cell
  if (b < B) {
    const float* o = O + (size_t)b * P;
    for (int d = 0; d < L; ++d) {
      float loc = o[d], sc = softplus_t(o[L + 1 + d]), e = eps_z[(size_t)b * L + d];
      float zv = __fmaf_rn(e, sc, loc);
      z[(size_t)b * L + d] = zv;
      // Normal(0,1).log_prob(z) - Normal(loc, sc).log_prob(z); (z - loc)/sc == e up to rounding
      t += (-0.5f * zv * zv) - (-0.5f * e * e - logf(sc));
    }
    float loc = o[L], sc = softplus_t(o[2 * L + 1]), e = eps_l[b];
    float lv = __fmaf_rn(e, sc, loc);
    l[b] = lv;
    float dl = (lv - lib_loc) / lib_scale;
    t += (-0.5f * dl * dl - logf(lib_scale)) - (-0.5f * e * e - logf(sc));
  }
  t = warp_sum(t);
  __shared__ float sh[ST / 32];
  if (threadIdx.x % 32 == 0) sh[threadIdx.x / 32] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < ST / 32; ++i) s += sh[i];
    atomicAdd(&out->loss, -(double)c * (double)s);
  }
}

// d loss / d (encoder output), loss = -c sum_b (log p_z + log p_l + LL - log q_z - log q_l) + drs terms
__global__ void __launch_bounds__(ST) k_sample_bwd(const float* __restrict__ O, const float* __restrict__ eps_z,
                                                   const float* __restrict__ eps_l, const float* __restrict__ z,
                                                   const float* __restrict__ l, const float* __restrict__ dz_dec,
                                                   const float* __restrict__ dz_drs, const float* __restrict__ asum,
                                                   int B, int L, float c, float lib_loc, float lib_scale,
                                                   float* __restrict__ dO) {
  pdl_prologue();
  int b = blockIdx.x * ST + threadIdx.x;
  if (b >= B) return;
  const int P = 2 * L + 2;
  const float* o = O + (size_t)b * P;
  float* g = dO + (size_t)b * P;
  for (int d = 0; d < L; ++d) {
    size_t i = (size_t)b * L + d;
    float dz = c * z[i] + dz_dec[i] + (dz_drs ? dz_drs[i] : 0.f);
    float raw = o[L + 1 + d];
    float sc = softplus_t(raw);
    g[d] = dz;
    g[L + 1 + d] = (dz * eps_z[i] - c / sc) * softplus_grad_t(raw);
  }
  // dLL/dl = A[b]   -> d loss/dl = -c A + c (l - mu)/sigma^2
  float dl = -c * asum[b] + c * (l[b] - lib_loc) / (lib_scale * lib_scale);
  float raw = o[2 * L + 1];
  float sc = softplus_t(raw);
  g[L] = dl;
  g[2 * L + 1] = (dl * eps_l[b] - c / sc) * softplus_grad_t(raw);
}

__global__ void k_encoder_outputs(const float* __restrict__ O, int B, int L, float* __restrict__ z_loc,
                                  float* __restrict__ z_scale, float* __restrict__ l_loc, float* __restrict__ l_scale) {
  pdl_prologue();
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * (L + 1)) return;
  int b = idx / (L + 1), d = idx % (L + 1);
  const float* o = O + (size_t)b * (2 * L + 2);
  if (d < L) {
    z_loc[(size_t)b * L + d] = o[d];
    z_scale[(size_t)b * L + d] = softplus_t(o[L + 1 + d]);
  } else {
    l_loc[b] = o[L];
    l_scale[b] = softplus_t(o[2 * L + 1]);
  }
}

// The Philox counter is offset + step_counter[0] * stride: the step counter lives in device memory
// so that a captured CUDA graph draws fresh noise on every replay.
template <bool LAPLACE>
__global__ void k_fill(float* __restrict__ out, int64_t n, uint64_t seed, uint64_t offset,
                       const unsigned long long* __restrict__ step_counter, uint64_t stride) {
  pdl_prologue();
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // quad index
  if (q * 4 >= n) return;
  if (step_counter) offset += (uint64_t)(*step_counter) * stride;
  uint32_t r[4];
  philox4(seed, offset / 4 + (uint64_t)q, r);
  float v[4];
  if (LAPLACE) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float u = 2.f * u01(r[i]) - 1.f;  // (-1,1); Laplace icdf as torch laplace.py:73-85
      v[i] = -copysignf(1.f, u) * log1pf(-fabsf(u));
    }
  } else {
    // Box-Muller on two pairs
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float rad = sqrtf(-2.f * logf(u01(r[2 * i])));
      float s, co;
      sincospif(2.f * u01(r[2 * i + 1]), &s, &co);
      v[2 * i] = rad * co;
      v[2 * i + 1] = rad * s;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (q * 4 + i < n) out[q * 4 + i] = v[i];
}

// All noise draws of one ELBO step in ONE launch (the five drv_fill_* calls + the prior scalings +
// the counter bump).  Segment i covers blocks [blk0[i], blk0[i+1]); its values are exactly those of
// the per-buffer generators at the same Philox offsets.  The last CTA to finish advances the step
// counter (every CTA has read it before taking its ticket) and re-arms the ticket.
struct NoiseSegs {
  float* out[5];
  long long n[5];
  unsigned long long off[5];
  float scale[5];
  int laplace[5];
  int blk0[6];
};

__device__ __forceinline__ void draw4(bool laplace, uint64_t seed, uint64_t ctr, float (&v)[4]) {
  uint32_t r[4];
  philox4(seed, ctr, r);
  if (laplace) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float u = 2.f * u01(r[i]) - 1.f;
      v[i] = -copysignf(1.f, u) * log1pf(-fabsf(u));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float rad = sqrtf(-2.f * logf(u01(r[2 * i])));
      float s, co;
      sincospif(2.f * u01(r[2 * i + 1]), &s, &co);
      v[2 * i] = rad * co;
      v[2 * i + 1] = rad * s;
    }
  }
}

__global__ void __launch_bounds__(256) k_draw_step_noise(NoiseSegs sg, uint64_t seed, unsigned long long* state,
                                                         uint64_t stride) {
  pdl_prologue();
  int seg = 0;
#pragma unroll
  for (int i = 1; i < 5; ++i)
    if ((int)blockIdx.x >= sg.blk0[i]) seg = i;
  const unsigned long long step = *(volatile unsigned long long*)state;
  int64_t q = (int64_t)(blockIdx.x - sg.blk0[seg]) * blockDim.x + threadIdx.x;
  int64_t n = sg.n[seg];
  if (q * 4 < n) {
    float v[4];
    draw4(sg.laplace[seg] != 0, seed, (sg.off[seg] + step * stride) / 4 + (uint64_t)q, v);
    float* out = sg.out[seg];
    float sc = sg.scale[seg];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (q * 4 + i < n) out[q * 4 + i] = v[i] * sc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int* ticket = reinterpret_cast<unsigned int*>(state + 1);
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
      *ticket = 0u;
      *state = step + 1;
    }
  }
}

}  // namespace

void draw_step_noise(float* const* outs, const int64_t* ns, const int* laplace, const float* scales, uint64_t seed,
                     uint64_t offset, unsigned long long* state, uint64_t stride, cudaStream_t st) {
  NoiseSegs sg{};
  int blk = 0;
  uint64_t off = offset;
  for (int i = 0; i < 5; ++i) {
    int64_t n = outs[i] ? ns[i] : 0;
    sg.out[i] = outs[i];
    sg.n[i] = n;
    sg.off[i] = off;
    sg.scale[i] = scales[i];
    sg.laplace[i] = laplace[i];
    sg.blk0[i] = blk;
    blk += (int)cdiv(cdiv(n, 4), 256);
    off += 4 * (uint64_t)cdiv(n, 4);
  }
  sg.blk0[5] = blk;
  if (blk == 0) blk = 1;
  launch_k(k_draw_step_noise, blk, 256, 0, st, sg, seed, state, stride);
  note_launch();
}

void sample_forward(const float* O, const float* eps_z, const float* eps_l, int B, int L, const drv_hparams& hp,
                    float* z, float* l, drv_step_out* out, cudaStream_t st) {
  launch_k(k_sample_fwd, cdiv(B, ST), ST, 0, st, O, eps_z, eps_l, B, L, hp.scale_factor, hp.lib_loc, hp.lib_scale, z, l, out);
  note_launch();
}

void sample_backward(const float* O, const float* eps_z, const float* eps_l, const float* z, const float* l,
                     const float* dz_dec, const float* dz_drs, const float* asum, int B, int L, const drv_hparams& hp,
                     float* dO, cudaStream_t st) {
  launch_k(k_sample_bwd, cdiv(B, ST), ST, 0, st, O, eps_z, eps_l, z, l, dz_dec, dz_drs, asum, B, L, hp.scale_factor,
                                           hp.lib_loc, hp.lib_scale, dO);
  note_launch();
}

void encoder_outputs(const float* O, int B, int L, float* z_loc, float* z_scale, float* l_loc, float* l_scale,
                     cudaStream_t st) {
  launch_k(k_encoder_outputs, cdiv((int64_t)B * (L + 1), 256), 256, 0, st, O, B, L, z_loc, z_scale, l_loc, l_scale);
  note_launch();
}

void fill_normal(float* out, int64_t n, uint64_t seed, uint64_t offset, const unsigned long long* ctr, uint64_t stride,
                 cudaStream_t st) {
  if (n <= 0) return;
  launch_k(k_fill<false>, cdiv(cdiv(n, 4), 256), 256, 0, st, out, n, seed, offset, ctr, stride);
  note_launch();
}
void fill_laplace(float* out, int64_t n, uint64_t seed, uint64_t offset, const unsigned long long* ctr, uint64_t stride,
                  cudaStream_t st) {
  if (n <= 0) return;
  launch_k(k_fill<true>, cdiv(cdiv(n, 4), 256), 256, 0, st, out, n, seed, offset, ctr, stride);
  note_launch();
}

__global__ void k_axpby(const float* __restrict__ a, const float* __restrict__ b, float ca, float cb, int n,
                        float* __restrict__ out) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = ca * a[i] + cb * b[i];
}
void axpby(const float* a, const float* b, float ca, float cb, int n, float* out, cudaStream_t st) {
  launch_k(k_axpby, cdiv(n, 256), 256, 0, st, a, b, ca, cb, n, out);
  note_launch();
}

__global__ void k_counter_add(unsigned long long* c, unsigned long long v) {
  pdl_prologue(); *c += v; }
void counter_add(unsigned long long* ctr, unsigned long long v, cudaStream_t st) {
  launch_k(k_counter_add, 1, 1, 0, st, ctr, v);
  note_launch();
}

}  // namespace drv
