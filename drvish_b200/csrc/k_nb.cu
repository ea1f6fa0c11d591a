// Gene-softmax + library-size scaling + negative-binomial log-likelihood and its backward on
// MATERIALISED decoder logits (general-shape path; the tcgen05 path fuses the same math into the
// GEMM epilogue).  Reference: NBDecoder.forward modules.py:93-96, the "obs" site nbvae.py:72-78,
// torch NegativeBinomial.log_prob negative_binomial.py:130-150; gradients SURVEY Appendix A.3.
// One CTA per cell: the 2G-float logits row and the G-float count row stay in L1/L2 between passes.
#include "kernels.cuh"

namespace drv {

namespace {

constexpr int NBT = 256;

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  int w = threadIdx.x / 32, l = threadIdx.x % 32;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (threadIdx.x < NBT / 32) ? sh[threadIdx.x] : 0.f;
  if (w == 0) t = warp_sum(t);
  if (threadIdx.x == 0) sh[0] = t;
  __syncthreads();
  return sh[0];
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  v = warp_max(v);
  int w = threadIdx.x / 32, l = threadIdx.x % 32;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (threadIdx.x < NBT / 32) ? sh[threadIdx.x] : -INFINITY;
  if (w == 0) t = warp_max(t);
  if (threadIdx.x == 0) sh[0] = t;
  __syncthreads();
  return sh[0];
}

template <bool BWD>
__global__ void __launch_bounds__(NBT) k_nb_rows(float* __restrict__ logits, const float* __restrict__ x,
                                                 const float* __restrict__ lib, const float* __restrict__ rshift, int G,
                                                 float eps, float c, float* __restrict__ lse_out,
                                                 float* __restrict__ ll_out, float* __restrict__ asum_out,
                                                 drv_step_out* out) {
  pdl_prologue();
  __shared__ float sh[NBT / 32];
  __shared__ float lgtab[LGTAB_N];
  load_lgamma_table(lgtab, threadIdx.x);
  __syncthreads();
  const int b = blockIdx.x;
  float* s = logits + (size_t)b * 2 * G;
  float* rho = s + G;
  const float* xr = x + (size_t)b * G;
  const float l = lib[b];
  const float rs = rshift ? rshift[b] * 1.4426950408889634f : 0.f;  // MCV: shift of log_r for the count parameter

  // pass 1: logsumexp over the scale half
  float mx = -INFINITY;
  for (int g = threadIdx.x; g < G; g += NBT) mx = fmaxf(mx, s[g]);
  mx = block_max(mx, sh);
  float se = 0.f;
  for (int g = threadIdx.x; g < G; g += NBT) se += __expf(s[g] - mx);
  se = block_sum(se, sh);
  const float lse = mx + logf(se);

  // pass 2: log-likelihood and sum_g dLL/d ell (warp-uniform loops: nb_elems needs converged
  // warps; each lane handles 4 genes per iteration for instruction-level parallelism)
  const int lane = threadIdx.x % 32, wbase = (threadIdx.x / 32) * 128;
  const float lml = l - lse;
  float ll = 0.f, asum = 0.f;
  for (int g0 = wbase; g0 < G; g0 += NBT * 4) {
    float a_lsl[4], a_rho[4], a_x[4];
    bool ok[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int g = g0 + 32 * v + lane;
      ok[v] = g < G;
      a_lsl[v] = ok[v] ? lml + s[g] : 0.f;
      a_rho[v] = ok[v] ? rho[g] : 0.f;
      a_x[v] = ok[v] ? xr[g] : 0.f;
    }
    NbElem e[4];
    nb_elems<BWD, 4>(a_lsl, a_rho, a_x, eps, lgtab, e, rs);
#pragma unroll
    for (int v = 0; v < 4; ++v)
      if (ok[v]) {
        ll += e[v].ll;
        if (BWD) asum += e[v].a;
      }
  }
  ll = block_sum(ll, sh);
  if (BWD) asum = block_sum(asum, sh);
  if (threadIdx.x == 0) {
    lse_out[b] = lse;
    ll_out[b] = ll;
    asum_out[b] = asum;
    atomicAdd(&out->loss, -(double)c * (double)ll);
  }
  if (!BWD) return;

  // pass 3: d loss / d logits in place (loss = -c * LL):  ds = -c (a - p A),  drho = -c dLL/drho
  for (int g0 = wbase; g0 < G; g0 += NBT * 4) {
    float a_lsl[4], a_rho[4], a_x[4], ls[4];
    bool ok[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int g = g0 + 32 * v + lane;
      ok[v] = g < G;
      ls[v] = ok[v] ? s[g] - lse : 0.f;
      a_lsl[v] = l + ls[v];
      a_rho[v] = ok[v] ? rho[g] : 0.f;
      a_x[v] = ok[v] ? xr[g] : 0.f;
    }
    NbElem e[4];
    nb_elems<true, 4>(a_lsl, a_rho, a_x, eps, lgtab, e, rs);
#pragma unroll
    for (int v = 0; v < 4; ++v)
      if (ok[v]) {
        const int g = g0 + 32 * v + lane;
        s[g] = -c * (e[v].a - __expf(ls[v]) * asum);
        rho[g] = -c * e[v].drho;
      }
  }
}

__global__ void k_nb_decoder_outputs(const float* __restrict__ logits, const float* __restrict__ lib, int G,
                                     float* __restrict__ log_r, float* __restrict__ logit) {
  pdl_prologue();
  __shared__ float sh[NBT / 32];
  const int b = blockIdx.x;
  const float* s = logits + (size_t)b * 2 * G;
  const float* rho = s + G;
  float mx = -INFINITY;
  for (int g = threadIdx.x; g < G; g += NBT) mx = fmaxf(mx, s[g]);
  mx = block_max(mx, sh);
  float se = 0.f;
  for (int g = threadIdx.x; g < G; g += NBT) se += expf(s[g] - mx);
  se = block_sum(se, sh);
  const float lse = mx + logf(se);
  const float l = lib[b];
  for (int g = threadIdx.x; g < G; g += NBT) {
    float r = rho[g];
    log_r[(size_t)b * G + g] = r;
    logit[(size_t)b * G + g] = l + (s[g] - lse) - r;
  }
}

}  // namespace

void nb_rows(float* logits, const float* x, const float* lib, const float* rshift, int B, int G, float eps,
             float scale_factor, bool bwd, float* lse, float* ll, float* asum, drv_step_out* out, cudaStream_t st) {
  if (bwd) launch_k(k_nb_rows<true>, B, NBT, 0, st, logits, x, lib, rshift, G, eps, scale_factor, lse, ll, asum, out);
  else launch_k(k_nb_rows<false>, B, NBT, 0, st, logits, x, lib, rshift, G, eps, scale_factor, lse, ll, asum, out);
  note_launch();
}

void nb_decoder_outputs(const float* logits, const float* lib, int B, int G, float* log_r, float* logit, cudaStream_t st) {
  launch_k(k_nb_decoder_outputs, B, NBT, 0, st, logits, lib, G, log_r, logit);
  note_launch();
}

}  // namespace drv
