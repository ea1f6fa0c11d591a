// Dose-response head: all drugs of DRNBVAE in a handful of launches instead of the reference's
// Python loop over drugs (drnbvae.py:110-116) with ~15 tiny ops and a sort-based unique() each.
// Reference: LinearMultiBias.logit_mean_sigmoid modules.py:159-180, the drs_i Normal likelihood
// drnbvae.py:112-116, the priors modules.py:130-135 and the guide-side monotonic factor
// modules.py:189-196; semantics as pinned down in SURVEY Appendix A.4.
// Layouts: a "slot" is one (drug i, dose c) pair, slot = dose_off[i] + c, T slots in total;
// per-drug outputs m_i (K, C_i, 1) are stored drug-major at K*dose_off[i] + k*C_i + c.
#include "kernels.cuh"

namespace drv {

namespace {

__device__ __forceinline__ int drug_of_slot(const int32_t* __restrict__ off, int D, int slot) {
  int lo = 0, hi = D - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (off[mid] <= slot) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

__global__ void k_dr_project(const float* __restrict__ z, const float* __restrict__ w, int B, int L, int D,
                             float* __restrict__ u) {
  pdl_prologue();
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * D) return;
  int b = idx / D, i = idx % D;
  float s = 0.f;
  for (int d = 0; d < L; ++d) s = __fmaf_rn(z[(size_t)b * L + d], w[(size_t)i * L + d], s);
  u[idx] = s;
}

__global__ void __launch_bounds__(256) k_dr_means(const float* __restrict__ u, const float* __restrict__ bias,
                                                  const int64_t* __restrict__ labels, const int32_t* __restrict__ off,
                                                  DrShape s, const float* __restrict__ y, float sigma,
                                                  float* __restrict__ m_out, float* __restrict__ gfac,
                                                  drv_step_out* out, uint32_t* status) {
  pdl_prologue();
  extern __shared__ float sh[];  // [K] sums, [K] counts
  float* acc = sh;
  float* cnt = sh + s.K;
  const int slot = blockIdx.x;
  const int i = drug_of_slot(off, s.D, slot);
  const int c = slot - off[i];
  const int Ci = off[i + 1] - off[i];
  for (int k = threadIdx.x; k < 2 * s.K; k += blockDim.x) sh[k] = 0.f;
  __syncthreads();
  const float bi = bias[slot];
  bool bad = false;
  if (s.K <= 16) {
    // few classes: per-thread sums selected by label (no dynamic indexing), one warp reduction per
    // class, one shared-memory atomic per warp and class - instead of 256 threads contending for K words
    float ra[16], ca[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) ra[k] = ca[k] = 0.f;
    for (int b0 = threadIdx.x; b0 < s.B; b0 += 4 * blockDim.x) {
      // four cells per trip: the eight loads are issued before any of them is used
      int64_t lab[4];
      float uv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int b = b0 + j * blockDim.x;
        lab[j] = b < s.B ? labels[b] : 0;
        uv[j] = b < s.B ? u[(size_t)b * s.D + i] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (b0 + j * (int)blockDim.x >= s.B) continue;
        if (lab[j] < 0 || lab[j] >= s.K) { bad = true; continue; }
        const float r = sigmoidf_(uv[j] + bi);
        const int li = (int)lab[j];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          ra[k] += li == k ? r : 0.f;
          ca[k] += li == k ? 1.f : 0.f;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (k < s.K) {
        const float tr = warp_sum(ra[k]), tc = warp_sum(ca[k]);
        if (threadIdx.x % 32 == 0) {
          atomicAdd(&acc[k], tr);
          atomicAdd(&cnt[k], tc);
        }
      }
    }
  } else {
    for (int b = threadIdx.x; b < s.B; b += blockDim.x) {
      int64_t lab = labels[b];
      if (lab < 0 || lab >= s.K) { bad = true; continue; }
      float r = sigmoidf_(u[(size_t)b * s.D + i] + bi);
      atomicAdd(&acc[lab], r);
      atomicAdd(&cnt[lab], 1.f);
    }
  }
  if (bad) atomicOr(status, DRV_STATUS_BAD_LABEL);
  __syncthreads();
  float nll = 0.f;
  for (int k = threadIdx.x; k < s.K; k += blockDim.x) {
    float n = cnt[k];
    if (n == 0.f) { atomicOr(status, DRV_STATUS_MISSING_CLASS); n = 1.f; }
    float mean = acc[k] / n;
    float mc = fminf(fmaxf(mean, 1e-5f), 1.f - 1e-5f);
    float m = logf(mc / (1.f - mc));  // torch.logit(., eps=1e-5)
    size_t oi = (size_t)s.K * off[i] + (size_t)k * Ci + c;
    m_out[oi] = m;
    if (y) {
      float d = y[oi] - m;
      nll -= -0.5f * d * d / (sigma * sigma) - logf(sigma) - DRV_HALF_LOG_2PI;
      bool inside = (mean >= 1e-5f) && (mean <= 1.f - 1e-5f);
      // d loss/d m = -(y - m)/sigma^2 ; d m/d mean = 1/(mc (1-mc)) ; d mean / d resp_b = 1/n
      gfac[(size_t)k * s.T + slot] = inside ? (-d / (sigma * sigma)) / (mc * (1.f - mc)) / n : 0.f;
    }
  }
  if (y) {
    nll = warp_sum(nll);
    if (threadIdx.x % 32 == 0 && nll != 0.f) atomicAdd(&out->loss, (double)nll);
  }
}

__global__ void k_dr_du(const float* __restrict__ u, const float* __restrict__ bias, const int64_t* __restrict__ labels,
                        const int32_t* __restrict__ off, const float* __restrict__ gfac, DrShape s,
                        float* __restrict__ du) {
  pdl_prologue();
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= s.B * s.D) return;
  int b = idx / s.D, i = idx % s.D;
  int64_t lab = labels[b];
  float g = 0.f;
  if (lab >= 0 && lab < s.K) {
    float uv = u[idx];
    for (int slot = off[i]; slot < off[i + 1]; ++slot) {
      float r = sigmoidf_(uv + bias[slot]);
      g = __fmaf_rn(gfac[(size_t)lab * s.T + slot], r * (1.f - r), g);
    }
  }
  du[idx] = g;
}

__global__ void k_dr_dz(const float* __restrict__ du, const float* __restrict__ w, int B, int L, int D,
                        float* __restrict__ dz) {
  pdl_prologue();
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * L) return;
  int b = idx / L, d = idx % L;
  float sacc = 0.f;
  for (int i = 0; i < D; ++i) sacc = __fmaf_rn(du[(size_t)b * D + i], w[(size_t)i * L + d], sacc);
  dz[idx] = sacc;
}

__global__ void __launch_bounds__(256) k_dr_prior_factor(const float* __restrict__ w_prior, const float* __restrict__ b_prior,
                                                         const float* __restrict__ eps_bias,
                                                         const float* __restrict__ bias_loc,
                                                         const float* __restrict__ bias_scale_unc,
                                                         const int32_t* __restrict__ off, DrShape s, float lam, float bsc,
                                                         float alpha, int include_factor, float* __restrict__ d_bias_loc,
                                                         float* __restrict__ d_bias_scale_unc, drv_step_out* out) {
  pdl_prologue();
  double t = 0.0;  // contribution to the LOSS (= -ELBO)
  for (int j = threadIdx.x; j < s.D * s.L; j += blockDim.x)
    t -= (double)(-logf(2.f * lam) - fabsf(w_prior[j]) / lam);  // Laplace(0, lam).log_prob
  for (int j = threadIdx.x; j < s.T; j += blockDim.x) {
    float b = b_prior[j];
    t -= (double)(-0.5f * b * b / (bsc * bsc) - logf(bsc) - DRV_HALF_LOG_2PI);  // Normal(0, bsc).log_prob
  }
  for (int slot = threadIdx.x; slot < s.T; slot += blockDim.x) {
    float gb = 0.f;
    if (include_factor) {
      int i = drug_of_slot(off, s.D, slot);
      int c = slot - off[i], Ci = off[i + 1] - off[i];
      auto bias_at = [&](int sl) { return __fmaf_rn(expf(bias_scale_unc[sl]), eps_bias[sl], bias_loc[sl]); };
      float here = bias_at(slot);
      if (c < Ci - 1) {
        float d = here - bias_at(slot + 1);
        if (d > 0.f) {
          // guide-side factor log f = -alpha sum relu(d): ELBO -= log f, loss += log f
          t += (double)(-alpha * d);
          gb -= alpha;
        }
      }
      if (c > 0) {
        float d = bias_at(slot - 1) - here;
        if (d > 0.f) gb += alpha;
      }
    }
    if (d_bias_loc) {
      d_bias_loc[slot] = gb;
      d_bias_scale_unc[slot] = gb * eps_bias[slot] * expf(bias_scale_unc[slot]);
    }
  }
  t = warp_sum(t);
  __shared__ double shd[8];
  if (threadIdx.x % 32 == 0) shd[threadIdx.x / 32] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int k = 0; k < 8; ++k) tot += shd[k];
    atomicAdd(&out->loss, tot);
  }
}

}  // namespace

void dr_project(const float* z, const float* w, DrShape s, float* u, cudaStream_t st) {
  launch_k(k_dr_project, cdiv((int64_t)s.B * s.D, 256), 256, 0, st, z, w, s.B, s.L, s.D, u);
  note_launch();
}

void dr_means(const float* u, const float* bias, const int64_t* labels, const int32_t* dose_off, DrShape s,
              const float* y, float sigma, float* m, float* gfac, drv_step_out* out, uint32_t* status, cudaStream_t st) {
  launch_k(k_dr_means, s.T, 256, 2 * s.K * sizeof(float), st, u, bias, labels, dose_off, s, y, sigma, m, gfac, out, status);
  note_launch();
}

void dr_backward(const float* u, const float* bias, const float* w, const int64_t* labels, const int32_t* dose_off,
                 const float* gfac, DrShape s, float* du, float* dz, cudaStream_t st) {
  launch_k(k_dr_du, cdiv((int64_t)s.B * s.D, 256), 256, 0, st, u, bias, labels, dose_off, gfac, s, du);
  note_launch();
  launch_k(k_dr_dz, cdiv((int64_t)s.B * s.L, 256), 256, 0, st, du, w, s.B, s.L, s.D, dz);
  note_launch();
}

void dr_prior_factor(const float* w_prior, const float* b_prior, const float* eps_bias, const float* bias_loc,
                     const float* bias_scale_unc, const int32_t* dose_off, DrShape s, const drv_hparams& hp,
                     float* d_bias_loc, float* d_bias_scale_unc, drv_step_out* out, cudaStream_t st) {
  launch_k(k_dr_prior_factor, 1, 256, 0, st, w_prior, b_prior, eps_bias, bias_loc, bias_scale_unc, dose_off, s, hp.lam_scale,
                                       hp.bias_scale, hp.alpha, hp.include_guide_factor, d_bias_loc, d_bias_scale_unc,
                                       out);
  note_launch();
}

}  // namespace drv
