// Fused Aggregated-Momentum optimizer step over the flat parameter / gradient buffers
// (SURVEY 8f row 1).  Reference: AggMo.step, src/drvish/train/aggmo.py:42-79 (lr already divided by
// len(betas) at construction, :22), run once PER PARAMETER TENSOR by Pyro's PyroOptim after its
// per-tensor gradient clipping (clip_args={"clip_norm": 20.0}, notebook cell 10:
// torch.nn.utils.clip_grad_norm_(p, clip_norm) on each parameter separately).  The reference spends
// ~10 elementwise launches per tensor (~250 per step at 2 hidden layers); here the whole step is
// two launches: per-tensor squared gradient norms, then clip + weight decay + all momentum buffers
// + (Nesterov) update in one pass (reads g, p and the n_betas buffers once, writes p and the buffers).
#include "kernels.cuh"

namespace drv {

namespace {

constexpr int OPT_T = 256;            // threads per CTA
constexpr int OPT_PER_CTA = 4096;     // floats per CTA (4 float4 per thread)

struct TensorTable {
  int n;
  int blk0[DRV_MAX_TENSORS + 1];  // first CTA of every tensor (prefix sum of its chunk count)
  long long off[DRV_MAX_TENSORS];
  long long len[DRV_MAX_TENSORS];
};

// CTA -> (tensor, chunk): binary search in the block prefix (n <= 96)
__device__ __forceinline__ int tensor_of_block(const TensorTable& tb, int blk) {
  int lo = 0, hi = tb.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tb.blk0[mid] <= blk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// norms[t] += sum of squares of tensor t's gradient (fp64 accumulation)
__global__ void __launch_bounds__(OPT_T) k_grad_sqnorm(const float* __restrict__ g, TensorTable tb, double* __restrict__ norms) {
  pdl_prologue();
  const int t = tensor_of_block(tb, blockIdx.x);
  const long long len = tb.len[t];
  const long long c0 = (long long)(blockIdx.x - tb.blk0[t]) * OPT_PER_CTA;
  const float* gp = g + tb.off[t];
  double s = 0.0;
  const long long c1 = c0 + OPT_PER_CTA < len ? c0 + OPT_PER_CTA : len;
  for (long long i = c0 + threadIdx.x * 4; i < c1; i += OPT_T * 4) {
    if (i + 4 <= c1) {  // tensors start on 128-byte boundaries: aligned
      const float4 v = *reinterpret_cast<const float4*>(gp + i);
      s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    } else {
      for (long long j = i; j < c1; ++j) s += (double)gp[j] * gp[j];
    }
  }
  s = warp_sum(s);
  __shared__ double sh[OPT_T / 32];
  if (threadIdx.x % 32 == 0) sh[threadIdx.x / 32] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < OPT_T / 32; ++i) tot += sh[i];
    atomicAdd(&norms[t], tot);
  }
}

__device__ __forceinline__ float aggmo_elem(float p, float g, float coef, float wd, float lr, int n_betas, const float* betas,
                                            bool nesterov, float* const* bufs, long long idx) {
  // d_p = clipped grad (+ weight decay); for each beta: buf = beta buf + d_p; p -= lr (nesterov ? d_p + beta buf : buf)
  float d = g * coef;
  if (wd != 0.f) d = __fmaf_rn(wd, p, d);
#pragma unroll
  for (int b = 0; b < DRV_MAX_BETAS; ++b) {
    if (b < n_betas) {
      float m = __fmaf_rn(betas[b], bufs[b][idx], d);
      bufs[b][idx] = m;
      const float upd = nesterov ? __fmaf_rn(betas[b], m, d) : m;
      p = __fmaf_rn(-lr, upd, p);
    }
  }
  return p;
}

__global__ void __launch_bounds__(OPT_T) k_aggmo(float* __restrict__ params, const float* __restrict__ grads,
                                                 float* __restrict__ mom, long long mom_stride, TensorTable tb,
                                                 const double* __restrict__ norms, drv_aggmo hp) {
  pdl_prologue();
  const int t = tensor_of_block(tb, blockIdx.x);
  const long long len = tb.len[t];
  const long long c0 = (long long)(blockIdx.x - tb.blk0[t]) * OPT_PER_CTA;
  const long long base = tb.off[t];
  float coef = 1.f;
  if (hp.clip_norm > 0.f) {
    // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
    const float total = (float)sqrt(norms[t]);
    coef = fminf(hp.clip_norm / (total + 1e-6f), 1.f);
  }
  float* bufs[DRV_MAX_BETAS];
  float betas[DRV_MAX_BETAS];
#pragma unroll
  for (int b = 0; b < DRV_MAX_BETAS; ++b) {
    bufs[b] = mom + (long long)b * mom_stride + base;
    betas[b] = hp.betas[b];
  }
  const long long c1 = c0 + OPT_PER_CTA < len ? c0 + OPT_PER_CTA : len;
  const bool nest = hp.nesterov != 0;
  for (long long i = c0 + threadIdx.x * 4; i < c1; i += OPT_T * 4) {
    if (i + 4 <= c1) {  // tensors start on 128-byte boundaries: 128-bit accesses throughout
      float4 p = *reinterpret_cast<float4*>(params + base + i);
      const float4 g = *reinterpret_cast<const float4*>(grads + base + i);
      float pv[4] = {p.x, p.y, p.z, p.w};
      float d[4] = {g.x * coef, g.y * coef, g.z * coef, g.w * coef};
      if (hp.weight_decay != 0.f) {
#pragma unroll
        for (int j = 0; j < 4; ++j) d[j] = __fmaf_rn(hp.weight_decay, pv[j], d[j]);
      }
#pragma unroll
      for (int b = 0; b < DRV_MAX_BETAS; ++b) {
        if (b < hp.n_betas) {
          float4 m = *reinterpret_cast<float4*>(bufs[b] + i);
          float mv[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            mv[j] = __fmaf_rn(betas[b], mv[j], d[j]);
            const float upd = nest ? __fmaf_rn(betas[b], mv[j], d[j]) : mv[j];
            pv[j] = __fmaf_rn(-hp.lr, upd, pv[j]);
          }
          *reinterpret_cast<float4*>(bufs[b] + i) = make_float4(mv[0], mv[1], mv[2], mv[3]);
        }
      }
      *reinterpret_cast<float4*>(params + base + i) = make_float4(pv[0], pv[1], pv[2], pv[3]);
    } else {
      for (long long j = i; j < c1; ++j)
        params[base + j] = aggmo_elem(params[base + j], grads[base + j], coef, hp.weight_decay, hp.lr, hp.n_betas, betas,
                                      nest, bufs, j);
    }
  }
}

// Dose-response parameters: one small tensor per drug (weight_loc / weight_scale: L floats, bias_loc /
// bias_scale: C_i floats - each its own pyro.param, so each clipped on its own).  One CTA per tensor:
// norm and update in the same launch.  which = blockIdx.y: 0 w_loc, 1 w_scale, 2 b_loc, 3 b_scale.
struct DrRegions {
  long long base[4];
  int L;
};
__global__ void __launch_bounds__(128) k_aggmo_small(float* __restrict__ params, const float* __restrict__ grads,
                                                     float* __restrict__ mom, long long mom_stride, DrRegions rg,
                                                     const int32_t* __restrict__ dose_off, drv_aggmo hp) {
  pdl_prologue();
  const int drug = blockIdx.x, which = blockIdx.y;
  long long off;
  int len;
  if (which < 2) { off = rg.base[which] + (long long)drug * rg.L; len = rg.L; }
  else { off = rg.base[which] + dose_off[drug]; len = dose_off[drug + 1] - dose_off[drug]; }
  double s = 0.0;
  for (int i = threadIdx.x; i < len; i += blockDim.x) s += (double)grads[off + i] * grads[off + i];
  s = warp_sum(s);
  __shared__ double sh[4];
  if (threadIdx.x % 32 == 0) sh[threadIdx.x / 32] = s;
  __syncthreads();
  float coef = 1.f;
  if (hp.clip_norm > 0.f) {
    const float total = (float)sqrt(sh[0] + sh[1] + sh[2] + sh[3]);
    coef = fminf(hp.clip_norm / (total + 1e-6f), 1.f);
  }
  float* bufs[DRV_MAX_BETAS];
  float betas[DRV_MAX_BETAS];
#pragma unroll
  for (int b = 0; b < DRV_MAX_BETAS; ++b) {
    bufs[b] = mom + (long long)b * mom_stride + off;
    betas[b] = hp.betas[b];
  }
  for (int i = threadIdx.x; i < len; i += blockDim.x)
    params[off + i] = aggmo_elem(params[off + i], grads[off + i], coef, hp.weight_decay, hp.lr, hp.n_betas, betas,
                                 hp.nesterov != 0, bufs, i);
}

}  // namespace

void aggmo_step_dose_response(const long long* bases, int L, int n_drugs, const int32_t* dose_off, const drv_aggmo& hp,
                              float* params, const float* grads, float* momentum, long long mom_stride, cudaStream_t st) {
  DrRegions rg{};
  for (int i = 0; i < 4; ++i) rg.base[i] = bases[i];
  rg.L = L;
  launch_k(k_aggmo_small, dim3((unsigned)n_drugs, 4), 128, 0, st, params, grads, momentum, mom_stride, rg, dose_off, hp);
  note_launch();
}

int aggmo_step(const long long* offs, const long long* lens, int n_tensors, const drv_aggmo& hp, float* params,
               const float* grads, float* momentum, long long mom_stride, double* norms, cudaStream_t st) {
  if (n_tensors < 1 || n_tensors > DRV_MAX_TENSORS) return DRV_EINVAL;
  TensorTable tb{};
  tb.n = n_tensors;
  int blocks = 0;
  for (int i = 0; i < n_tensors; ++i) {
    tb.off[i] = offs[i];
    tb.len[i] = lens[i];
    tb.blk0[i] = blocks;
    blocks += cdiv(lens[i], OPT_PER_CTA);
  }
  tb.blk0[n_tensors] = blocks;
  dim3 grid((unsigned)blocks);
  if (hp.clip_norm > 0.f) {
    cudaMemsetAsync(norms, 0, sizeof(double) * n_tensors, st);
    launch_k(k_grad_sqnorm, grid, OPT_T, 0, st, grads, tb, norms);
    note_launch();
  }
  launch_k(k_aggmo, grid, OPT_T, 0, st, params, grads, momentum, mom_stride, tb, (const double*)norms, hp);
  note_launch();
  return 0;
}

}  // namespace drv
