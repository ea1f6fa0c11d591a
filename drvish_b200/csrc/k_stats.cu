// BatchNorm1d (training mode) statistics, running-buffer update and backward
// (reference modules.py:19: every Linear of make_fc is preceded by BatchNorm1d -> ReLU; the
// reference never calls .eval(), so batch statistics are used in validation too - SURVEY 3.3).
// HBM-bound column reductions: 128-byte coalesced rows, fp64 accumulators, the batch split over
// enough CTAs to fill the 148 SMs, one double atomic per column per CTA.
#include "kernels.cuh"

namespace drv {

namespace {

constexpr int CW = 32;   // columns per CTA (one 128-byte line per row)
constexpr int RY = 8;    // row lanes per CTA

inline int row_splits(int B, int D) {
  int col_blocks = cdiv(D, CW);
  int want = cdiv(148 * 4, col_blocks);
  int max_split = cdiv(B, RY * 4);
  int s = want < 1 ? 1 : want;
  if (s > max_split) s = max_split;
  return s < 1 ? 1 : s;
}

template <bool SQRT>
__global__ void __launch_bounds__(CW* RY) k_colstats(const float* __restrict__ U, int B, int D, int rows_per_cta,
                                                      double* __restrict__ acc, StatFinal fin) {
  pdl_prologue();
  int col = blockIdx.x * CW + threadIdx.x;
  int r0 = blockIdx.y * rows_per_cta;
  int r1 = min(B, r0 + rows_per_cta);
  double s = 0.0, ss = 0.0;
  if (col < D) {
    int r = r0 + threadIdx.y;
    for (; r + 3 * RY < r1; r += 4 * RY) {
      float v0 = U[(size_t)r * D + col], v1 = U[(size_t)(r + RY) * D + col];
      float v2 = U[(size_t)(r + 2 * RY) * D + col], v3 = U[(size_t)(r + 3 * RY) * D + col];
      if (SQRT) { v0 = sqrt_count(v0); v1 = sqrt_count(v1); v2 = sqrt_count(v2); v3 = sqrt_count(v3); }
      s += (double)v0 + (double)v1 + (double)v2 + (double)v3;
      ss += (double)v0 * v0 + (double)v1 * v1 + (double)v2 * v2 + (double)v3 * v3;
    }
    for (; r < r1; r += RY) {
      float v = U[(size_t)r * D + col];
      if (SQRT) v = sqrt_count(v);
      s += v;
      ss += (double)v * v;
    }
  }
  __shared__ double sh[2][RY][CW];
  sh[0][threadIdx.y][threadIdx.x] = s;
  sh[1][threadIdx.y][threadIdx.x] = ss;
  __syncthreads();
  if (threadIdx.y < 2 && col < D) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < RY; ++i) t += sh[threadIdx.y][i][threadIdx.x];
    atomicAdd(&acc[(size_t)threadIdx.y * D + col], t);
  }
  stat_finalize_if_last(fin, acc, D, blockIdx.x * CW, CW);
}

// Wide version (D % 4 == 0): 128-bit loads, 4 columns per thread, 2 rows in flight per thread.
// fp32 partial sums over <= 32 rows are folded into fp64 accumulators.
template <bool SQRT>
__global__ void __launch_bounds__(CW* RY) k_colstats4(const float* __restrict__ U, int B, int D, int rows_per_cta,
                                                       double* __restrict__ acc, StatFinal fin) {
  pdl_prologue();
  const int col = (blockIdx.x * CW + threadIdx.x) * 4;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(B, r0 + rows_per_cta);
  double s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
  if (col < D) {
    int r = r0 + threadIdx.y;
    while (r < r1) {
      float ps[4] = {0.f, 0.f, 0.f, 0.f}, pss[4] = {0.f, 0.f, 0.f, 0.f};
      int chunk_end = min(r1, r + 32 * RY);
      for (; r + 3 * RY < chunk_end; r += 4 * RY) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(U + (size_t)r * D + col));
        const float4 b = __ldcs(reinterpret_cast<const float4*>(U + (size_t)(r + RY) * D + col));
        const float4 c = __ldcs(reinterpret_cast<const float4*>(U + (size_t)(r + 2 * RY) * D + col));
        const float4 d = __ldcs(reinterpret_cast<const float4*>(U + (size_t)(r + 3 * RY) * D + col));
        float v[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (SQRT) v[j] = sqrt_count(v[j]);
          ps[j & 3] += v[j];
          pss[j & 3] = __fmaf_rn(v[j], v[j], pss[j & 3]);
        }
      }
      for (; r < chunk_end; r += RY) {
        const float4 a = *reinterpret_cast<const float4*>(U + (size_t)r * D + col);
        float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (SQRT) v[j] = sqrt_count(v[j]);
          ps[j] += v[j];
          pss[j] = __fmaf_rn(v[j], v[j], pss[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[j] += (double)ps[j];
        ss[j] += (double)pss[j];
      }
    }
  }
  __shared__ double sh[2][RY][CW * 4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sh[0][threadIdx.y][threadIdx.x * 4 + j] = s[j];
    sh[1][threadIdx.y][threadIdx.x * 4 + j] = ss[j];
  }
  __syncthreads();
  const int tid = threadIdx.y * CW + threadIdx.x;
  const int which = tid / (CW * 4), c = tid % (CW * 4);
  const int gcol = blockIdx.x * CW * 4 + c;
  if (gcol < D) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < RY; ++i) t += sh[which][i][c];
    atomicAdd(&acc[(size_t)which * D + gcol], t);
  }
  stat_finalize_if_last(fin, acc, D, blockIdx.x * CW * 4, CW * 4);
}

// GATE: dY holds the UNGATED gradient w.r.t. the ReLU output; the gate [bn(U) > 0] is recomputed here
// (same fused-multiply-add sequence as the forward), saving a separate gating pass.
template <bool GATE>
__global__ void __launch_bounds__(CW* RY) k_bn_bwd_reduce(const float* __restrict__ dY, const float* __restrict__ U, int B,
                                                           int D, int rows_per_cta, BnParams bn,
                                                           double* __restrict__ acc) {
  pdl_prologue();
  int col = blockIdx.x * CW + threadIdx.x;
  int r0 = blockIdx.y * rows_per_cta;
  int r1 = min(B, r0 + rows_per_cta);
  double s1 = 0.0, s2 = 0.0;
  if (col < D) {
    float mean = bn.mean[col], rstd = bn.rstd[col], gam = bn.gamma[col], bet = bn.beta[col];
    for (int r = r0 + threadIdx.y; r < r1; r += RY) {
      float g = dY[(size_t)r * D + col];
      float xh = bn_xhat(U[(size_t)r * D + col], mean, rstd);
      if (GATE && !(bn_y(xh, gam, bet) > 0.f)) g = 0.f;
      s1 += g;
      s2 += (double)g * xh;
    }
  }
  __shared__ double sh[2][RY][CW];
  sh[0][threadIdx.y][threadIdx.x] = s1;
  sh[1][threadIdx.y][threadIdx.x] = s2;
  __syncthreads();
  if (threadIdx.y < 2 && col < D) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < RY; ++i) t += sh[threadIdx.y][i][threadIdx.x];
    atomicAdd(&acc[(size_t)threadIdx.y * D + col], t);
  }
}

template <bool GATE>
__global__ void k_bn_bwd_apply(float* __restrict__ dY, const float* __restrict__ U, int B, int D, BnParams bn,
                               const double* __restrict__ acc, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_prologue();
  size_t n = (size_t)B * D;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    int col = (int)(idx % D);
    const float invB = 1.0f / (float)B;
    float s1 = (float)acc[col] * invB, s2 = (float)acc[D + col] * invB;
    float rstd = bn.rstd[col];
    float xh = bn_xhat(U[idx], bn.mean[col], rstd);
    float g = dY[idx];
    if (GATE && !(bn_y(xh, bn.gamma[col], bn.beta[col]) > 0.f)) g = 0.f;
    dY[idx] = bn.gamma[col] * rstd * (g - s1 - xh * s2);
    if (idx < (size_t)D) {
      dbeta[col] = (float)acc[col];
      dgamma[col] = (float)acc[D + col];
    }
  }
}

// BatchNorm backward apply fused with the preparation of the NEXT Linear's backward operands: dU is
// written once, tf32-rounded (the MMA operand of both the dA and the dW GEMM of the block below),
// and its unrounded column sums are accumulated into that block's bias gradient.
template <bool GATE>
__global__ void __launch_bounds__(256) k_bn_bwd_apply_round(const float* __restrict__ dY, const float* __restrict__ U, int B,
                                                            int D, int rows_per_cta, BnParams bn,
                                                            const double* __restrict__ acc, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, float* __restrict__ dU_r,
                                                            float* __restrict__ db_next) {
  pdl_prologue();
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(B, r0 + rows_per_cta);
  float s = 0.f;
  if (col < D) {
    const float invB = 1.0f / (float)B;
    const float s1 = (float)acc[col] * invB, s2 = (float)acc[D + col] * invB;
    const float rstd = bn.rstd[col], mean = bn.mean[col], gamma = bn.gamma[col], beta = bn.beta[col];
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      const size_t idx = (size_t)r * D + col;
      const float xh = bn_xhat(U[idx], mean, rstd);
      float g = dY[idx];
      if (GATE && !(bn_y(xh, gamma, beta) > 0.f)) g = 0.f;
      const float du = gamma * rstd * (g - s1 - xh * s2);
      uint32_t rr;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(rr) : "f"(du));
      dU_r[idx] = __uint_as_float(rr);
      s += du;
    }
    if (blockIdx.y == 0 && threadIdx.y == 0) {
      dbeta[col] = (float)acc[col];
      dgamma[col] = (float)acc[D + col];
    }
  }
  __shared__ float sh[8][32];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && col < D) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
    atomicAdd(&db_next[col], t);
  }
}

__global__ void k_bn_grad_finalize(const double* __restrict__ acc, int D, float* __restrict__ dgamma,
                                   float* __restrict__ dbeta) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D) return;
  dbeta[i] = (float)acc[i];
  dgamma[i] = (float)acc[D + i];
}

__global__ void k_bump(int64_t* c, int n) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) c[i] += 1;
}

// Last launch of a step: BatchNorm batch counters, the decoder's log-likelihood sum folded into the
// loss (ll may be null: the CUDA-core NB kernel adds it itself), status word + NaN flag.
__global__ void __launch_bounds__(256) k_finalize_step(drv_step_out* out, const uint32_t* status, int64_t* c, int n,
                                                       const float* __restrict__ ll, int B, float scale,
                                                       float* __restrict__ loss_slot, uint32_t path) {
  pdl_prologue();
  const int i = threadIdx.x;
  if (c && i < n) c[i] += 1;
  double t = 0.0;
  if (ll)
    for (int b = i; b < B; b += blockDim.x) t += (double)ll[b];
  t = warp_sum(t);
  __shared__ double sh[8];
  if (i % 32 == 0) sh[i / 32] = t;
  __syncthreads();
  if (i == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)blockDim.x / 32; ++w) s += sh[w];
    const double loss = out->loss - (double)scale * s;
    out->loss = loss;
    uint32_t st = *status;
    if (isnan(loss)) st |= DRV_STATUS_NAN_LOSS;
    if (loss_slot) {
      // data parallel: the loss and the status bits ride in the gradient buffer's tail slots, one float per bit, so
      // that the SUM all-reduce leaves "how many ranks raised it" and every rank sees the same loss and status
      loss_slot[0] = (float)loss;
      loss_slot[1] = (st & DRV_STATUS_NAN_LOSS) ? 1.f : 0.f;
      loss_slot[2] = (st & DRV_STATUS_MISSING_CLASS) ? 1.f : 0.f;
      loss_slot[3] = (st & DRV_STATUS_BAD_LABEL) ? 1.f : 0.f;
    }
    out->status = st;
    out->path = path;
  }
}

}  // namespace

void colstats(const float* U, int B, int D, bool sqrt_in, double* acc, unsigned int* ticket, float eps, float momentum,
              float* mean, float* rstd, float* running_mean, float* running_var, cudaStream_t st) {
  StatFinal fin{ticket, B, eps, momentum, mean, rstd, running_mean, running_var};
  if (D % 4 == 0 && D >= 512 && ((uintptr_t)U % 16) == 0) {
    int col_blocks = cdiv(D, CW * 4);
    int S = cdiv(148 * 4, col_blocks);
    int max_split = cdiv(B, RY * 32);
    if (S > max_split) S = max_split;
    if (S < 1) S = 1;
    int rows = cdiv(cdiv(B, S), RY) * RY;
    S = cdiv(B, rows);
    dim3 grid(col_blocks, S), block(CW, RY);
    if (sqrt_in) launch_k(k_colstats4<true>, grid, block, 0, st, U, B, D, rows, acc, fin);
    else launch_k(k_colstats4<false>, grid, block, 0, st, U, B, D, rows, acc, fin);
    note_launch();
    return;
  }
  int S = row_splits(B, D);
  int rows = cdiv(B, S);
  rows = cdiv(rows, RY) * RY;
  S = cdiv(B, rows);
  dim3 grid(cdiv(D, CW), S), block(CW, RY);
  if (sqrt_in) launch_k(k_colstats<true>, grid, block, 0, st, U, B, D, rows, acc, fin);
  else launch_k(k_colstats<false>, grid, block, 0, st, U, B, D, rows, acc, fin);
  note_launch();
}

void bn_bwd_reduce(const float* dY, const float* U, int B, int D, BnParams bn, double* acc, bool gate, cudaStream_t st) {
  int S = row_splits(B, D);
  int rows = cdiv(B, S);
  rows = cdiv(rows, RY) * RY;
  S = cdiv(B, rows);
  dim3 grid(cdiv(D, CW), S), block(CW, RY);
  if (gate) launch_k(k_bn_bwd_reduce<true>, grid, block, 0, st, dY, U, B, D, rows, bn, acc);
  else launch_k(k_bn_bwd_reduce<false>, grid, block, 0, st, dY, U, B, D, rows, bn, acc);
  note_launch();
}

void bn_bwd_apply(float* dY_dU, const float* U, int B, int D, BnParams bn, const double* acc, float* dgamma,
                  float* dbeta, bool gate, cudaStream_t st) {
  size_t n = (size_t)B * D;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (gate) launch_k(k_bn_bwd_apply<true>, blocks, 256, 0, st, dY_dU, U, B, D, bn, acc, dgamma, dbeta);
  else launch_k(k_bn_bwd_apply<false>, blocks, 256, 0, st, dY_dU, U, B, D, bn, acc, dgamma, dbeta);
  note_launch();
}

void bn_bwd_apply_round(const float* dY, const float* U, int B, int D, BnParams bn, const double* acc, float* dgamma,
                        float* dbeta, bool gate, float* dU_r, float* db_next, cudaStream_t st) {
  const int rows = 64;
  dim3 grid(cdiv(D, 32), cdiv(B, rows)), block(32, 8);
  if (gate) launch_k(k_bn_bwd_apply_round<true>, grid, block, 0, st, dY, U, B, D, rows, bn, acc, dgamma, dbeta, dU_r, db_next);
  else launch_k(k_bn_bwd_apply_round<false>, grid, block, 0, st, dY, U, B, D, rows, bn, acc, dgamma, dbeta, dU_r, db_next);
  note_launch();
}

void bn_grad_finalize(const double* acc, int D, float* dgamma, float* dbeta, cudaStream_t st) {
  launch_k(k_bn_grad_finalize, cdiv(D, 256), 256, 0, st, acc, D, dgamma, dbeta);
  note_launch();
}

void bump_counters(int64_t* counters, int n, cudaStream_t st) {
  launch_k(k_bump, cdiv(n, 64), 64, 0, st, counters, n);
  note_launch();
}

void finalize_step(drv_step_out* out, const uint32_t* status, int64_t* counters, int n, const float* ll, int B,
                   float scale, cudaStream_t st, float* loss_slot, uint32_t path) {
  launch_k(k_finalize_step, 1, 256, 0, st, out, status, counters, n, ll, B, scale, loss_slot, path);
  note_launch();
}

}  // namespace drv
