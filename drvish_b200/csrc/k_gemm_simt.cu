// General-shape fp32 GEMMs (CUDA-core FFMA, shared-memory tiled).  They serve every shape the
// tcgen05/TMA kernels do not take (ragged gene counts, tiny trunk layers) and are the on-device
// A/B check for the tensor-core path.  One kernel template covers the three contractions of a
// BatchNorm1d -> ReLU -> Linear block (reference modules.py:15-23) and its backward:
//
//   forward   V[b,j]  = sum_i act(U)[b,i] W[j,i] + bias[j]           A: k-contiguous + prologue, B: k-contiguous
//   dW / db   dW[j,i] = sum_b dV[b,j] act(U)[b,i]; db[j] = sum_b dV   A: m-contiguous, B: n-contiguous + prologue (+ ones column)
//   dY        dY[b,i] = (sum_j dV[b,j] W[j,i]) * [y(U)[b,i] > 0]      A: k-contiguous, B: n-contiguous
//   d gamma/beta of the first encoder block: same as dY but reduced over b on the fly.
#include "kernels.cuh"

namespace drv {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4, NT = 256;

struct Prologue {
  int mode;  // 0 none, 1 BN+ReLU, 2 sqrt then BN+ReLU
  BnParams bn;
};

__device__ __forceinline__ float apply_pro(const Prologue& p, float u, int col) {
  if (p.mode == 0) return u;
  if (p.mode == 2) u = sqrt_count(u);
  return bn_relu(u, p.bn.mean[col], p.bn.rstd[col], p.bn.gamma[col], p.bn.beta[col]);
}

enum { EPI_BIAS = 0, EPI_DW = 1, EPI_GATE = 2, EPI_BNGRAD = 3, EPI_PLAIN = 4 };

struct Epilogue {
  const float* bias;  // EPI_BIAS
  float* db;          // EPI_DW: receives column N-1
  const float* U;     // EPI_GATE / EPI_BNGRAD: block input [M][N] (ld = ldu)
  int ldu;
  Prologue gate;      // statistics of that block
  double* acc;        // EPI_BNGRAD: [2][N] (sum dY, sum dY*xhat)
};

// AK: A is k-contiguous (A[m*lda+k]) else m-contiguous (A[k*lda+m]).
// BKC: B is k-contiguous (B[n*ldb+k]) else n-contiguous (B[k*ldb+n]).
// ONES: logical column N-1 of B is all ones (bias-gradient trick, only with !BKC).
template <bool AK, bool BKC, bool ONES, int EPI>
__global__ void __launch_bounds__(NT) k_gemm(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
                                            float* __restrict__ C, int ldc, int M, int N, int K, Prologue proA,
                                            Prologue proB, Epilogue ep) {
  pdl_prologue();
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = tid % 16, ty = tid / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int idx = tid + NT * r;
      int m, k;
      if (AK) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
      int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < K) {
        v = AK ? A[(size_t)gm * lda + gk] : A[(size_t)gk * lda + gm];
        if (AK) v = apply_pro(proA, v, gk);
      }
      ra[r] = v;
      int n;
      if (BKC) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
      int gn = n0 + n;
      gk = k0 + k;
      v = 0.f;
      if (gn < N && gk < K) {
        if (ONES && gn == N - 1) {
          v = 1.f;
        } else {
          v = BKC ? Bm[(size_t)gn * ldb + gk] : Bm[(size_t)gk * ldb + gn];
          if (!BKC) v = apply_pro(proB, v, gn);
        }
      }
      rb[r] = v;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int idx = tid + NT * r;
      int m, k, n;
      if (AK) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
      As[k][m] = ra[r];
      if (BKC) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
      Bs[k][n] = rb[r];
    }
  };

  // split-K over gridDim.z (EPI_DW only: partial tiles are combined with atomics into a zeroed C)
  const int k_per = ((K + gridDim.z - 1) / gridDim.z + BK - 1) / BK * BK;
  const int k_begin = blockIdx.z * k_per;
  const int k_end = min(K, k_begin + k_per);
  if (k_begin >= k_end) return;
  const int K_full = K;
  K = k_end;
  (void)K_full;
  load_tiles(k_begin);
  for (int k0 = k_begin; k0 < K; k0 += BK) {
    __syncthreads();
    store_tiles();
    __syncthreads();
    if (k0 + BK < K) load_tiles(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(av[i], bv[j], acc[i][j]);
    }
  }

  if (EPI == EPI_BNGRAD) {
    // gate with the recomputed ReLU mask, reduce over the rows of this tile, one double atomic
    // per column per CTA.
    __shared__ float red[2][16][BN];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      float s1 = 0.f, s2 = 0.f;
      if (gn < N) {
        float mean = ep.gate.bn.mean[gn], rstd = ep.gate.bn.rstd[gn], g = ep.gate.bn.gamma[gn], be = ep.gate.bn.beta[gn];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int gm = m0 + ty * 4 + i;
          if (gm < M) {
            float u = ep.U[(size_t)gm * ep.ldu + gn];
            if (ep.gate.mode == 2) u = sqrt_count(u);
            float xh = bn_xhat(u, mean, rstd);
            if (bn_y(xh, g, be) > 0.f) {
              s1 += acc[i][j];
              s2 = __fmaf_rn(acc[i][j], xh, s2);
            }
          }
        }
      }
      red[0][ty][tx * 4 + j] = s1;
      red[1][ty][tx * 4 + j] = s2;
    }
    __syncthreads();
    if (tid < 2 * BN) {
      int which = tid / BN, n = tid % BN;
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < 16; ++t) s += red[which][t][n];
      int gn = n0 + n;
      if (gn < N) atomicAdd(&ep.acc[(size_t)which * N + gn], (double)s);
    }
    return;
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (EPI == EPI_BIAS) {
        C[(size_t)gm * ldc + gn] = v + ep.bias[gn];
      } else if (EPI == EPI_DW) {
        float* dst = (ONES && gn == N - 1) ? &ep.db[gm] : &C[(size_t)gm * ldc + gn];
        if (gridDim.z > 1) atomicAdd(dst, v);
        else *dst = v;
      } else if (EPI == EPI_PLAIN) {
        C[(size_t)gm * ldc + gn] = v;
      } else if (EPI == EPI_GATE) {
        float u = ep.U[(size_t)gm * ep.ldu + gn];
        if (ep.gate.mode == 2) u = sqrt_count(u);
        float y = bn_y(bn_xhat(u, ep.gate.bn.mean[gn], ep.gate.bn.rstd[gn]), ep.gate.bn.gamma[gn], ep.gate.bn.beta[gn]);
        C[(size_t)gm * ldc + gn] = y > 0.f ? v : 0.f;
      }
    }
  }
}

inline dim3 grid_for(int M, int N) { return dim3(cdiv(N, BN), cdiv(M, BM)); }

}  // namespace

void simt_linear_forward(const float* U, int B, int din, int pro_mode, BnParams bn, const float* W, const float* bias,
                         int dout, float* V, cudaStream_t st) {
  Prologue pa{pro_mode, bn}, pb{0, {}};
  Epilogue ep{};
  ep.bias = bias;
  launch_k(k_gemm<true, true, false, EPI_BIAS>, grid_for(B, dout), NT, 0, st, U, din, W, din, V, dout, B, dout, din, pa, pb, ep);
  note_launch();
}

// V = relu(bn(U)) W^T + b for a NARROW input (din <= 64: the decoder's first Linear, latent -> hidden) with the
// BatchNorm batch statistics of V for the next block in the same launch (replaces the tiled GEMM, whose tiles are
// mostly empty at K = 10, plus a separate statistics kernel: two latency-bound links of the critical chain, 12 + 17 us
// at cfg 3).  Thread = output column (its W row in registers), CTA = NROWS cells whose normalised inputs sit in shared
// memory; the column sums stay in registers, one fp64 atomic per column, moment and CTA, the last CTA of a column
// block finalises (stat_finalize_if_last; grid = (column blocks, row blocks)).
constexpr int NROWS = 32;
template <int DIN_MAX>
__global__ void __launch_bounds__(256) k_narrow_linear_stats(const float* __restrict__ U, int B, int din, BnParams bn,
                                                             const float* __restrict__ W, const float* __restrict__ bias,
                                                             int dout, float* __restrict__ V, double* __restrict__ acc,
                                                             StatFinal fin) {
  pdl_prologue();
  __shared__ float a[NROWS][DIN_MAX];
  const int r0 = blockIdx.y * NROWS;
  for (int idx = threadIdx.x; idx < NROWS * din; idx += blockDim.x) {
    const int r = idx / din, i = idx - r * din;
    const int row = r0 + r;
    a[r][i] = row < B ? bn_relu(U[(size_t)row * din + i], bn.mean[i], bn.rstd[i], bn.gamma[i], bn.beta[i]) : 0.f;
  }
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const bool col_ok = j < dout;
  float w[DIN_MAX];
#pragma unroll
  for (int i = 0; i < DIN_MAX; ++i) w[i] = (col_ok && i < din) ? __ldg(&W[(size_t)j * din + i]) : 0.f;
  const float bj = col_ok ? __ldg(&bias[j]) : 0.f;
  __syncthreads();
  float s1 = 0.f, s2 = 0.f;
  const int nr = min(NROWS, B - r0);
  if (col_ok) {
    for (int r = 0; r < nr; ++r) {
      float v = bj;
#pragma unroll
      for (int i = 0; i < DIN_MAX; ++i)
        if (i < din) v = __fmaf_rn(a[r][i], w[i], v);
      V[(size_t)(r0 + r) * dout + j] = v;
      s1 += v;
      s2 = __fmaf_rn(v, v, s2);
    }
    atomicAdd(&acc[j], (double)s1);
    atomicAdd(&acc[(size_t)dout + j], (double)s2);
  }
  stat_finalize_if_last(fin, acc, dout, blockIdx.x * blockDim.x, blockDim.x);
}

bool simt_narrow_linear_forward_stats(const float* U, int B, int din, BnParams bn, const float* W, const float* bias,
                                      int dout, float* V, const ColStatsOut& ns, cudaStream_t st) {
  if (din > 64 || din < 1) return false;
  dim3 grid(cdiv(dout, 256), cdiv(B, NROWS));
  if (din <= 16) launch_k(k_narrow_linear_stats<16>, grid, 256, 0, st, U, B, din, bn, W, bias, dout, V, ns.acc, ns.fin);
  else if (din <= 32) launch_k(k_narrow_linear_stats<32>, grid, 256, 0, st, U, B, din, bn, W, bias, dout, V, ns.acc, ns.fin);
  else launch_k(k_narrow_linear_stats<64>, grid, 256, 0, st, U, B, din, bn, W, bias, dout, V, ns.acc, ns.fin);
  note_launch();
  return true;
}

void simt_linear_wgrad(const float* dV, int B, int dout, const float* U, int din, int pro_mode, BnParams bn, float* dW,
                       float* db, cudaStream_t st) {
  Prologue pa{0, {}}, pb{pro_mode, bn};
  Epilogue ep{};
  ep.db = db;
  // M = dout, N = din + 1 (ones column -> db), K = B; split-K when the output is too small to
  // fill the SMs (the reduction runs over the whole minibatch)
  dim3 grid = grid_for(dout, din + 1);
  int ctas = grid.x * grid.y;
  int splits = 1;
  if (ctas < 296) {
    splits = 592 / ctas;
    int max_splits = B / 128;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  if (splits > 1) grid.z = splits;  // split-K accumulation: dW / db were cleared at the start of the step
  launch_k(k_gemm<false, false, true, EPI_DW>, grid, NT, 0, st, dV, dout, U, din, dW, din, dout, din + 1, B, pa, pb, ep);
  note_launch();
}

// dA = dV W for a narrow input (din <= 32, e.g. the decoder's first Linear, latent -> hidden): the
// generic tile kernel would run 64 CTAs with most of each tile empty.  One warp per row: the lanes
// split the dout-long dot products (coalesced reads of the dV row, W^T staged in shared memory with
// the dout index contiguous -> conflict-free), then a butterfly leaves column i's sum in every lane.
template <int DIN_MAX>
__global__ void __launch_bounds__(256) k_dgrad_skinny(const float* __restrict__ dV, int B, int dout,
                                                      const float* __restrict__ W, int din, float* __restrict__ dA) {
  pdl_prologue();
  extern __shared__ float sk_sh[];  // W^T: [din][dout]
  for (int i = threadIdx.x; i < dout * din; i += blockDim.x) {
    const int j = i / din, c = i % din;
    sk_sh[c * dout + j] = W[i];
  }
  __syncthreads();
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  for (int r = blockIdx.x * 8 + warp; r < B; r += gridDim.x * 8) {
    float acc[DIN_MAX];
#pragma unroll
    for (int i = 0; i < DIN_MAX; ++i) acc[i] = 0.f;
    const float* row = dV + (size_t)r * dout;
    for (int j = lane; j < dout; j += 32) {
      const float v = row[j];
#pragma unroll
      for (int i = 0; i < DIN_MAX; ++i)
        if (i < din) acc[i] = __fmaf_rn(v, sk_sh[i * dout + j], acc[i]);
    }
    float mine = 0.f;
#pragma unroll
    for (int i = 0; i < DIN_MAX; ++i) {
      if (i < din) {
        const float t = warp_sum(acc[i]);
        if (lane == i) mine = t;
      }
    }
    if (lane < din) dA[(size_t)r * din + lane] = mine;
  }
}

// dA = dV W (ungated): the ReLU gate is applied by the BatchNorm backward kernels
void simt_linear_dgrad_gate(const float* dV, int B, int dout, const float* W, int din, const float* U, int pro_mode,
                            BnParams bn, float* dY, cudaStream_t st) {
  (void)U; (void)pro_mode; (void)bn;
  const size_t sk_bytes = sizeof(float) * (size_t)dout * din;
  if (din <= 32 && sk_bytes <= 48 * 1024) {
    const int grid = cdiv(B, 8) < 148 * 4 ? cdiv(B, 8) : 148 * 4;
    if (din <= 16) launch_k(k_dgrad_skinny<16>, grid, 256, sk_bytes, st, dV, B, dout, W, din, dY);
    else launch_k(k_dgrad_skinny<32>, grid, 256, sk_bytes, st, dV, B, dout, W, din, dY);
    note_launch();
    return;
  }
  Prologue none{0, {}};
  Epilogue ep{};
  launch_k(k_gemm<true, false, false, EPI_PLAIN>, grid_for(B, din), NT, 0, st, dV, dout, W, din, dY, din, B, din, dout, none,
                                                                         none, ep);
  note_launch();
}

void simt_linear_dgrad_bngrad(const float* dV, int B, int dout, const float* W, int din, const float* U, int pro_mode,
                              BnParams bn, double* acc, cudaStream_t st) {
  Prologue none{0, {}};
  Epilogue ep{};
  ep.U = U;
  ep.ldu = din;
  ep.gate = Prologue{pro_mode, bn};
  ep.acc = acc;
  launch_k(k_gemm<true, false, false, EPI_BNGRAD>, grid_for(B, din), NT, 0, st, dV, dout, W, din, nullptr, 0, B, din, dout,
                                                                          none, none, ep);
  note_launch();
}

}  // namespace drv
