// Molecular-cross-validation split of a UMI count tile on the device: the
// reference's `MCVDataLoader.split_molecules` (src/drvish/train/mcv_train.py:20-40) draws, per
// count n,
//     x_data ~ Binomial(n, data_split - overlap_factor)
//     y_data ~ Binomial(n - x_data, (1 - data_split) / (1 - data_split + overlap_factor))
//     overlap = n - x_data - y_data
// and returns (x_data + overlap, y_data + overlap, log data_split, log data_split_complement) with
// torch.distributions on the host, every iteration.  Here one launch does it next to the data:
// HBM-bound (4 bytes read, 8 written per count), counter-based Philox draws keyed by
// (seed, call index, element, draw), so a split is reproducible and independent of the launch shape.
//
// Binomial sampler: exact inversion of the CDF by the pmf recurrence on min(p, 1-p), in chunks of at
// most 64 trials (Binomial(n, p) = sum of independent chunk draws; q^64 >= 5e-20 never underflows in
// fp32).  Expected work n*min(p,1-p) + n/64 steps per count; single-cell counts are mostly 0-10.
#include "kernels.cuh"
#include "philox.cuh"

namespace drv {
namespace {

constexpr int SPLIT_CHUNK = 64;

struct BinomP {  // per-row constants of one draw
  float pp;      // min(p, 1-p)
  float ratio;   // pp / (1 - pp)
  float log2q;   // log2(1 - pp)
  int flip;      // p > 1/2: sample the failures
  int degenerate;  // 0: sample, 1: always 0, 2: always n
};

__device__ __forceinline__ BinomP make_binom(float p) {
  BinomP b;
  b.degenerate = !(p > 0.f) ? 1 : (p >= 1.f ? 2 : 0);
  b.flip = p > 0.5f;
  b.pp = b.flip ? 1.f - p : p;
  const float q = 1.f - b.pp;
  b.ratio = b.pp / q;
  b.log2q = log2f(q);
  return b;
}

// number of successes among m <= SPLIT_CHUNK trials, by inversion with uniform u
__device__ __forceinline__ int binom_chunk(int m, const BinomP& b, float u) {
  float pm = exp2f((float)m * b.log2q);
  float cdf = pm;
  int k = 0;
  while (u > cdf && k < m) {
    ++k;
    pm *= b.ratio * (float)(m - k + 1) / (float)k;
    cdf += pm;
  }
  return k;
}

__device__ __forceinline__ float binom_draw(float nf, const BinomP& b, uint64_t seed, uint64_t elem, uint32_t call,
                                            uint32_t draw) {
  if (!(nf >= 1.f) || b.degenerate == 1) return 0.f;
  if (b.degenerate == 2) return nf;
  int n = (int)nf;
  int succ = 0;
  uint32_t group = 0;
  while (n > 0) {
    uint32_t r[4];
    philox4(seed, elem, r, 0x73706C74u ^ call, (draw << 28) | group);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (n > 0) {
        const int m = n < SPLIT_CHUNK ? n : SPLIT_CHUNK;
        succ += binom_chunk(m, b, u01(r[i]));
        n -= m;
      }
    }
    ++group;
  }
  const int total = (int)nf;
  return (float)(b.flip ? total - succ : succ);
}

__device__ __forceinline__ void split_one(float n, const BinomP& b1, const BinomP& b2, uint64_t seed, uint64_t elem,
                                          uint32_t call, float& x0, float& x1) {
  const float xd = binom_draw(n, b1, seed, elem, call, 0u);
  const float yd = binom_draw(n - xd, b2, seed, elem, call, 1u);
  // x_data + overlap = n - y_data, y_data + overlap = n - x_data (mcv_train.py:33-36)
  x0 = n - yd;
  x1 = n - xd;
}

template <bool VEC4>
__global__ void __launch_bounds__(256) k_split_molecules(const float* __restrict__ umis, const float* __restrict__ data_split,
                                                         const float* __restrict__ data_split_complement,
                                                         const float* __restrict__ overlap_factor, int B, int G,
                                                         uint64_t seed, uint32_t call, float* __restrict__ x0,
                                                         float* __restrict__ x1, float* __restrict__ log_split,
                                                         float* __restrict__ log_split_complement) {
  pdl_prologue();
  const int per_row = VEC4 ? G / 4 : G;
  const int64_t total = (int64_t)B * per_row;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(t / per_row), c = (int)(t - (int64_t)row * per_row);
    const float s = data_split[row], ov = overlap_factor[row];
    const BinomP b1 = make_binom(s - ov);
    const BinomP b2 = make_binom((1.f - s) / (1.f - s + ov));
    if (c == 0) {
      log_split[row] = logf(s);
      log_split_complement[row] = logf(data_split_complement[row]);
    }
    if (VEC4) {
      const float4 n = __ldcs(reinterpret_cast<const float4*>(umis) + t);
      const uint64_t e = (uint64_t)t * 4;
      float4 a, b;
      split_one(n.x, b1, b2, seed, e + 0, call, a.x, b.x);
      split_one(n.y, b1, b2, seed, e + 1, call, a.y, b.y);
      split_one(n.z, b1, b2, seed, e + 2, call, a.z, b.z);
      split_one(n.w, b1, b2, seed, e + 3, call, a.w, b.w);
      reinterpret_cast<float4*>(x0)[t] = a;
      reinterpret_cast<float4*>(x1)[t] = b;
    } else {
      float a, b;
      split_one(umis[t], b1, b2, seed, (uint64_t)t, call, a, b);
      x0[t] = a;
      x1[t] = b;
    }
  }
}

}  // namespace
}  // namespace drv

extern "C" int drv_split_molecules(const float* umis, const float* data_split, const float* data_split_complement,
                                   const float* overlap_factor, int n_rows, int n_genes, uint64_t seed, uint64_t call_index,
                                   float* x0, float* x1, float* log_split, float* log_split_complement, void* stream) {
  using namespace drv;
  g_launches = 0;
  g_err = cudaSuccess;
  if (!umis || !data_split || !data_split_complement || !overlap_factor || !x0 || !x1 || !log_split ||
      !log_split_complement || n_rows < 0 || n_genes < 1)
    return DRV_EINVAL;
  if (n_rows == 0) return 0;
  const bool vec = (n_genes % 4 == 0) && !((reinterpret_cast<uintptr_t>(umis) | reinterpret_cast<uintptr_t>(x0) |
                                            reinterpret_cast<uintptr_t>(x1)) & 15);
  const int64_t threads = (int64_t)n_rows * (vec ? n_genes / 4 : n_genes);
  int64_t blocks = (threads + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  const uint32_t call = (uint32_t)(call_index ^ (call_index >> 32));
  if (vec)
    launch_k(k_split_molecules<true>, (unsigned)blocks, 256, 0, (cudaStream_t)stream, umis, data_split, data_split_complement,
             overlap_factor, n_rows, n_genes, seed, call, x0, x1, log_split, log_split_complement);
  else
    launch_k(k_split_molecules<false>, (unsigned)blocks, 256, 0, (cudaStream_t)stream, umis, data_split, data_split_complement,
             overlap_factor, n_rows, n_genes, seed, call, x0, x1, log_split, log_split_complement);
  note_launch();
  cudaError_t e = g_err;
  g_err = cudaSuccess;
  return (int)e;
}
