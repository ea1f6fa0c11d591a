// Shape padding for the tensor-core path.  The tcgen05 / TMA kernels want 16-byte-aligned rows and halves
// (genes % 8 == 0) and hidden widths that are multiples of 32; real gene panels and layer widths are arbitrary
// (the reference has no shape restrictions: nbvae.py:28-39).  Instead of dropping such shapes onto the CUDA-core
// path (13x slower), drv_elbo_step runs them as the next aligned problem:
//   * pad genes get count 0, a zero column in the encoder's first weight, zero rows in the decoder's last weight and
//     the scale-half bias -1e30, so their softmax probability, NB log-likelihood and every gradient are exactly 0;
//   * pad hidden units get zero weight rows / columns and zero bias: their BatchNorm input is identically 0, their
//     activation 0, their gradients 0.
// k_pad_params builds the padded flat parameter (or BatchNorm-buffer) vector from the caller's, k_unpad copies the
// gradients (or BatchNorm buffers) of the real entries back, k_pad_rows widens a count tile.  One launch each.
#include "kernels.cuh"

namespace drv {
namespace {

__device__ __forceinline__ int find_seg(const PadSegs& s, long long i, bool dst_space) {
  int lo = 0, hi = s.n - 1;
  while (lo < hi) {  // last segment whose offset is <= i
    const int mid = (lo + hi + 1) / 2;
    const long long off = dst_space ? s.seg[mid].dst_off : s.seg[mid].src_off;
    if (off <= i) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256) k_pad_params(const float* __restrict__ src, float* __restrict__ dst, PadSegs s,
                                                    long long n_dst) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_dst; i += (long long)gridDim.x * blockDim.x) {
    const PadSeg& g = s.seg[find_seg(s, i, true)];
    const long long j = i - g.dst_off;
    float v = 0.f;  // alignment gaps between tensors
    if (j < (long long)g.dst_rows * g.dst_cols) {
      const int r = (int)(j / g.dst_cols), c = (int)(j % g.dst_cols);
      const int blk = (g.n_blocks == 2 && r >= g.blk_stride) ? 1 : 0;
      const int rr = r - blk * g.blk_stride;
      v = (rr < g.src_rows && c < g.src_cols) ? src[g.src_off + ((long long)blk * g.src_rows + rr) * g.src_cols + c] : g.pad[blk];
    }
    dst[i] = v;
  }
}

// real entries only: dst_real[src layout] = padded[dst layout]
__global__ void __launch_bounds__(256) k_unpad(const float* __restrict__ padded, float* __restrict__ real, PadSegs s,
                                               long long n_real, int n_tail, long long tail_padded, long long tail_real) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_real; i += (long long)gridDim.x * blockDim.x) {
    const PadSeg& g = s.seg[find_seg(s, i, false)];
    const long long j = i - g.src_off;
    if (j >= (long long)g.n_blocks * g.src_rows * g.src_cols) continue;  // alignment gap
    const int r = (int)(j / g.src_cols), c = (int)(j % g.src_cols);
    const int blk = r / g.src_rows, rr = r % g.src_rows;
    real[i] = padded[g.dst_off + ((long long)blk * g.blk_stride + rr) * g.dst_cols + c];
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) real[tail_real + threadIdx.x] = padded[tail_padded + threadIdx.x];
}

__global__ void __launch_bounds__(256) k_pad_rows(const float* __restrict__ x, int B, int G, int Gp, float* __restrict__ xp) {
  pdl_prologue();
  const long long n = (long long)B * Gp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / Gp), g = (int)(i % Gp);
    xp[i] = g < G ? __ldcs(&x[(long long)b * G + g]) : 0.f;
  }
}

int blocks_for(long long n) {
  long long b = (n + 255) / 256;
  return (int)(b > 148 * 8 ? 148 * 8 : (b < 1 ? 1 : b));
}

}  // namespace

void pad_params(const float* src, float* dst, const PadSegs& s, long long n_dst, cudaStream_t st) {
  launch_k(k_pad_params, blocks_for(n_dst), 256, 0, st, src, dst, s, n_dst);
  note_launch();
}
void unpad(const float* padded, float* real, const PadSegs& s, long long n_real, int n_tail, long long tail_padded,
           long long tail_real, cudaStream_t st) {
  launch_k(k_unpad, blocks_for(n_real), 256, 0, st, padded, real, s, n_real, n_tail, tail_padded, tail_real);
  note_launch();
}
void pad_rows(const float* x, int B, int G, int Gp, float* xp, cudaStream_t st) {
  launch_k(k_pad_rows, blocks_for((long long)B * Gp), 256, 0, st, x, B, G, Gp, xp);
  note_launch();
}

}  // namespace drv
