// Gradients of the encoder's input BatchNorm (d gamma, d beta over G genes) without materialising
// dA = dH W (B x G):  the GEMM is computed TRANSPOSED, D[g][b] = sum_j W[j][g] dH[b][j], so that a TMEM
// lane is a gene and the accumulator columns are cells - the reduction over the batch is then a
// private sum in each epilogue thread (no cross-lane work), with the ReLU gate and xhat recomputed
// from a TMA-staged count tile.  Replaces a 82 MB store, a 164 MB re-read and a reduction kernel.
// (reference: backward of BatchNorm1d(G) -> ReLU -> Linear(G, h), modules.py:19-21; x is data, so
// no dX is needed.)
//
// A = W_hi^T: MN-major (genes contiguous, 128-byte swizzle with 32-byte atoms), B = dH (tf32-rounded)
// K-major, tile = 128 genes x 128 cells, K = h.  Persistent CTAs, 10 warps: TMA producer, MMA issuer,
// 8 epilogue warps; two TMEM accumulator stages and two count-tile buffers overlap the epilogue of a
// tile with the loads / MMAs of the next.
#include "kernels.cuh"
#include "tc.cuh"
#include "tc_common.cuh"

namespace drv {
namespace tc {

namespace {

constexpr int BG = 128;   // genes per tile (TMEM lanes)
constexpr int BC = 128;   // cells per tile (accumulator columns)
constexpr int BK = 32;
constexpr int NST = 3;
constexpr int A_BYTES = BG * BK * 4;    // 16 KB: 4 boxes of 32 genes x 32 k
constexpr int B_BYTES = BC * BK * 4;    // 16 KB
constexpr int STAGE = A_BYTES + B_BYTES;
constexpr int X_BYTES = 4 * BC * 32 * 4;  // count tile: 4 boxes of 128 cells x 32 genes = 64 KB
constexpr int SMEM_BYTES = NST * STAGE + 2 * X_BYTES + 1024 + 256;
// 8 epilogue warps: two per TMEM lane quarter, each taking half of the tile's cell columns.  With one warp
// per scheduler the epilogue's dependent instruction stream was the bound (ncu: issue slots 35 % busy with
// ~1 eligible warp per scheduler, tensor pipe 26 %).
constexpr int N_EPI = 8;
constexpr int NTHREADS = (2 + N_EPI) * 32;

__global__ void __launch_bounds__(NTHREADS, 1)
k_bn0grad(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH,
          const __grid_constant__ CUtensorMap tmX, BnParams bn, int B, int G, int h, double* __restrict__ gacc) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* xbuf = smem + NST * STAGE;
  uint64_t* full = (uint64_t*)(xbuf + 2 * X_BYTES);
  uint64_t* empty = full + NST;
  uint64_t* acc_full = empty + NST;   // [2]
  uint64_t* acc_empty = acc_full + 2; // [2]
  uint64_t* x_full = acc_empty + 2;   // [2]
  uint64_t* x_empty = x_full + 2;     // [2]
  uint32_t* tmem_slot = (uint32_t*)(x_empty + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int n_gt = (G + BG - 1) / BG, n_ct = (B + BC - 1) / BC;
  const int n_tiles = n_gt * n_ct;
  // gene-major tile order: a CTA sweeps the cell tiles of one gene tile, so that its threads keep
  // their gene's partial sums in registers and issue one pair of atomics per gene tile
  const int t0 = (int)((long long)n_tiles * blockIdx.x / gridDim.x);
  const int t1 = (int)((long long)n_tiles * (blockIdx.x + 1) / gridDim.x);
  const int nkb = (h + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmH);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], N_EPI);
      mbar_init(&x_full[s], 1);
      mbar_init(&x_empty[s], N_EPI);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int kit = 0;
      for (int t = t0, it = 0; t < t1; ++t, ++it) {
        const int g0 = (t / n_ct) * BG, b0 = (t % n_ct) * BC;
        const int xs = it & 1, xph = (it >> 1) & 1;
        mbar_wait(&x_empty[xs], xph ^ 1);
        mbar_expect_tx(&x_full[xs], X_BYTES);
#pragma unroll
        for (int j = 0; j < 4; ++j) tma_load_2d(xbuf + xs * X_BYTES + j * (BC * 128), &tmX, &x_full[xs], g0 + 32 * j, b0);
        for (int kb = 0; kb < nkb; ++kb, ++kit) {
          const int s = kit % NST, ph = (kit / NST) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = smem + s * STAGE;
          mbar_expect_tx(&full[s], STAGE);
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_2d(sa + j * 4096, &tmW, &full[s], g0 + 32 * j, kb * BK);  // W[k][g]
          tma_load_2d(sa + A_BYTES, &tmH, &full[s], kb * BK, b0);                                          // dH[b][k]
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(BG, BC, 1, 0);
      int kit = 0;
      for (int t = t0, it = 0; t < t1; ++t, ++it) {
        const int as = it & 1, aph = (it >> 1) & 1;
        mbar_wait(&acc_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BC);
        for (int kb = 0; kb < nkb; ++kb, ++kit) {
          const int s = kit % NST, ph = (kit / NST) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE), sb = sa + A_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk)
            umma_tf32(d_tmem, smem_desc(sa + kk * 1024, 4096, 512, 1), smem_desc(sb + kk * 32, 16, 1024), idesc,
                      (kb | kk) != 0);
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[as]);
      }
    }
  } else {
    // ===== epilogue: thread = gene, sums over the cells of the tile =====
    const int q = warp % 4;
    const int ch = (warp - 2) / 4;  // which half of the tile's cells
    const int r = q * 32 + lane;  // gene within the tile
    float s1 = 0.f, s2 = 0.f;
    int cur_gt = -1, g = 0;
    float mean = 0.f, rstd = 0.f, gam = 0.f, bet = 0.f;
    auto flush = [&]() {
      if (cur_gt >= 0 && g < G) {
        atomicAdd(&gacc[g], (double)s1);
        atomicAdd(&gacc[(size_t)G + g], (double)s2);
      }
      s1 = s2 = 0.f;
    };
    for (int t = t0, it = 0; t < t1; ++t, ++it) {
      const int gt = t / n_ct, b0 = (t % n_ct) * BC;
      if (gt != cur_gt) {
        flush();
        cur_gt = gt;
        g = gt * BG + r;
        if (g < G) { mean = bn.mean[g]; rstd = bn.rstd[g]; gam = bn.gamma[g]; bet = bn.beta[g]; }
      }
      const int as = it & 1, aph = (it >> 1) & 1;
      mbar_wait(&acc_full[as], aph);
      mbar_wait(&x_full[as], aph);
      tc_fence_after();
      const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BC);
      // count tile: box q (this warp's 32 genes), row = cell, this thread's gene = column `lane`
      const uint8_t* xq = xbuf + as * X_BYTES + q * (BC * 128);
      const int chunk = lane >> 2, within = (lane & 3) * 4;
#pragma unroll 1
      for (int c = ch * (BC / 2); c < (ch + 1) * (BC / 2); c += 16) {
        float v[16];
        tmem_ld16(tb + (uint32_t)c, v);
        float xv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int cell = c + j;
          xv[j] = *reinterpret_cast<const float*>(xq + cell * 128 + ((chunk ^ (cell & 7)) << 4) + within);
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float xh = bn_xhat(sqrt_count(xv[j]), mean, rstd);
          const bool on = (b0 + c + j < B) && (bn_y(xh, gam, bet) > 0.f);
          const float d = on ? v[j] : 0.f;
          s1 += d;
          s2 = __fmaf_rn(d, xh, s2);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&acc_empty[as]);
        mbar_arrive(&x_empty[as]);
      }
    }
    flush();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

int sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace

// gacc[0][g] += sum_b dA[b][g] gate[b][g],  gacc[1][g] += sum_b dA gate xhat,  dA = dH W_hi
int bn0grad_fused(const float* dH_r, const float* W_hi, const float* x, int B, int G, int h, BnParams bn, double* gacc,
                  cudaStream_t st) {
  CUtensorMap tW, tH, tX;
  if (!tmap_2d(&tW, W_hi, h, G, G, BK, 32, 2)) return -3;   // rows = k (h), cols = genes: MN-major boxes 32 x 32
  if (!tmap_2d(&tH, dH_r, B, h, h, BC, BK)) return -3;       // [cells][h] K-major
  if (!tmap_2d(&tX, x, B, G, G, BC, 32)) return -3;          // [cells][genes]
  const int n_tiles = ((G + BG - 1) / BG) * ((B + BC - 1) / BC);
  const int grid = n_tiles < sms() ? n_tiles : sms();
  cudaFuncSetAttribute(k_bn0grad, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  launch_k(k_bn0grad, grid, NTHREADS, SMEM_BYTES, st, tW, tH, tX, bn, B, G, h, gacc);
  note_launch();
  return 0;
}

}  // namespace tc
}  // namespace drv
