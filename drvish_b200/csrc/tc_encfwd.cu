// Encoder input layer forward with the prologue fused into the tensor-core GEMM:
//   V = relu(bn(sqrt(x))) W^T + b        (reference Encoder.forward modules.py:56, make_fc :19-21)
// The count tile is TMA-loaded ONCE, transformed in shared memory by 4 "transform" warps (sqrt,
// BatchNorm, ReLU, split into tf32 hi + lo parts at the very same swizzled offsets, so no layout
// knowledge is needed), and consumed by three MMAs per k-step (3xTF32: A_lo W_hi + A_hi W_lo +
// A_hi W_hi).  The hi tile is also written back to HBM by a TMA store: it is
// the B operand of the weight-gradient GEMM in the backward.  Compared with materialising A_hi /
// A_lo first (prep kernel + 3 operand passes) this removes 2/3 of the HBM traffic of the layer.
//
// Warp roles: 0 TMA producer, 1 TMEM alloc + MMA issuer, 2-9 transform (tf32 variant: 2-5), 10-13 epilogue,
// 14 weight-tile TMA producer.
// Pipeline barriers per stage: full (TMA landed) -> ready (tile transformed, async-proxy fenced)
// -> empty (MMAs retired AND the hi-tile store has finished reading the stage).
// Work = (cell tile, k-part) chunks; every k-part is stored to its own slab and the slabs are
// summed in fixed order (deterministic; accumulation chains <= 32 k-blocks).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "kernels.cuh"
#include "tc.cuh"
#include "tc_common.cuh"

namespace drv {
namespace tc {

namespace {

constexpr int BM = 128, BK = 32, BN = 256;
// Two independent rings: the count tiles need a transform pass (3 stages of hi + lo = 32 KB), the
// weight tiles go straight from TMA to the MMA (2 stages of W_hi + W_lo = 64 KB) and have their own
// producer thread, so a full weight ring never delays the count loads (and vice versa).
constexpr int XST = 3, WST = 2;
constexpr int X_BYTES = BM * BK * 4;        // 16 KB count tile -> becomes the hi tile in place
constexpr int W_BYTES = BN * BK * 4;        // 32 KB
constexpr int XSTAGE = 2 * X_BYTES;         // x/hi, lo
constexpr int WSTAGE = 2 * W_BYTES;         // W_hi, W_lo
constexpr int SMEM_BYTES = XST * XSTAGE + WST * WSTAGE + 1024 + 256;
// H16: TG thread groups of 128 threads share a tile row (64 / TG genes each): the transform pass, not the
// MMA or the loads, bounded the kernel with 4 warps (86 us), 8 warps gave 66 us
constexpr int TG = 4;
constexpr int TW = 4 * TG;               // transform warps (H16); the tf32 variant uses the first 4
constexpr int NTHREADS = (TW + 7) * 32;  // warps: 0 x-TMA, 1 MMA, 2..2+TW-1 transform, then 4 epilogue, 1 W-TMA
constexpr int TMEM_COLS = 512;

__device__ __forceinline__ float rna_tf32_(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// H16: the same 3-term split on fp16 operands (kind::f16).  hi = fp16(a), lo = fp16((a - hi) * 2^11): the
// pair carries the 22 mantissa bits of the tf32 split, the 2^11 keeps the low part in fp16's normal range
// and is taken out of the correction accumulator by the epilogue.  A k-block is 64 genes: two raw fp32
// count boxes (32 KB) become the fp16 hi tile (over box 0) and lo tile (over box 1), W_hi / W_lo tiles are
// 256 x 64 fp16 - half the operand bytes and ring hand-shakes per unit of K, twice the MMA rate; the hi
// tile written back for the weight gradient is fp16 too (half the HBM bytes both ways).
constexpr float LO_SCALE = 2048.f;

// Column sums of 16 values over the 32 lanes of a warp by a transposing butterfly: afterwards w[0] of lane l holds the
// sum of column (l >> 1) (lanes l and l ^ 1 hold the same column); returns that column index.
__device__ __forceinline__ int warp_colsum16(float (&w)[16], int lane) {
#pragma unroll
  for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
    const bool hi = lane & bit;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = hi ? w[i] : w[i + half];
      const float keep = hi ? w[i + half] : w[i];
      w[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  w[0] += __shfl_xor_sync(0xffffffffu, w[0], 1);
  // lane bits 4..1 selected halves 8, 4, 2, 1 in turn: column = bit4 * 8 + bit3 * 4 + bit2 * 2 + bit1
  return (lane >> 1) & 15;
}

// SQRT: the encoder input layer (sqrt of the counts before the BatchNorm); without it the same kernel is one h x h
// trunk block relu(bn(U)) W^T + b.  stats_acc (unsplit chunks only): the epilogue also accumulates the column sums /
// sums of squares of the OUTPUT (the next block's BatchNorm statistics) and the last CTA finalises them.
template <bool H16, bool SQRT>
__global__ void __launch_bounds__(NTHREADS, 1)
k_encfwd(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWhi,
         const __grid_constant__ CUtensorMap tmWlo, const __grid_constant__ CUtensorMap tmAhi, BnParams bn, int B, int G,
         int h, int n_tile, int splits, float* __restrict__ out, long long part_stride, const float* __restrict__ bias,
         double* __restrict__ stats_acc, StatFinal stats_fin) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* xring = smem;
  uint8_t* wring = smem + XST * XSTAGE;
  uint64_t* full = (uint64_t*)(wring + WST * WSTAGE);  // [XST] count tile landed
  uint64_t* ready = full + XST;                        // [XST] tile transformed
  uint64_t* empty = ready + XST;                       // [XST] MMAs retired + hi-tile store drained
  uint64_t* wfull = empty + XST;                       // [WST]
  uint64_t* wempty = wfull + WST;                      // [WST]
  uint64_t* acc_full = wempty + WST;  // [2]
  uint64_t* acc_empty = acc_full + 2; // [2]
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);

  constexpr int BKE = H16 ? 64 : BK;  // genes per k-block
  auto pre = [](float v) { return SQRT ? sqrt_count(v) : v; };
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int kb_total = (G + BKE - 1) / BKE;
  const int m_tiles = (B + BM - 1) / BM, n_tiles = (h + n_tile - 1) / n_tile;
  const long long chunks = (long long)m_tiles * n_tiles * splits;
  const int c0 = (int)(chunks * blockIdx.x / gridDim.x);
  const int c1 = (int)(chunks * (blockIdx.x + 1) / gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmWhi);
    tma_prefetch_desc(&tmWlo);
    tma_prefetch_desc(&tmAhi);
    for (int s = 0; s < XST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&ready[s], H16 ? TW : 4);   // one arrival per transform warp
      mbar_init(&empty[s], 2);   // tcgen05.commit + hi-tile store drained
    }
    for (int s = 0; s < WST; ++s) {
      mbar_init(&wfull[s], 1);
      mbar_init(&wempty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto chunk_range = [&](int ch, int& tile, int& part, int& kb_b, int& kb_e) {
    tile = ch / splits;
    part = ch % splits;
    kb_b = (int)((long long)kb_total * part / splits);
    kb_e = (int)((long long)kb_total * (part + 1) / splits);
  };

  if (warp == 0) {
    // ===== TMA producer: count tiles =====
    if (lane == 0) {
      int i = 0;
      for (int ch = c0; ch < c1; ++ch) {
        int tile, part, kb_b, kb_e;
        chunk_range(ch, tile, part, kb_b, kb_e);
        const int m0 = (tile / n_tiles) * BM;
        for (int kb = kb_b; kb < kb_e; ++kb, ++i) {
          const int s = i % XST, ph = (i / XST) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], H16 ? 2 * X_BYTES : X_BYTES);
          tma_load_2d(xring + s * XSTAGE, &tmX, &full[s], kb * BKE, m0);
          if (H16) tma_load_2d(xring + s * XSTAGE + X_BYTES, &tmX, &full[s], kb * BKE + 32, m0);
        }
      }
    }
  } else if (warp == TW + 6) {
    // ===== TMA producer: weight tiles =====
    if (lane == 0) {
      int i = 0;
      for (int ch = c0; ch < c1; ++ch) {
        int tile, part, kb_b, kb_e;
        chunk_range(ch, tile, part, kb_b, kb_e);
        const int n0 = (tile % n_tiles) * n_tile;
        for (int kb = kb_b; kb < kb_e; ++kb, ++i) {
          const int s = i % WST, ph = (i / WST) & 1;
          mbar_wait(&wempty[s], ph ^ 1);
          mbar_expect_tx(&wfull[s], 2 * (uint32_t)n_tile * 128);
          tma_load_2d(wring + s * WSTAGE, &tmWhi, &wfull[s], kb * BKE, n0);
          tma_load_2d(wring + s * WSTAGE + W_BYTES, &tmWlo, &wfull[s], kb * BKE, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = H16 ? idesc_f16(BM, n_tile, 0, 0) : idesc_tf32(BM, n_tile, 0, 0);
      int i = 0, seg = 0;
      for (int ch = c0; ch < c1; ++ch, ++seg) {
        int tile, part, kb_b, kb_e;
        chunk_range(ch, tile, part, kb_b, kb_e);
        // Two TMEM accumulators per chunk, added by the epilogue in fp32 (round-to-nearest): the main
        // products A_hi W_hi and the small correction products A_lo W_hi + A_hi W_lo.  The tcgen05
        // accumulator is updated with truncation - about half an ulp OF THE ACCUMULATOR per MMA step -
        // so adding the 2^-12-sized corrections into the big sum tripled the bias (measured: worst
        // gradient error 2.8e-4 -> 9e-4); kept apart, the corrections meet a tiny ulp and the main sum
        // sees only a third of the steps.  (Cost: no accumulator double-buffering across chunks,
        // ~10 % of this kernel, the chunks are long.)
        const int aph = seg & 1;
        mbar_wait(&acc_empty[0], aph ^ 1);
        tc_fence_after();
        for (int kb = kb_b; kb < kb_e; ++kb, ++i) {
          const int s = i % XST, ph = (i / XST) & 1;
          const int ws = i % WST, wph = (i / WST) & 1;
          mbar_wait(&ready[s], ph);
          mbar_wait(&wfull[ws], wph);
          tc_fence_after();
          const uint32_t hi = smem_u32(xring + s * XSTAGE), lo = hi + X_BYTES;
          const uint32_t whi = smem_u32(wring + ws * WSTAGE), wlo = whi + W_BYTES;
#pragma unroll
          for (int pair = 0; pair < 3; ++pair) {
            const uint32_t sa = pair == 0 ? lo : hi;
            const uint32_t sb = pair == 1 ? wlo : whi;
            const uint32_t d_tmem = tmem_base + (uint32_t)(pair == 2 ? 0 : BN);
#pragma unroll
            for (int kk = 0; kk < BK / 8; ++kk) {  // 4 steps of 32 bytes of K per row
              if (H16) umma_f16(d_tmem, smem_desc(sa + kk * 32, 16, 1024), smem_desc(sb + kk * 32, 16, 1024), idesc,
                                (kb != kb_b) || kk != 0 || pair == 1);
              else umma_tf32(d_tmem, smem_desc(sa + kk * 32, 16, 1024), smem_desc(sb + kk * 32, 16, 1024), idesc,
                             (kb != kb_b) || kk != 0 || pair == 1);
            }
          }
          umma_commit(&empty[s]);
          umma_commit(&wempty[ws]);
        }
        umma_commit(&acc_full[0]);
      }
    }
  } else if (warp >= 2 && warp < (H16 ? 2 + TW : 6)) {
    // ===== transform warps: x tile -> (hi, lo) of relu(bn(sqrt(x))) in place =====
    const int tid = (warp - 2) * 32 + lane;
    const int t = tid % 128;      // row of the tile
    const int tg = tid / 128;     // H16: which 64 / TG genes of the row this thread converts
    constexpr int NTR = H16 ? 128 * TG : 128;
    constexpr int GPT = 64 / TG;  // genes per thread
    int i = 0;
    int pending_store_stage = -1;
    for (int ch = c0; ch < c1; ++ch) {
      int tile, part, kb_b, kb_e;
      chunk_range(ch, tile, part, kb_b, kb_e);
      const int m0 = (tile / n_tiles) * BM;
      const bool store_hi = (tile % n_tiles) == 0;
      for (int kb = kb_b; kb < kb_e; ++kb, ++i) {
        const int s = i % XST, ph = (i / XST) & 1;
        if (tid == 0 && pending_store_stage >= 0) {
          // the previous hi-tile store has finished reading its stage: release it
          tma_store_wait_read0();
          mbar_arrive(&empty[pending_store_stage]);
          pending_store_stage = -1;
        }
        mbar_wait(&full[s], ph);
        uint8_t* xt = xring + s * XSTAGE;
        uint8_t* lt = xt + X_BYTES;
        const int k0 = kb * BKE;
        if (H16) {
          // thread (tg, t): GPT genes of row t -> GPT / 8 chunks of the fp16 hi row (tile over raw box 0) and of
          // the fp16 lo row (tile over raw box 1).  Outputs land where other threads' raw values live, so
          // every raw value is in registers (barrier) before anything is written.
          float a[GPT];
          const int box = (tg * GPT) / 32, rc0 = ((tg * GPT) % 32) / 4;  // raw box, first raw 16-byte chunk
#pragma unroll
          for (int c = 0; c < GPT / 4; ++c) {
            const float4 xv = *reinterpret_cast<const float4*>(xt + box * X_BYTES + t * 128 + (((rc0 + c) ^ (t & 7)) << 4));
            a[4 * c] = xv.x; a[4 * c + 1] = xv.y; a[4 * c + 2] = xv.z; a[4 * c + 3] = xv.w;
          }
          named_bar_sync(2, NTR);
#pragma unroll
          for (int c = 0; c < GPT / 4; ++c) {
            const int col = k0 + tg * GPT + c * 4;
            if (col + 4 <= G) {
              const float4 mean = __ldg(reinterpret_cast<const float4*>(bn.mean + col));
              const float4 rstd = __ldg(reinterpret_cast<const float4*>(bn.rstd + col));
              const float4 gam = __ldg(reinterpret_cast<const float4*>(bn.gamma + col));
              const float4 bet = __ldg(reinterpret_cast<const float4*>(bn.beta + col));
              a[4 * c] = bn_relu(pre(a[4 * c]), mean.x, rstd.x, gam.x, bet.x);
              a[4 * c + 1] = bn_relu(pre(a[4 * c + 1]), mean.y, rstd.y, gam.y, bet.y);
              a[4 * c + 2] = bn_relu(pre(a[4 * c + 2]), mean.z, rstd.z, gam.z, bet.z);
              a[4 * c + 3] = bn_relu(pre(a[4 * c + 3]), mean.w, rstd.w, gam.w, bet.w);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                a[4 * c + j] = (col + j < G) ? bn_relu(pre(a[4 * c + j]), bn.mean[col + j], bn.rstd[col + j],
                                                       bn.gamma[col + j], bn.beta[col + j])
                                             : 0.f;
            }
          }
          // 8 genes -> one 16-byte chunk of hi and of lo
#pragma unroll
          for (int j = 0; j < GPT / 8; ++j) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const float a0 = a[8 * j + 2 * v], a1 = a[8 * j + 2 * v + 1];
              const __half2 hh = __floats2half2_rn(a0, a1);
              const float2 hf = __half22float2(hh);
              hw[v] = *reinterpret_cast<const uint32_t*>(&hh);
              const __half2 ll = __floats2half2_rn((a0 - hf.x) * LO_SCALE, (a1 - hf.y) * LO_SCALE);
              lw[v] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            const int lc = tg * (GPT / 8) + j;  // logical 16-byte chunk of the 64-gene fp16 row
            *reinterpret_cast<uint4*>(xt + t * 128 + ((lc ^ (t & 7)) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(lt + t * 128 + ((lc ^ (t & 7)) << 4)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
        }
#pragma unroll
        for (int c = 0; c < (H16 ? 0 : 8); ++c) {
          // logical 16-byte chunk c of row t lives at physical chunk c ^ (t % 8) (128-byte swizzle):
          // the 8 threads of a quarter-warp phase touch 8 different chunks -> conflict-free
          const int off = t * 128 + ((c ^ (t & 7)) << 4);
          const float4 xv = *reinterpret_cast<const float4*>(xt + off);
          const float in[4] = {xv.x, xv.y, xv.z, xv.w};
          float hv[4], lv[4];
          const int col = k0 + c * 4;
          if (col + 4 <= G) {
            const float4 mean = __ldg(reinterpret_cast<const float4*>(bn.mean + col));
            const float4 rstd = __ldg(reinterpret_cast<const float4*>(bn.rstd + col));
            const float4 gam = __ldg(reinterpret_cast<const float4*>(bn.gamma + col));
            const float4 bet = __ldg(reinterpret_cast<const float4*>(bn.beta + col));
            const float mv[4] = {mean.x, mean.y, mean.z, mean.w}, rv[4] = {rstd.x, rstd.y, rstd.z, rstd.w};
            const float gv[4] = {gam.x, gam.y, gam.z, gam.w}, bv[4] = {bet.x, bet.y, bet.z, bet.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float a = bn_relu(pre(in[j]), mv[j], rv[j], gv[j], bv[j]);
              hv[j] = rna_tf32_(a);
              lv[j] = rna_tf32_(a - hv[j]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float a = 0.f;
              if (col + j < G) a = bn_relu(pre(in[j]), bn.mean[col + j], bn.rstd[col + j], bn.gamma[col + j], bn.beta[col + j]);
              hv[j] = rna_tf32_(a);
              lv[j] = rna_tf32_(a - hv[j]);
            }
          }
          *reinterpret_cast<float4*>(xt + off) = make_float4(hv[0], hv[1], hv[2], hv[3]);
          *reinterpret_cast<float4*>(lt + off) = make_float4(lv[0], lv[1], lv[2], lv[3]);
        }
        fence_proxy_async();        // generic-proxy writes -> visible to tcgen05.mma / TMA store
        named_bar_sync(1, NTR);     // the whole tile is transformed
        if (tid == 0) {
          if (store_hi) {
            tma_store_2d(&tmAhi, xt, k0, m0);
            tma_store_commit();
            pending_store_stage = s;
          } else {
            mbar_arrive(&empty[s]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[s]);
      }
    }
    if (tid == 0) {
      if (pending_store_stage >= 0) {
        tma_store_wait_read0();
        mbar_arrive(&empty[pending_store_stage]);
      }
      tma_store_wait_all();  // global writes of the hi tiles complete before the kernel ends
    }
  } else if (warp >= TW + 2 && warp < TW + 6) {
    // ===== epilogue: TMEM -> slab of this k-part =====
    const int q = warp % 4;
    int seg = 0;
    for (int ch = c0; ch < c1; ++ch, ++seg) {
      int tile, part, kb_b, kb_e;
      chunk_range(ch, tile, part, kb_b, kb_e);
      const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * n_tile;
      const int aph = seg & 1;
      const int m = m0 + q * 32 + lane;
      const bool two = true;  // main + correction accumulators
      mbar_wait(&acc_full[0], aph);
      tc_fence_after();
      const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int c = 0; c < n_tile; c += 16) {
        float v[16];
        tmem_ld16(tb + (uint32_t)c, v);
        if (two) {
          float v2[16];
          tmem_ld16(tb + (uint32_t)(BN + c), v2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = H16 ? __fmaf_rn(v2[j], 1.f / LO_SCALE, v[j]) : v[j] + v2[j];
        } else {
          tmem_ld_wait();
        }
        if (part_stride == 0 && bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + c + j < h) v[j] += __ldg(&bias[n0 + c + j]);
        }
        if (stats_acc != nullptr) {
          // BatchNorm statistics of the output for the next block: column sums over this warp's 32 rows by a
          // transposing butterfly (16 shuffles per moment), then one fp64 atomic per column, moment and warp
          float s1[16], s2[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            s1[j] = m < B ? v[j] : 0.f;
            s2[j] = s1[j] * s1[j];
          }
          const int col = warp_colsum16(s1, lane);
          warp_colsum16(s2, lane);
          if ((lane & 1) == 0 && n0 + c + col < h) {
            atomicAdd(&stats_acc[n0 + c + col], (double)s1[0]);
            atomicAdd(&stats_acc[h + n0 + c + col], (double)s2[0]);
          }
        }
        if (m < B) {
          float* dst = out + (size_t)part * part_stride + (size_t)m * h + n0 + c;
          if (n0 + c + 16 <= h && ((((uintptr_t)dst) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n0 + c + j < h) dst[j] = v[j];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[0]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
  if (stats_acc != nullptr) {
    // the last CTA to get here turns the sums into mean / rstd and updates the running buffers
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(stats_fin.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (is_last) {
      __threadfence();
      for (int i = threadIdx.x; i < h; i += blockDim.x) {
        const double mu = __ldcg(&stats_acc[i]) / stats_fin.B;
        double var = __ldcg(&stats_acc[h + i]) / stats_fin.B - mu * mu;
        if (var < 0.0) var = 0.0;
        stats_fin.mean[i] = (float)mu;
        stats_fin.rstd[i] = 1.0f / sqrtf((float)(var + (double)stats_fin.eps));
        if (stats_fin.rmean) {
          const float unbiased = (float)var * ((float)stats_fin.B / (float)(stats_fin.B - 1));
          stats_fin.rmean[i] = (1.0f - stats_fin.momentum) * stats_fin.rmean[i] + stats_fin.momentum * (float)mu;
          stats_fin.rvar[i] = (1.0f - stats_fin.momentum) * stats_fin.rvar[i] + stats_fin.momentum * unbiased;
        }
      }
    }
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

// fp16 hi / scaled-lo split of the first Linear's weight ([h][G], once per step)
__global__ void k_split_h16(const float* __restrict__ src, size_t n2, __half2* __restrict__ hi, __half2* __restrict__ lo) {
  pdl_prologue();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(src) + i);
    const __half2 hh = __floats2half2_rn(v.x, v.y);
    const float2 hf = __half22float2(hh);
    hi[i] = hh;
    lo[i] = __floats2half2_rn((v.x - hf.x) * LO_SCALE, (v.y - hf.y) * LO_SCALE);
  }
}

bool encfwd_h16_supported(int G, int h) {
  static const bool off = getenv("DRV_ENC_TF32") != nullptr;  // debug / A-B: the tf32-operand kernel
  return !off && G % 8 == 0 && G >= 64 && h % 8 == 0;
}

void encfwd_split_weight_h16(const float* W, int h, int G, void* w_hi_h, void* w_lo_h, cudaStream_t st) {
  const size_t n2 = (size_t)h * G / 2;
  int blocks = (int)((n2 + 255) / 256);
  if (blocks > sms() * 8) blocks = sms() * 8;
  launch_k(k_split_h16, blocks, 256, 0, st, W, n2, reinterpret_cast<__half2*>(w_hi_h), reinterpret_cast<__half2*>(w_lo_h));
  note_launch();
}

int encfwd_fused(const float* x, int B, int G, BnParams bn, const float* w_hi, const float* w_lo, int h,
                 const float* bias, float* V, float* a_hi, float* partials, int max_parts, cudaStream_t st,
                 const ColStatsOut* stats, bool h16, bool sqrt_in) {
  CUtensorMap tX, tWh, tWl, tA;
  int n_tile = h >= BN ? BN : ((h + 15) / 16) * 16;
  const int bke = h16 ? 64 : BK;
  const int kb_total = (G + bke - 1) / bke;
  const int m_tiles = (B + BM - 1) / BM;
  // Short K (the h x h trunk blocks: <= 20 k-blocks, one accumulation chain): no k-split, hence no slabs and no
  // reduction pass - bias and the next block's BatchNorm statistics are finished in the epilogue.  Narrow N tiles
  // spread the few cell tiles over the SMs (every CTA of a cell tile transforms the same small input tile).
  const bool unsplit = kb_total <= 20 && h % 4 == 0;
  if (unsplit) {
    while (n_tile > 64 && n_tile % 32 == 0 && (long long)m_tiles * ((h + n_tile - 1) / n_tile) * 2 <= sms()) n_tile /= 2;
  }
  if (!tmap_2d(&tX, x, B, G, G, BM, BK)) return -3;
  if (h16) {  // w_hi / w_lo / a_hi are fp16 buffers
    if (!tmap_2d_h(&tWh, w_hi, h, G, G, (uint32_t)n_tile, 64)) return -3;
    if (!tmap_2d_h(&tWl, w_lo, h, G, G, (uint32_t)n_tile, 64)) return -3;
    if (!tmap_2d_h(&tA, a_hi, B, G, G, BM, 64)) return -3;
  } else {
    if (!tmap_2d(&tWh, w_hi, h, G, G, (uint32_t)n_tile, BK)) return -3;
    if (!tmap_2d(&tWl, w_lo, h, G, G, (uint32_t)n_tile, BK)) return -3;
    if (!tmap_2d(&tA, a_hi, B, G, G, BM, BK)) return -3;
  }
  const long long tiles = (long long)m_tiles * ((h + n_tile - 1) / n_tile);
  // k-parts: minimise waves x k-blocks per chunk (chunks beyond a multiple of the SM count cost a
  // whole extra wave), with accumulation chains <= 20 k-blocks (240 MMA steps, bias ~1e-6, see
  // DESIGN.md) and at most max_parts slabs
  int splits = 1;
  if (!unsplit) {
    long long best = -1;
    for (int sp = 1; sp <= max_parts && sp <= kb_total; ++sp) {
      const int per = (kb_total + sp - 1) / sp;
      if (per > 20 && sp < max_parts) continue;
      const long long waves = (tiles * sp + sms() - 1) / sms();
      const long long cost = waves * (per + 2);
      if (best < 0 || cost < best) { best = cost; splits = sp; }
    }
  }
  const long long chunks = tiles * splits;
  const int grid = (int)(chunks < sms() ? chunks : sms());
  auto kern = h16 ? (sqrt_in ? k_encfwd<true, true> : k_encfwd<true, false>) : (sqrt_in ? k_encfwd<false, true> : k_encfwd<false, false>);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  StatFinal nofin{};
  if (splits > 1) {
    launch_k(kern, grid, NTHREADS, SMEM_BYTES, st, tX, tWh, tWl, tA, bn, B, G, h, n_tile, splits, partials,
                                             (long long)B * h, (const float*)nullptr, (double*)nullptr, nofin);
    note_launch();
    return reduce_parts(partials, splits, B, h, bias, V, h, st, stats) ? 1 : 0;
  }
  const bool fuse_stats = stats != nullptr && B > 1;
  launch_k(kern, grid, NTHREADS, SMEM_BYTES, st, tX, tWh, tWl, tA, bn, B, G, h, n_tile, 1, V, 0LL, bias,
           fuse_stats ? stats->acc : (double*)nullptr, fuse_stats ? stats->fin : nofin);
  note_launch();
  return fuse_stats ? 1 : 0;
}

}  // namespace tc
}  // namespace drv
