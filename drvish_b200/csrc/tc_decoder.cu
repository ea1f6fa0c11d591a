// Fused decoder-last-layer GEMM + gene softmax + library scaling + NB likelihood on the sm_100a
// tensor cores (tcgen05.mma, TMEM accumulators, TMA-fed 128-byte-swizzled operands; kind::f16 with fp16
// operands by default - see DL_SCALE below - or kind::tf32 with fp32 containers, template flag H16).
// The (cells x 2*genes) logits / mean matrix never touches HBM in the forward pass:
//
//   phase 1  logits(scale half) tile in TMEM -> per-row online (max, sum exp) partials  -> lse[b]
//   phase 2  logits(both halves) tile in TMEM + TMA-staged count tile -> NB log-prob row sums
//            LL[b], A[b] = sum_g dLL/d ell  (= dLL/d library)                 [evaluate_loss stops here]
//            training: the NB part of d loss/d logits (ds', drho) leaves in the same pass through
//            shared-memory tiles and TMA stores
//   phase 3  recompute the scale-half tile, add the softmax coupling ds = ds' + A[b] softmax(s) in place
//            to the TMA-staged ds' tile, TMA store, bias gradient.
//            Hidden width <= 256 (k_dec3w, below): the same kernel feeds the finished tile, still in shared memory,
//            into the weight-gradient MMA dW = ds^T act (gene-major walk, accumulator in TMEM); otherwise (k_dec<3>)
//            dW = dlogits^T act is a tcgen05 GEMM.  dAct = dlogits W and the log_r half of dW are GEMMs that read
//            dlogits / act / W in the layouts they already have (MN-major descriptors, tc_gemm.cu).
//
// Reference arithmetic: NBDecoder.forward modules.py:93-96, "obs" site nbvae.py:72-78.
// Persistent CTAs (one per SM), 19 warps: 16 epilogue warps (4 per TMEM lane quarter, each a
// 16-gene column slice), a TMA producer warp, an MMA-issuer warp and a TMA store warp.  Tiles: 128 cells x
// 64 genes x 2 halves (phase 2) or 128 x 128 genes (phases 1, 3) = 128 accumulator columns,
// double-buffered in TMEM (256 columns).
#include <cuda_fp16.h>
#include <stdlib.h>
#include <type_traits>

#include "kernels.cuh"
#include "tc.cuh"
#include "tc_common.cuh"

extern "C" void drv_internal_kmark(int k, int which, void* stream);

namespace drv {

using namespace tc;

namespace {

constexpr int BM = 128;            // cells per tile
constexpr int BK = 32;             // k-block (fp32 elements)
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 4;       // 16 KB
constexpr int B_BYTES = 128 * BK * 4;      // 16 KB (128 weight rows)
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int X_BYTES = 2 * BM * 32 * 4;   // one count tile: two 128x32 boxes = 32 KB
constexpr int N_EPI_WARPS = 16;
// + TMA producer, MMA issuer, TMA store warp and one idle warp that completes the fifth warpgroup: setmaxnreg acts on
// whole warpgroups (phase 2 moves registers from the four service warps to the sixteen epilogue warps)
constexpr int NTHREADS = (N_EPI_WARPS + 4) * 32;
constexpr int EPI_REGS = 112, SERVICE_REGS = 32;  // 16 x 32 x 112 + 4 x 32 x 32 = 640 x 96, the launch-time allocation
// phase 2 (training): 3 operand stages + 2 count tiles (overwritten in place by ds') + 2 drho tiles
constexpr int SMEM_BYTES = 3 * STAGE_BYTES + 4 * X_BYTES + 1024 + 1024;
// phases 1, 2 (forward only) and 3 need less: leave room for the small kernels of the other stream
constexpr int SMEM_SMALL = STAGES * STAGE_BYTES + 2 * X_BYTES + 1024 + 1024;
static_assert(SMEM_SMALL >= 2 * STAGE_BYTES + 4 * X_BYTES + 1024 + 1024, "phase 3 layout");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
// fp16-operand variant (H16): act, W and d loss/d logits are stored as fp16 - the same 10-bit mantissa
// as the tf32 operands they replace, half the bytes through HBM / L2 / shared memory, 64-element
// k-blocks (half the pipeline hand-shakes) and twice the MMA rate.  d logits is stored divided by the
// plate scale c and multiplied by DL_SCALE, so that its magnitude is that of the per-element count
// residual whatever scale_factor is; the backward GEMMs multiply by c / DL_SCALE (exact powers of two).
// fp16 range with DL_SCALE = 4: residuals up to 1.6e4 are representable (larger ones saturate: cvt.satfinite),
// residuals down to 1.5e-5 keep the full 11-bit significand (below: 1.5e-8 absolute resolution).  The lower end
// matters: with a small dispersion parameter (r = e^rho ~ 3e-4, log_r ~ -8) EVERY entry is of size r, and the former
// scale of 1/16 put all of them into fp16's subnormal range (measured: decoder-trunk gradients off by 2.7e-3,
// tests/test_cuda_parity.py::test_dispersion_sweep[mid--8.0-False]).
constexpr float DL_SCALE = 4.0f;

// tcgen05 kind::tf32 TRUNCATES the low 13 mantissa bits of its fp32 operands, which biases every
// product by ~2^-11; rounding the operands to nearest when they are produced keeps the error
// unbiased (measured: gradient error 1.3e-3 -> see DESIGN.md).
__device__ __forceinline__ float rna_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

struct DecParams {
  int B, G, h;
  const float* bias;  // [2G]
  const float* lib;   // [B]
  const float* rshift;  // [B] or NULL (MCVNBVAE data-split shift of log_r, count parameter only)
  float eps, c;
  // phase 1
  float* part_m;      // [n_slots][B]
  float* part_s;
  // phase 2/3: lse[b] is combined from the phase-1 partials when a CTA enters a cell tile
  int lse_n_gt, lse_n_tiles, lse_grid;  // phase 1's gene tiles per cell tile, tile count and grid
  float* lse;         // [B] logsumexp per cell, written by phase 2 for the fused phase 3 (k_dec3w)
  float* ll;          // [B] (atomic accumulation, zeroed)
  float* asum;        // [B]
  float* dlogits;     // [B][2G]
  float* db;          // [2G] (atomic accumulation, zeroed)
  const unsigned long long* trace;  // non-null: k_dec3w stamps its timeline (debug)
};

// Column sums of 4 values over the 32 lanes of a warp by a transposing butterfly (6 shuffles instead
// of 20): afterwards w[0] of every lane with (lane & 7) == 0 holds the sum of column
// 2*bit4(lane) + bit3(lane); returns that column index or -1.
__device__ __forceinline__ int butterfly4(float (&w)[4], int lane) {
  const bool hi16 = lane & 16;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = hi16 ? w[i] : w[i + 2];
    const float keep = hi16 ? w[i + 2] : w[i];
    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  const bool hi8 = lane & 8;
  {
    const float send = hi8 ? w[0] : w[1];
    const float keep = hi8 ? w[1] : w[0];
    w[0] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  w[0] += __shfl_xor_sync(0xffffffffu, w[0], 4);
  w[0] += __shfl_xor_sync(0xffffffffu, w[0], 2);
  w[0] += __shfl_xor_sync(0xffffffffu, w[0], 1);
  return (lane & 7) == 0 ? ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1) : -1;
}

// logsumexp of row b from the (max, sum exp) partials phase 1 left per (CTA, column group) slot
__device__ __forceinline__ float combine_lse(const DecParams& p, int ct, int b) {
  const int first_cta = (int)((((long long)ct * p.lse_n_gt + 1) * p.lse_grid - 1) / p.lse_n_tiles);
  const int last_cta = (int)((((long long)ct * p.lse_n_gt + p.lse_n_gt) * p.lse_grid - 1) / p.lse_n_tiles);
  const int n_slots = (last_cta - first_cta + 1) * 4;
  float m = -INFINITY;
  for (int s = 0; s < n_slots; ++s) m = fmaxf(m, __ldcg(&p.part_m[(size_t)s * p.B + b]));
  float acc = 0.f;
  for (int s = 0; s < n_slots; ++s) {
    const float pmv = __ldcg(&p.part_m[(size_t)s * p.B + b]);
    if (pmv > -INFINITY) acc += __ldcg(&p.part_s[(size_t)s * p.B + b]) * __expf(pmv - m);
  }
  return m + logf(acc);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory accesses of the epilogue loops by 32-bit shared address: the tile pointers are derived from the
// run-time-aligned dynamic buffer, which the compiler treats as generic (LD / ST + 64-bit address arithmetic)
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

template <int PHASE, bool WITH_A, bool H16>
__global__ void __launch_bounds__(NTHREADS, 1)
k_dec(const __grid_constant__ CUtensorMap tmAct, const __grid_constant__ CUtensorMap tmW,
      const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmD,
      const __grid_constant__ CUtensorMap tmR, DecParams p) {
  pdl_prologue();
  // The gradient tiles leave through shared memory and TMA stores (TSTORE): a warp-wide 128-bit store
  // from the TMEM register layout touches 32 different rows - 32 half-filled sectors per instruction -
  // and phase 2 turned out to be bound by exactly those stores (a third store stream cost +50%).
  //   phase 2 (training): ds' overwrites the count tile in place, drho has its own two buffers; the
  //     operand ring shrinks to 3 stages to make room (the MMA is far from critical here)
  //   phase 3: stages a 128 x 128 ds' tile (4 boxes, 64 KB), 2-deep operand ring, ds written in place
  constexpr bool TSTORE = (PHASE == 2 && WITH_A) || PHASE == 3;
  // phase 1 moves TWO 32-float k-blocks per ring stage: the producer / MMA-thread hand-shake per stage,
  // not bytes or MMA time, is what bounds it (profiles/r01_notes.md), so fewer, fatter stages win
  constexpr int KG = PHASE == 1 ? 2 : 1;
  // H16 phase 3: the staged ds' tile is fp16 (32 KB), which leaves room for a 4-deep operand ring
  constexpr int NST = PHASE == 3 ? (H16 ? 4 : 2) : (TSTORE ? 3 : (PHASE == 1 ? 3 : STAGES));
  constexpr int XB = PHASE == 3 ? (H16 ? X_BYTES : 2 * X_BYTES) : (PHASE == 1 ? 0 : X_BYTES);  // phase 1 stages no count tile
  constexpr int RB = (PHASE == 2 && WITH_A) ? X_BYTES : 0;
  // H16 phase 2 (training): the fp16 ds' / drho tiles (16 KB each) have their own double buffer (rbuf); the
  // count tile is released by the epilogue warps, the output buffer by the store warp (o_empty)
  constexpr bool X_BY_STORE = TSTORE && !(H16 && PHASE == 2);
  constexpr int BKE = H16 ? 64 : BK;  // operand elements per k-block (one 128-byte swizzle row)
  // TMEM accumulator stages.  Four stages in phase 2 (all 512 columns, the MMAs two to three tiles ahead of the epilogue)
  // were tried because 48 % of the warp-tiles enter the spin loop on acc_full (ncu source page): no change, 110.3 us -
  // the polls are cheap, the epilogue's own instruction stream is the bound.
  constexpr int NACC = 2;
  constexpr int TM_COLS = NACC * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* xbuf = smem + NST * KG * STAGE_BYTES;
  uint8_t* rbuf = xbuf + 2 * XB;  // [2][RB]
  uint64_t* full = (uint64_t*)(rbuf + 2 * RB);
  uint64_t* empty = full + NST;
  uint64_t* acc_full = empty + NST;       // [NACC]
  uint64_t* acc_empty = acc_full + NACC;  // [NACC]
  uint64_t* x_full = acc_empty + NACC;    // [2]
  uint64_t* x_empty = x_full + 2;         // [2]
  uint64_t* o_full = x_empty + 2;         // [2] output tile complete in shared memory (TSTORE)
  uint64_t* o_empty = o_full + 2;         // [2] output tile read by the TMA store (H16 phase 2)
  uint32_t* tmem_slot = (uint32_t*)(o_empty + 2);
  __shared__ float lgtab[LGTAB_N];  // static: the compiler then knows the table lookups are shared-memory loads
  load_lgamma_table(lgtab, threadIdx.x);

  constexpr int TN = PHASE == 2 ? 64 : 128;  // genes per tile (phases 1 and 3 cover the scale half only)
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int n_gt = (p.G + TN - 1) / TN;
  const int n_ct = (p.B + BM - 1) / BM;
  const int n_tiles = n_gt * n_ct;
  const int t0 = (int)((long long)n_tiles * blockIdx.x / gridDim.x);
  const int t1 = (int)((long long)n_tiles * (blockIdx.x + 1) / gridDim.x);
  const int nkb = (p.h + BKE - 1) / BKE;  // TMA zero-fills a partial last k-block

  if (warp == N_EPI_WARPS && lane == 0) {
    tma_prefetch_desc(&tmAct);
    tma_prefetch_desc(&tmW);
    if (PHASE != 1) tma_prefetch_desc(&tmX);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < NACC; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], N_EPI_WARPS);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&x_empty[s], X_BY_STORE ? 1 : N_EPI_WARPS);  // released by the store warp / the epilogue warps
      mbar_init(&o_full[s], N_EPI_WARPS);
      mbar_init(&o_empty[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == N_EPI_WARPS + 1) tmem_alloc<TM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Phase 2 (training): the NB epilogue needs two element pairs and the next group's operands in registers, the
  // service warps need few - registers move from the fifth warpgroup to the four epilogue warpgroups.  The
  // setmaxnreg sits at the top of every role branch, so that ptxas allocates each branch under its own budget.
  constexpr bool REBALANCE = PHASE == 2 && WITH_A;
  auto service_regs = [] {
    if (REBALANCE) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SERVICE_REGS));
  };

  // phase 2 is bound by its epilogue warps: the service warps back off between barrier polls
  auto swait = [](uint64_t* bar, uint32_t parity) {
    if (PHASE == 2) mbar_wait_relaxed(bar, parity, 100);
    else mbar_wait(bar, parity);
  };
  if (warp == N_EPI_WARPS) {
    // ===== TMA producer =====
    service_regs();
    if (lane == 0) {
      int kit = 0;
      for (int t = t0, it = 0; t < t1; ++t, ++it) {
        const int ct = t / n_gt, gt = t % n_gt;
        const int b0 = ct * BM, g0 = gt * TN;
        if (PHASE != 1) {
          const int xs = it & 1, xph = (it >> 1) & 1;
          swait(&x_empty[xs], xph ^ 1);
          mbar_expect_tx(&x_full[xs], XB);
          constexpr int XW = (H16 && PHASE == 3) ? 64 : 32;  // elements per 128-byte box row
#pragma unroll
          for (int j = 0; j < XB / (BM * 128); ++j)
            tma_load_2d(xbuf + xs * XB + j * (BM * 128), &tmX, &x_full[xs], g0 + XW * j, b0);
        }
        for (int kb = 0; kb < nkb; kb += KG, ++kit) {
          const int s = kit % NST, ph = (kit / NST) & 1;
          swait(&empty[s], ph ^ 1);
          const int ng = min(KG, nkb - kb);
          mbar_expect_tx(&full[s], ng * STAGE_BYTES);
          for (int j = 0; j < ng; ++j) {
            uint8_t* sa = smem + (s * KG + j) * STAGE_BYTES;
            tma_load_2d(sa, &tmAct, &full[s], (kb + j) * BKE, b0);
            if (PHASE != 2) tma_load_2d(sa + A_BYTES, &tmW, &full[s], (kb + j) * BKE, g0);
            else tma_load_3d(sa + A_BYTES, &tmW, &full[s], (kb + j) * BKE, g0, 0);
          }
        }
      }
    }
  } else if (warp == N_EPI_WARPS + 1) {
    // ===== MMA issuer =====
    service_regs();
    if (lane == 0) {
      const uint32_t idesc = H16 ? idesc_f16(BM, 128, 0, 0) : idesc_tf32(BM, 128, 0, 0);
      int kit = 0;
      for (int t = t0, it = 0; t < t1; ++t, ++it) {
        const int as = it % NACC, aph = (it / NACC) & 1;
        swait(&acc_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * 128);
        for (int kb = 0; kb < nkb; kb += KG, ++kit) {
          const int s = kit % NST, ph = (kit / NST) & 1;
          swait(&full[s], ph);
          tc_fence_after();
          const int ng = min(KG, nkb - kb);
          for (int j = 0; j < ng; ++j) {
            const uint32_t sa = smem_u32(smem + (s * KG + j) * STAGE_BYTES);
            const uint32_t sb = sa + A_BYTES;
#pragma unroll
            for (int kk = 0; kk < BK / 8; ++kk) {  // 4 MMAs of 32 bytes of K per row (8 tf32 / 16 fp16 elements)
              if (H16) umma_f16(d_tmem, smem_desc(sa + kk * 32, 16, 1024), smem_desc(sb + kk * 32, 16, 1024), idesc,
                                (kb | j | kk) != 0);
              else umma_tf32(d_tmem, smem_desc(sa + kk * 32, 16, 1024), smem_desc(sb + kk * 32, 16, 1024), idesc,
                             (kb | j | kk) != 0);
            }
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[as]);
      }
    }
  } else if (warp == N_EPI_WARPS + 2) {
    // ===== TMA store warp: finished gradient tiles, shared memory -> global =====
    service_regs();
    if (TSTORE && lane == 0) {
      for (int t = t0, it = 0; t < t1; ++t, ++it) {
        const int ct = t / n_gt, gt = t % n_gt;
        const int b0 = ct * BM, g0 = gt * TN;
        const int xs = it & 1, ph = (it >> 1) & 1;
        swait(&o_full[xs], ph);  // the writers fenced their generic-proxy stores (fence.proxy.async)
        if (H16 && PHASE == 2) {
          // one 128 x 64 fp16 box each: ds' (first 16 KB of the slot), drho (second 16 KB)
          tma_store_2d(&tmD, rbuf + xs * RB, g0, b0);
          tma_store_2d(&tmR, rbuf + xs * RB + BM * 128, g0, b0);
        } else {
          constexpr int XW = H16 ? 64 : 32;
#pragma unroll
          for (int j = 0; j < XB / (BM * 128); ++j)
            tma_store_2d(&tmD, xbuf + xs * XB + j * (BM * 128), g0 + XW * j, b0);
          if (PHASE == 2) {
#pragma unroll
            for (int j = 0; j < RB / (BM * 128); ++j)
              tma_store_2d(&tmR, rbuf + xs * RB + j * (BM * 128), g0 + 32 * j, b0);
          }
        }
        tma_store_commit();
        tma_store_wait_read0();      // shared memory may be refilled; the global writes complete asynchronously
        mbar_arrive(X_BY_STORE ? &x_empty[xs] : &o_empty[xs]);
      }
      tma_store_wait_all();
    }
  } else if (warp > N_EPI_WARPS + 2) {
    service_regs();  // idle warp: completes the service warpgroup
  } else {
    // ===== epilogue warps =====
    if (REBALANCE) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPI_REGS));
    const int q = warp % 4;    // TMEM lane quarter
    const int cg = warp / 4;   // column group
    const int r = q * 32 + lane;
    f2 ll_acc2 = bc2(0.f), a_acc2 = bc2(0.f);  // phase 2: packed partial sums (both halves are added at the flush)
    float m_run = -INFINITY, s_run = 0.f;  // phase 1: online (max, sum exp) over this CTA's tiles of a cell tile
    auto flush_lse = [&](int ct_done, int row) {
      // slot = ordinal of this CTA among the CTAs whose tile range overlaps the cell tile
      const int first_cta = (int)((((long long)ct_done * n_gt + 1) * gridDim.x - 1) / n_tiles);
      const size_t slot = (size_t)((blockIdx.x - first_cta) * 4 + cg) * p.B + row;
      p.part_m[slot] = m_run;
      p.part_s[slot] = s_run;
    };
    // shared-space addresses and swizzle terms of this thread's row, made opaque so that the compiler keeps them in
    // registers (it otherwise re-derives them from %tid and the shared window in every trip of the hot loop)
    uint32_t xbuf_row_s = smem_u32(xbuf) + (uint32_t)((cg >> 1) * (BM * 128) + r * 128);
    uint32_t rbuf_row_s = smem_u32(rbuf) + (uint32_t)(r * 128);
    uint32_t r7 = (uint32_t)(r & 7);
    uint32_t c4 = (uint32_t)((cg & 1) * 4), cg2 = (uint32_t)(2 * cg), cg16 = (uint32_t)(cg * 16);
    asm volatile("" : "+r"(xbuf_row_s), "+r"(rbuf_row_s), "+r"(r7), "+r"(c4), "+r"(cg2), "+r"(cg16));
    int cur_ct = -1;
    float lib_b = 0.f, lse_b = 0.f, asum_b = 0.f, rs_b = 0.f;
    int b = 0;
    bool row_ok = false;
    auto flush_ll = [&]() {
      float s0, s1;
      un2(ll_acc2, s0, s1);
      atomicAdd(&p.ll[b], s0 + s1);
      if (WITH_A) {
        un2(a_acc2, s0, s1);
        atomicAdd(&p.asum[b], s0 + s1);
      }
    };
    for (int t = t0, it = 0; t < t1; ++t, ++it) {
      const int ct = t / n_gt, gt = t % n_gt;
      const int as = it % NACC, aph = (it / NACC) & 1;
      if (ct != cur_ct) {
        if (PHASE == 2 && cur_ct >= 0 && row_ok) flush_ll();
        if (PHASE == 1 && cur_ct >= 0 && row_ok) flush_lse(cur_ct, b);
        ll_acc2 = a_acc2 = bc2(0.f);
        m_run = -INFINITY;
        s_run = 0.f;
        cur_ct = ct;
        b = ct * BM + r;
        row_ok = b < p.B;
        if (PHASE != 1 && row_ok) {
          lib_b = p.lib[b];
          lse_b = combine_lse(p, ct, b);
          // every CTA that enters this cell tile computes the same bits: the duplicate stores are benign
          if (PHASE == 2 && WITH_A && cg == 0) p.lse[b] = lse_b;
          if (PHASE == 2) rs_b = p.rshift ? p.rshift[b] * 1.4426950408889634f : 0.f;
          if (PHASE == 3) asum_b = p.asum[b];
        }
        if (PHASE == 3 && !row_ok) asum_b = 0.f;
      }
      const int g0 = gt * TN;
      mbar_wait(&acc_full[as], aph);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 128);

      if (PHASE == 1) {
        // 32 genes of the scale half per warp: online (max, sum exp)
        float v0[16], v1[16];
        tmem_ld16(tbase + cg * 32, v0);
        tmem_ld16(tbase + cg * 32 + 16, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);
        const int gbase = g0 + cg * 32;
        float m = -INFINITY;
        if (gbase + 32 <= p.G) {  // interior: vector bias loads (warp-uniform addresses), no bounds selects
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 f0 = __ldg(reinterpret_cast<const float4*>(p.bias + gbase + 4 * j4));
            const float4 f1 = __ldg(reinterpret_cast<const float4*>(p.bias + gbase + 16 + 4 * j4));
            v0[4 * j4] += f0.x; v0[4 * j4 + 1] += f0.y; v0[4 * j4 + 2] += f0.z; v0[4 * j4 + 3] += f0.w;
            v1[4 * j4] += f1.x; v1[4 * j4 + 1] += f1.y; v1[4 * j4 + 2] += f1.z; v1[4 * j4 + 3] += f1.w;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) m = fmaxf(m, fmaxf(v0[j], v1[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            v0[j] = (gbase + j < p.G) ? v0[j] + __ldg(&p.bias[min(gbase + j, p.G - 1)]) : -INFINITY;
            v1[j] = (gbase + 16 + j < p.G) ? v1[j] + __ldg(&p.bias[min(gbase + 16 + j, p.G - 1)]) : -INFINITY;
            m = fmaxf(m, fmaxf(v0[j], v1[j]));
          }
        }
        float s = 0.f;
        if (m > -INFINITY) {
          const float L2E = 1.4426950408889634f, mm = m * L2E;  // exp(v - m) = 2^(v log2e - m log2e): one FFMA + MUFU
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            s0 += ex2_approx(__fmaf_rn(v0[j], L2E, -mm));
            s1 += ex2_approx(__fmaf_rn(v1[j], L2E, -mm));
          }
          s = s0 + s1;
        }
        if (m > -INFINITY) {
          const float M = fmaxf(m_run, m);
          s_run = s_run * __expf(m_run - M) + s * __expf(m - M);
          m_run = M;
        }
      } else if (PHASE == 3) {
        // backward, light phase: the NB part of d loss/d logits was written by phase 2; only the
        // softmax coupling is missing:  ds = ds' + c A[b] softmax(s)  (ds = a - p A).
        // This warp owns 32 genes of the 128-gene (scale half) tile, in 8 groups of 4 columns.
        const int gbase = g0 + cg * 32;
        const float cA = p.c * asum_b;
        // ds' tile: TMA-staged a tile ahead by the producer warp (box cg = this warp's 32 genes,
        // row r, 16-byte chunk g4 XOR-swizzled with r % 8)
        const int xs = it & 1, xph = (it >> 1) & 1;
        mbar_wait(&x_full[xs], xph);
        if (H16) {
          // staged ds' tile: fp16, pre-scaled by DL_SCALE / c; two boxes of 64 genes, this warp's 32 genes
          // are the 16-byte chunks (cg & 1) * 4 + j of box cg / 2.  ds = ds' + DL_SCALE A[b] softmax(s), in place.
          const float SA = DL_SCALE * asum_b, unscale = p.c * (1.f / DL_SCALE);
          uint8_t* drow_h = xbuf + xs * XB + (cg >> 1) * (BM * 128) + r * 128;
          float sv8_n[8];
          tmem_ld8(tbase + cg * 32, sv8_n);
          tmem_ld_wait();
#pragma unroll 1
          for (int j = 0; j < 4; ++j) {
            float sv[8], bs[8], ds[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) sv[v] = sv8_n[v];
            if (j + 1 < 4) tmem_ld8(tbase + cg * 32 + (j + 1) * 8, sv8_n);
            const int gq = gbase + 8 * j;  // G % 8 == 0 on this path: a group of 8 genes is all-valid or all-out
            const bool gok = gq < p.G;
            if (gok) {
              const float4 f0 = __ldg(reinterpret_cast<const float4*>(p.bias + gq));
              const float4 f1 = __ldg(reinterpret_cast<const float4*>(p.bias + gq + 4));
              bs[0] = f0.x; bs[1] = f0.y; bs[2] = f0.z; bs[3] = f0.w; bs[4] = f1.x; bs[5] = f1.y; bs[6] = f1.z; bs[7] = f1.w;
            } else {
#pragma unroll
              for (int v = 0; v < 8; ++v) bs[v] = 0.f;
            }
            uint4* cell = reinterpret_cast<uint4*>(drow_h + ((((cg & 1) * 4 + j) ^ (r & 7)) * 16));
            const uint4 raw = *cell;
            const uint32_t rw[4] = {raw.x, raw.y, raw.z, raw.w};
            const bool ok = row_ok && gok;
            uint32_t ow[4];
#pragma unroll
            for (int v2 = 0; v2 < 4; ++v2) {
              const float2 d2 = __half22float2(*reinterpret_cast<const __half2*>(&rw[v2]));
              const float p0 = ex2_approx((sv[2 * v2] + bs[2 * v2] - lse_b) * 1.4426950408889634f);
              const float p1 = ex2_approx((sv[2 * v2 + 1] + bs[2 * v2 + 1] - lse_b) * 1.4426950408889634f);
              ds[2 * v2] = ok ? __fmaf_rn(SA, p0, d2.x) : 0.f;
              ds[2 * v2 + 1] = ok ? __fmaf_rn(SA, p1, d2.y) : 0.f;
              ow[v2] = pack_h2_sat(ds[2 * v2], ds[2 * v2 + 1]);
            }
            *cell = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            {
              float w0[4] = {ds[0], ds[1], ds[2], ds[3]}, w1[4] = {ds[4], ds[5], ds[6], ds[7]};
              const int col = butterfly4(w0, lane);
              butterfly4(w1, lane);
              if (col >= 0 && gok) {
                atomicAdd(&p.db[gq + col], w0[0] * unscale);
                atomicAdd(&p.db[gq + 4 + col], w1[0] * unscale);
              }
            }
            tmem_ld_wait();
          }
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&acc_empty[as]);
            mbar_arrive(&o_full[xs]);
          }
          continue;
        }
        uint8_t* drow_s = xbuf + xs * XB + cg * (BM * 128) + r * 128;
        float sv_n[4];
        tmem_ld4(tbase + cg * 32, sv_n);
        tmem_ld_wait();
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4) {
          float sv[4], bs[4], ds[4];
#pragma unroll
          for (int v = 0; v < 4; ++v) sv[v] = sv_n[v];
          if (g4 + 1 < 8) tmem_ld4(tbase + cg * 32 + (g4 + 1) * 4, sv_n);  // next group, in flight during this one
          const int gq = gbase + 4 * g4;
          const bool vec = gq + 4 <= p.G;
          if (vec) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(p.bias + gq));
            bs[0] = f.x; bs[1] = f.y; bs[2] = f.z; bs[3] = f.w;
          } else {
#pragma unroll
            for (int v = 0; v < 4; ++v) bs[v] = (gq + v < p.G) ? __ldg(&p.bias[gq + v]) : 0.f;
          }
          const float4 dld = *reinterpret_cast<const float4*>(drow_s + ((g4 ^ (r & 7)) * 16));
          const float dsp[4] = {dld.x, dld.y, dld.z, dld.w};
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const bool ok = row_ok && (gq + v < p.G);
            const float pr = ex2_approx((sv[v] + bs[v] - lse_b) * 1.4426950408889634f);
            ds[v] = ok ? rna_tf32(__fmaf_rn(cA, pr, dsp[v])) : 0.f;
          }
          // in place over the staged ds' chunk; the store warp ships the tile (TMA clips rows >= B, genes >= G)
          *reinterpret_cast<float4*>(drow_s + ((g4 ^ (r & 7)) * 16)) = make_float4(ds[0], ds[1], ds[2], ds[3]);
          const int col = butterfly4(ds, lane);
          if (col >= 0 && gq + col < p.G) atomicAdd(&p.db[gq + col], ds[0]);
          tmem_ld_wait();
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&acc_empty[as]);
          mbar_arrive(&o_full[xs]);
        }
      } else {
        // phase 2: this warp owns 16 genes of the 64-gene tile, processed as a ROLLED loop over 4
        // groups of 4 columns (compact code: the fully unrolled epilogue thrashed the instruction
        // cache); nb_elems<.,4> provides the instruction-level parallelism.  In training mode
        // (WITH_A) the same pass also emits the NB part of d loss/d logits, so the transcendental
        // work is done once:  ds' = -c a (fp32, completed by phase 3),  drho = -c dLL/drho (final).
        const int xs = it & 1, xph = (it >> 1) & 1;
        mbar_wait(&x_full[xs], xph);
        if (H16 && WITH_A) mbar_wait(&o_empty[xs], xph ^ 1);  // the TMA store of this slot's previous tile has read it
        const int gbase = g0 + cg * 16;
        const f2 lml2 = bc2(lib_b - lse_b);
        // count tile: box (cg/2), row r; fp16 output tiles: ds' slot, drho slot = + BM * 128 (H16); fp32 tiles: ds'
        // over the counts, drho in rbuf
        const uint32_t xrow_s = xbuf_row_s + (uint32_t)(xs * XB);
        const uint32_t orow_s = rbuf_row_s + (uint32_t)(xs * RB);
        const uint32_t rrow_s = orow_s + (uint32_t)((cg >> 1) * (BM * 128));
        const float* bias_s = p.bias + gbase;  // scale half; the log_r half is p.G floats further
        const bool full_cols = gbase + 16 <= p.G;
        float sv_n[4], rv_n[4];
        const uint32_t tcol = tbase + cg16;  // this warp's 16 scale columns; the log_r columns are 64 further
        tmem_ld4(tcol, sv_n);
        tmem_ld4(tcol + 64, rv_n);
        tmem_ld_wait();
        // interior tiles (all 128 cells and 64 genes valid) take a copy of the loop without the validity
        // selects; TMA zero-fills / clips the edges, so only the edge tiles need masking for the sums
        auto tile_loop = [&](auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        // count chunk and bias values of column group g: 16-byte chunk (cg%2)*4 + g of the row, XOR-swizzled with r%8
        auto load_inputs = [&](int g, float4& xf, float4& fs, float4& fr) {
          xf = lds128(xrow_s + (((c4 + (uint32_t)g) ^ r7) * 16));
          if (!MASKED || full_cols) {  // warp-uniform addresses -> broadcast loads
            fs = __ldg(reinterpret_cast<const float4*>(bias_s + 4 * g));
            fr = __ldg(reinterpret_cast<const float4*>(bias_s + p.G + 4 * g));
          } else {
            const int gq = gbase + 4 * g;
            fs.x = gq + 0 < p.G ? __ldg(&p.bias[gq + 0]) : 0.f; fr.x = gq + 0 < p.G ? __ldg(&p.bias[p.G + gq + 0]) : 0.f;
            fs.y = gq + 1 < p.G ? __ldg(&p.bias[gq + 1]) : 0.f; fr.y = gq + 1 < p.G ? __ldg(&p.bias[p.G + gq + 1]) : 0.f;
            fs.z = gq + 2 < p.G ? __ldg(&p.bias[gq + 2]) : 0.f; fr.z = gq + 2 < p.G ? __ldg(&p.bias[p.G + gq + 2]) : 0.f;
            fs.w = gq + 3 < p.G ? __ldg(&p.bias[gq + 3]) : 0.f; fr.w = gq + 3 < p.G ? __ldg(&p.bias[p.G + gq + 3]) : 0.f;
          }
        };
        // software pipeline: the TMEM loads, the count chunk and the bias values of group g + 1 are in flight while
        // group g is evaluated (their latencies sat at the head of every trip: long / short scoreboard stalls)
        float4 xf_n, fs_n, fr_n;
        load_inputs(0, xf_n, fs_n, fr_n);
#pragma unroll 1
        for (int g4 = 0; g4 < 4; ++g4) {
          float sv[4], rv[4], bs[4], br[4];
#pragma unroll
          for (int v = 0; v < 4; ++v) { sv[v] = sv_n[v]; rv[v] = rv_n[v]; }
          float4 xf = xf_n;
          bs[0] = fs_n.x; bs[1] = fs_n.y; bs[2] = fs_n.z; bs[3] = fs_n.w;
          br[0] = fr_n.x; br[1] = fr_n.y; br[2] = fr_n.z; br[3] = fr_n.w;
          {
            const int gn = g4 + 1 < 4 ? g4 + 1 : g4;
            tmem_ld4(tcol + gn * 4, sv_n);
            tmem_ld4(tcol + 64 + gn * 4, rv_n);
            load_inputs(gn, xf_n, fs_n, fr_n);
          }
          const uint32_t chunk = (c4 + (uint32_t)g4) ^ r7;
          const int gq = gbase + 4 * g4;
          bool ok[4];
          if (MASKED) {
            // edge tile: out-of-range entries become benign (lsl = rho = x = 0) and are dropped from the sums
#pragma unroll
            for (int v = 0; v < 4; ++v) ok[v] = row_ok && (gq + v < p.G);
            if (!ok[0]) { sv[0] = bs[0] = rv[0] = br[0] = xf.x = 0.f; }
            if (!ok[1]) { sv[1] = bs[1] = rv[1] = br[1] = xf.y = 0.f; }
            if (!ok[2]) { sv[2] = bs[2] = rv[2] = br[2] = xf.z = 0.f; }
            if (!ok[3]) { sv[3] = bs[3] = rv[3] = br[3] = xf.w = 0.f; }
          }
          f2 lsl[2], rho[2], xv[2], ll[2], av[2], nd[2];
          const f2 lmlA = MASKED ? mk2(ok[0] ? lib_b - lse_b : 0.f, ok[1] ? lib_b - lse_b : 0.f) : lml2;
          const f2 lmlB = MASKED ? mk2(ok[2] ? lib_b - lse_b : 0.f, ok[3] ? lib_b - lse_b : 0.f) : lml2;
          lsl[0] = add2(add2(mk2(sv[0], sv[1]), mk2(bs[0], bs[1])), lmlA);
          lsl[1] = add2(add2(mk2(sv[2], sv[3]), mk2(bs[2], bs[3])), lmlB);
          rho[0] = add2(mk2(rv[0], rv[1]), mk2(br[0], br[1]));
          rho[1] = add2(mk2(rv[2], rv[3]), mk2(br[2], br[3]));
          xv[0] = mk2(xf.x, xf.y);
          xv[1] = mk2(xf.z, xf.w);
          av[0] = av[1] = nd[0] = nd[1] = bc2(0.f);
          bool fast = nb_pair<WITH_A>(lsl[0], rho[0], xv[0], p.eps, lgtab, rs_b, ll[0], av[0], nd[0]);
          fast &= nb_pair<WITH_A>(lsl[1], rho[1], xv[1], p.eps, lgtab, rs_b, ll[1], av[1], nd[1]);
          if (!fast) {
            // rare: counts >= LGTAB_N / non-integer, or r >= NB_R_BIG - those elements are redone by the generic route
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float l2[2], r2[2], x2[2], o_ll[2], o_a[2], o_nd[2];
              un2(lsl[h], l2[0], l2[1]); un2(rho[h], r2[0], r2[1]); un2(xv[h], x2[0], x2[1]);
              un2(ll[h], o_ll[0], o_ll[1]); un2(av[h], o_a[0], o_a[1]); un2(nd[h], o_nd[0], o_nd[1]);
#pragma unroll
              for (int v = 0; v < 2; ++v) {
                const float rr = exp2f(__fmaf_rn(r2[v], 1.4426950408889634f, rs_b)) + p.eps;
                if (!nb_fast_domain(x2[v], rr)) {  // (scalar form of the test: this block is off the hot path)
                  const float3 g = nb_elem_generic<WITH_A>(l2[v], r2[v], x2[v], p.eps, rs_b);
                  o_ll[v] = g.x; o_a[v] = g.y; o_nd[v] = g.z;
                }
              }
              ll[h] = mk2(o_ll[0], o_ll[1]); av[h] = mk2(o_a[0], o_a[1]); nd[h] = mk2(o_nd[0], o_nd[1]);
            }
          }
          if (MASKED) {
            float t0, t1, t2, t3;
            un2(ll[0], t0, t1); un2(ll[1], t2, t3);
            ll[0] = mk2(ok[0] ? t0 : 0.f, ok[1] ? t1 : 0.f); ll[1] = mk2(ok[2] ? t2 : 0.f, ok[3] ? t3 : 0.f);
            if (WITH_A) {
              un2(av[0], t0, t1); un2(av[1], t2, t3);
              av[0] = mk2(ok[0] ? t0 : 0.f, ok[1] ? t1 : 0.f); av[1] = mk2(ok[2] ? t2 : 0.f, ok[3] ? t3 : 0.f);
              un2(nd[0], t0, t1); un2(nd[1], t2, t3);
              nd[0] = mk2(ok[0] ? t0 : 0.f, ok[1] ? t1 : 0.f); nd[1] = mk2(ok[2] ? t2 : 0.f, ok[3] ? t3 : 0.f);
            }
          }
          ll_acc2 = add2(ll_acc2, add2(ll[0], ll[1]));
          if (WITH_A) a_acc2 = add2(a_acc2, add2(av[0], av[1]));
          if (WITH_A && H16) {
            // fp16 tiles, pre-scaled by DL_SCALE / c: ds' = -DL_SCALE a, drho' = -DL_SCALE dLL/drho = DL_SCALE ndrho;
            // 4 genes = 8 bytes of chunk 2 cg + g4 / 2 of this row
            float d[4], q[4];
            un2(mul2(av[0], bc2(-DL_SCALE)), d[0], d[1]);
            un2(mul2(av[1], bc2(-DL_SCALE)), d[2], d[3]);
            un2(mul2(nd[0], bc2(DL_SCALE)), q[0], q[1]);
            un2(mul2(nd[1], bc2(DL_SCALE)), q[2], q[3]);
            const uint32_t ocell = orow_s + (((cg2 + (uint32_t)(g4 >> 1)) ^ r7) * 16) + (uint32_t)((g4 & 1) * 8);
            sts64(ocell, pack_h2_sat(d[0], d[1]), pack_h2_sat(d[2], d[3]));
            sts64(ocell + BM * 128, pack_h2_sat(q[0], q[1]), pack_h2_sat(q[2], q[3]));
            // bias gradient of the log_r half: column sums over the 32 rows of this warp
            const int col = butterfly4(q, lane);
            if (col >= 0 && (!MASKED || gq + col < p.G)) atomicAdd(&p.db[p.G + gq + col], q[0] * (p.c * (1.f / DL_SCALE)));
          } else if (WITH_A) {
            float d[4], q[4];
            un2(mul2(av[0], bc2(-p.c)), d[0], d[1]);
            un2(mul2(av[1], bc2(-p.c)), d[2], d[3]);
            un2(mul2(nd[0], bc2(p.c)), q[0], q[1]);
            un2(mul2(nd[1], bc2(p.c)), q[2], q[3]);
#pragma unroll
            for (int v = 0; v < 4; ++v) q[v] = rna_tf32(q[v]);  // tf32 MMA operand of dW / dAct
            // ds' overwrites the counts it was computed from (same swizzled chunk), drho goes to its own
            // tile; the store warp ships both (TMA clips rows >= B and genes >= G)
            sts128(xrow_s + chunk * 16, d[0], d[1], d[2], d[3]);
            sts128(rrow_s + chunk * 16, q[0], q[1], q[2], q[3]);
            const int col = butterfly4(q, lane);
            if (col >= 0 && (!MASKED || gq + col < p.G)) atomicAdd(&p.db[p.G + gq + col], q[0]);
          }
          tmem_ld_wait();
        }
        };
        if (ct * BM + BM <= p.B && g0 + TN <= p.G) tile_loop(std::false_type{});
        else tile_loop(std::true_type{});
        if (TSTORE) fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&acc_empty[as]);
          if (TSTORE) mbar_arrive(&o_full[xs]);
          if (!X_BY_STORE) mbar_arrive(&x_empty[xs]);
        }
      }
    }
    if (PHASE == 2 && cur_ct >= 0 && row_ok) flush_ll();
    if (PHASE == 1 && cur_ct >= 0 && row_ok) flush_lse(cur_ct, b);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == N_EPI_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc<TM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Fused phase 3 + weight gradient (north star item 3: "the fused backward emits dlogits directly into the
// decoder weight-gradient GEMM").  The kernel that makes d loss / d logits(scale half) final -
//     ds = ds' + DL_SCALE A[b] softmax(s)           (ds = a - p A)
// - also consumes the finished fp16 tile, still in shared memory, as an operand of the weight-gradient MMA
//     dW_s[g][i] = sum_b ds[b][g] act[b][i]
// so the separate dW GEMM of the scale half and its 41 MB re-read of d logits disappear.
//
// Walk: GENE-major.  A CTA owns a contiguous range of (128-gene slab, 64-cell tile) tiles; the weight slab stays in
// shared memory and the dW accumulator (128 genes x h, fp32) stays in TMEM across the cell tiles of a slab, so dW is
// flushed (red.global.add) once per slab segment - twice per CTA - instead of once per tile.
// Per tile (64 cells x 128 genes):
//   MMA1  logits^T[128 genes][64 cells] = W_slab (A, K-major) act^T (B, K-major)                  -> TMEM D1 (2 stages)
//   epilogue warps: TMEM lane = gene, 16 cells per warp; ds' from the TMA-staged fp16 tile [cells][genes],
//                   ds written back in place (fp16), bias gradient summed in a register per gene
//   MMA2  dW[128 genes][h] += ds^T (A, MN-major: the staged tile as it lies) act (B, MN-major: the SAME act boxes
//                   MMA1 read as its K-major B operand)                                           -> TMEM D2
//   TMA store of the finished ds tile (the dAct GEMM reads it; that product accumulates over genes and cannot stay here)
// Shared memory: W slab 64 KB + 3 x (act tile 32 KB + ds tile 16 KB) = 208 KB; TMEM: 2 x 64 + 256 columns.
// Needs h <= 256 (accumulator columns; h = 512 keeps phase 3 + the dW GEMM).
constexpr int W3_SLAB = 128;                 // genes per slab
constexpr int W3_BM = 64;                    // cells per tile
constexpr int W3_NST = 3;                    // act / ds tile stages
constexpr int W3_WBYTES = 4 * W3_SLAB * 128; // weight slab: up to 4 k-blocks of [128 genes][64 h] fp16
constexpr int W3_ABYTES = 4 * W3_BM * 128;   // act tile: up to 4 boxes of [64 cells][64 h] fp16
constexpr int W3_DBYTES = 2 * W3_BM * 128;   // ds tile: 2 boxes of [64 cells][64 genes] fp16
constexpr int W3_CBYTES = 2 * W3_BM * 4;         // per-cell terms of a tile: lse[64], A[64] (fp32)
constexpr int W3_SMEM = W3_WBYTES + W3_NST * (W3_ABYTES + W3_DBYTES + W3_CBYTES) + 1024 + 512;
static_assert(W3_SMEM <= 227 * 1024, "shared memory");
constexpr int W3_D2_COL = 128;               // TMEM column of the dW accumulator (D1 stages at 0 and 64)

template <int OFF>
__device__ __forceinline__ unsigned short lds_u16_off(uint32_t a) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1+%2];" : "=h"(v) : "r"(a), "n"(OFF) : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ void sts_h_sat_off(uint32_t a, float f) {
  unsigned short v;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(v) : "f"(f));
  asm volatile("st.shared.u16 [%0+%2], %1;" ::"r"(a), "h"(v), "n"(OFF) : "memory");
}
// 1-D bulk copy global -> shared (bytes: multiple of 16, both addresses 16-byte aligned), completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Debug timeline of CTA 0 (DRV_W3_TRACE=1): globaltimer stamps per tile and role, read back with drv_debug_w3_trace
constexpr int W3_TRACE_TILES = 24, W3_TRACE_EVENTS = 10;
__device__ unsigned long long g_w3_trace[W3_TRACE_TILES * W3_TRACE_EVENTS];
__device__ __forceinline__ void w3_stamp(const unsigned long long* on, int it, int ev) {
  if (on && blockIdx.x == 0 && it < W3_TRACE_TILES) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_w3_trace[it * W3_TRACE_EVENTS + ev] = t;
  }
}

// kernel-level stamps of the first / middle / last CTA: slots [tile 21..23][entry, set-up done, walk done, exit]
__device__ __forceinline__ void w3_stamp_cta(const unsigned long long* on, int ev) {
  if (!on || threadIdx.x != 0) return;
  const int slot = blockIdx.x == 0 ? 21 : (blockIdx.x == gridDim.x / 2 ? 22 : (blockIdx.x == gridDim.x - 1 ? 23 : -1));
  if (slot < 0) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  g_w3_trace[slot * W3_TRACE_EVENTS + ev] = t;
}

__global__ void __launch_bounds__(NTHREADS, 1)
k_dec3w(const __grid_constant__ CUtensorMap tmAct, const __grid_constant__ CUtensorMap tmW,
        const __grid_constant__ CUtensorMap tmD, DecParams p, float* __restrict__ dW) {
  w3_stamp_cta(p.trace, 0);
  pdl_prologue();
  w3_stamp_cta(p.trace, 1);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* wbuf = smem;
  uint8_t* abuf = wbuf + W3_WBYTES;            // [W3_NST][W3_ABYTES]
  uint8_t* dbuf = abuf + W3_NST * W3_ABYTES;   // [W3_NST][W3_DBYTES]
  uint8_t* cbuf = dbuf + W3_NST * W3_DBYTES;   // [W3_NST][lse[64] | A[64]]
  uint64_t* bars = (uint64_t*)(cbuf + W3_NST * W3_CBYTES);
  uint64_t* w_full = bars;                     // weight slab landed
  uint64_t* w_empty = bars + 1;                // every MMA1 of the slab segment has read it
  uint64_t* act_full = bars + 2;               // [3]
  uint64_t* act_empty = act_full + W3_NST;     // [3] MMA2 of the tile has read the act stage
  uint64_t* x_full = act_empty + W3_NST;       // [3] ds' tile landed
  uint64_t* d_empty = x_full + W3_NST;         // [3] MMA2 and the TMA store have read the ds stage (2 arrivals)
  uint64_t* o_full = d_empty + W3_NST;         // [3] ds tile final in shared memory (16 warp arrivals)
  uint64_t* acc_full = o_full + W3_NST;        // [2] MMA1 complete
  uint64_t* acc_empty = acc_full + 2;          // [2] epilogue has read D1 (16 warp arrivals)
  uint64_t* d2_full = acc_empty + 2;           // dW accumulator of the segment complete
  uint64_t* d2_empty = d2_full + 1;            // ... flushed (16 warp arrivals)
  uint32_t* tmem_slot = (uint32_t*)(d2_empty + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int n_slabs = (p.G + W3_SLAB - 1) / W3_SLAB;
  const int n_ct = (p.B + W3_BM - 1) / W3_BM;
  const int n_tiles = n_slabs * n_ct;
  const int t0 = (int)((long long)n_tiles * blockIdx.x / gridDim.x);
  const int t1 = (int)((long long)n_tiles * (blockIdx.x + 1) / gridDim.x);
  const int nb = (p.h + 63) / 64;  // 64-element k-blocks of h (TMA zero-fills a partial one)

  if (warp == N_EPI_WARPS && lane == 0) {
    tma_prefetch_desc(&tmAct);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmD);
    mbar_init(w_full, 1);
    mbar_init(w_empty, 1);
    for (int s = 0; s < W3_NST; ++s) {
      mbar_init(&act_full[s], 1);
      mbar_init(&act_empty[s], 1);
      mbar_init(&x_full[s], 1);
      mbar_init(&d_empty[s], 2);
      mbar_init(&o_full[s], N_EPI_WARPS);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], N_EPI_WARPS);
    }
    mbar_init(d2_full, 1);
    mbar_init(d2_empty, N_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == N_EPI_WARPS + 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  w3_stamp_cta(p.trace, 2);

  auto seg_first = [&](int t) { return t == t0 || t % n_ct == 0; };
  auto seg_last = [&](int t) { return t == t1 - 1 || t % n_ct == n_ct - 1; };

  if (warp == N_EPI_WARPS) {
    // ===== TMA producer =====
    if (lane == 0) {
      int seg = 0;
      for (int t = t0, it = 0; t < t1; ++t, ++it) {
        const int slab = t / n_ct, ct = t % n_ct;
        const int g0 = slab * W3_SLAB, b0 = ct * W3_BM;
        if (seg_first(t)) {
          mbar_wait(w_empty, (seg & 1) ^ 1);
          mbar_expect_tx(w_full, nb * (W3_SLAB * 128));
          for (int j = 0; j < nb; ++j) tma_load_2d(wbuf + j * (W3_SLAB * 128), &tmW, w_full, j * 64, g0);
          ++seg;
        }
        const int s = it % W3_NST, ph = (it / W3_NST) & 1;
        mbar_wait(&act_empty[s], ph ^ 1);
        w3_stamp(p.trace, it, 0);
        mbar_expect_tx(&act_full[s], nb * (W3_BM * 128));
        for (int j = 0; j < nb; ++j) tma_load_2d(abuf + s * W3_ABYTES + j * (W3_BM * 128), &tmAct, &act_full[s], j * 64, b0);
        mbar_wait(&d_empty[s], ph ^ 1);
        w3_stamp(p.trace, it, 1);
        // per-cell terms of the tile (lse, A): whole 16-byte pieces; a ragged last piece reads into the 256-byte padding
        // of the workspace arrays, the epilogue masks those cells
        const uint32_t cb = (uint32_t)(((min(W3_BM, p.B - b0) + 3) / 4) * 16);
        mbar_expect_tx(&x_full[s], W3_DBYTES + 2 * cb);
        for (int j = 0; j < 2; ++j) tma_load_2d(dbuf + s * W3_DBYTES + j * (W3_BM * 128), &tmD, &x_full[s], g0 + 64 * j, b0);
        bulk_load_1d(cbuf + s * W3_CBYTES, p.lse + b0, cb, &x_full[s]);
        bulk_load_1d(cbuf + s * W3_CBYTES + W3_BM * 4, p.asum + b0, cb, &x_full[s]);
      }
    }
  } else if (warp == N_EPI_WARPS + 1) {
    // ===== MMA issuer =====
    // Order per trigger "epilogue of tile t - 1 done" (o_full): MMA2(t - 1) first - its commit frees the act / ds stage
    // for the TMA as early as possible - then MMA1(t + 1), which runs on the tensor pipe while the epilogue warps work
    // on tile t.  (tools/w3_trace.py: MMA1 = 16 N=64 MMAs takes 0.8 us, MMA2 0.4 us, the epilogue 0.7 us per tile.)
    // warp-uniform control flow (all lanes wait on the barriers), one elected lane issues: inside a divergent
    // `if (lane == 0)` the compiler wraps every tcgen05 instruction in a uniformity loop (ELECT / BRA.U.ANY)
    {
      const bool leader = elect_one();
      const uint32_t idesc1 = idesc_f16(128, W3_BM, 0, 0);
      const uint32_t idesc2 = idesc_f16(128, nb * 64, 1, 1);
      const uint32_t d2_tmem = tmem_base + (uint32_t)W3_D2_COL;
      int segW = 0, segD = 0;
      auto mma1 = [&](int it) {
        const int t = t0 + it;
        if (seg_first(t)) {
          mbar_wait(w_full, segW & 1);
          ++segW;
        }
        const int as = it & 1, aph = (it >> 1) & 1;
        const int s = it % W3_NST, ph = (it / W3_NST) & 1;
        mbar_wait(&acc_empty[as], aph ^ 1);
        mbar_wait(&act_full[s], ph);
        if (lane == 0) w3_stamp(p.trace, it, 2);
        tc_fence_after();
        const uint32_t d1 = tmem_base + (uint32_t)(as * W3_BM);
        const DescHalves da = desc_halves(smem_u32(wbuf), 16, 1024), db = desc_halves(smem_u32(abuf + s * W3_ABYTES), 16, 1024);
        const bool sl = seg_last(t);
        if (leader) {
          for (int j = 0; j < nb; ++j) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16(d1, desc_at(da, (uint32_t)(j * (W3_SLAB * 128) + kk * 32)), desc_at(db, (uint32_t)(j * (W3_BM * 128) + kk * 32)),
                       idesc1, (j | kk) != 0);
          }
          if (sl) umma_commit(w_empty);
          umma_commit(&acc_full[as]);
        }
        __syncwarp();
      };
      auto mma2 = [&](int j) {
        const int t = t0 + j;
        const int s = j % W3_NST, ph = (j / W3_NST) & 1;
        mbar_wait(&o_full[s], ph);
        const bool first = seg_first(t);
        if (first) mbar_wait(d2_empty, (segD & 1) ^ 1);  // the previous segment's accumulator has been flushed
        tc_fence_after();
        if (lane == 0) w3_stamp(p.trace, j, 3);
        const DescHalves dd = desc_halves(smem_u32(dbuf + s * W3_DBYTES), 8192, 1024), da = desc_halves(smem_u32(abuf + s * W3_ABYTES), 8192, 1024);
        const bool sl = seg_last(t);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < W3_BM / 16; ++kk)  // K = 64 cells; MN-major operands: 2048 bytes per 16 k-rows
            umma_f16(d2_tmem, desc_at(dd, (uint32_t)(kk * 2048)), desc_at(da, (uint32_t)(kk * 2048)), idesc2, !first || kk != 0);
          umma_commit(&act_empty[s]);
          umma_commit(&d_empty[s]);
          if (sl) umma_commit(d2_full);
        }
        __syncwarp();
        if (sl) ++segD;
      };
      const int n_it = t1 - t0;
      if (n_it > 0) mma1(0);
      for (int it = 0; it < n_it; ++it) {
        if (it + 1 < n_it) mma1(it + 1);  // issued right behind MMA2(it - 1): overlaps the epilogue of tile it
        mma2(it);
      }
    }
  } else if (warp == N_EPI_WARPS + 2) {
    // ===== TMA store warp: finished ds tiles -> d logits (scale half), read by the dAct GEMM =====
    if (lane == 0) {
      for (int t = t0, it = 0; t < t1; ++t, ++it) {
        const int slab = t / n_ct, ct = t % n_ct;
        const int g0 = slab * W3_SLAB, b0 = ct * W3_BM;
        const int s = it % W3_NST, ph = (it / W3_NST) & 1;
        mbar_wait(&o_full[s], ph);
        w3_stamp(p.trace, it, 4);
        for (int j = 0; j < 2; ++j) tma_store_2d(&tmD, dbuf + s * W3_DBYTES + j * (W3_BM * 128), g0 + 64 * j, b0);
        tma_store_commit();
        tma_store_wait_read0();
        w3_stamp(p.trace, it, 5);
        mbar_arrive(&d_empty[s]);
      }
      tma_store_wait_all();
    }
  } else if (warp < N_EPI_WARPS) {
    // ===== epilogue warps: TMEM lane = gene r of the slab, column group cg = 16 cells of the tile =====
    const int q = warp % 4, cg = warp / 4;
    const int r = q * 32 + lane;
    const float unscale = p.c * (1.f / DL_SCALE);
    const float L2E = 1.4426950408889634f;
    // shared-space addresses, made opaque so that they stay in registers (the compiler otherwise re-derives them from
    // %tid and the shared window every tile).  Entry (cell c, gene r) of a ds tile: box r / 64, row c, 16-byte chunk
    // ((r % 64) / 8) ^ (c % 8), byte 2 (r % 8); this warp's cells are c = 16 cg + j, so c % 8 = j % 8.
    uint32_t d_row0 = smem_u32(dbuf) + (uint32_t)((r >> 6) * (W3_BM * 128) + (r & 7) * 2 + cg * 16 * 128);
    uint32_t c_row0 = smem_u32(cbuf) + (uint32_t)(cg * 64);
    uint32_t tm_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t sw[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) sw[k] = (uint32_t)(((((r & 63) >> 3) ^ k) << 4));
    asm volatile("" : "+r"(d_row0), "+r"(c_row0), "+r"(tm_lane));
#pragma unroll
    for (int k = 0; k < 8; ++k) asm volatile("" : "+r"(sw[k]));
    float bias_g = 0.f, dbacc = 0.f;
    int segE = 0;
    int slab = t0 / n_ct, ct = t0 % n_ct;  // walked incrementally (no division per tile)
    int s = 0, sph = 0;
    for (int t = t0, it = 0; t < t1; ++t, ++it) {
      const int gene = slab * W3_SLAB + r, b0 = ct * W3_BM + cg * 16;
      const bool gene_ok = gene < p.G;
      const bool first = (it == 0) || (ct == 0), last = (t == t1 - 1) || (ct == n_ct - 1);
      if (first) {
        // exponent constant of this gene: (bias + log(DL_SCALE)) log2(e) - the softmax term comes out pre-scaled
        bias_g = gene_ok ? __fmaf_rn(__ldg(&p.bias[gene]), L2E, 2.f) : 0.f;
        static_assert(DL_SCALE == 4.0f, "log2(DL_SCALE) = 2 is folded into the exponent");
        dbacc = 0.f;
      }
      const int as = it & 1, aph = (it >> 1) & 1;
      // the staged ds' tile and the per-cell terms first: their shared-memory latency hides behind the wait for MMA1
      mbar_wait(&x_full[s], sph);
      if (threadIdx.x == 0) w3_stamp(p.trace, it, 6);
      const uint32_t drow = d_row0 + (uint32_t)(s * W3_DBYTES);
      uint32_t base[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) base[k] = drow + sw[k];
      unsigned short raw[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        raw[j] = lds_u16_off<0>(base[j] + (uint32_t)(j * 128));
        raw[j + 8] = lds_u16_off<8 * 128>(base[j] + (uint32_t)(j * 128));
      }
      const uint32_t crow = c_row0 + (uint32_t)(s * W3_CBYTES);
      float nl[16], sa[16];
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const float4 a = lds128(crow + j4 * 16), b = lds128(crow + W3_BM * 4 + j4 * 16);
        nl[4 * j4] = a.x; nl[4 * j4 + 1] = a.y; nl[4 * j4 + 2] = a.z; nl[4 * j4 + 3] = a.w;
        sa[4 * j4] = b.x; sa[4 * j4 + 1] = b.y; sa[4 * j4 + 2] = b.z; sa[4 * j4 + 3] = b.w;
      }
      mbar_wait(&acc_full[as], aph);
      if (threadIdx.x == 0) w3_stamp(p.trace, it, 7);
      tc_fence_after();
      float sv[16];
      tmem_ld16(tm_lane + (uint32_t)(as * W3_BM + cg * 16), sv);
      tmem_ld_wait();
      if (threadIdx.x == 0) w3_stamp(p.trace, it, 8);
      // ds = ds' + DL_SCALE A[b] softmax(s)  (both stored pre-scaled by DL_SCALE / c)
      float ds[16];
      if (gene_ok && b0 + 16 <= p.B) {  // interior: no validity selects
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float pr = ex2_approx(__fmaf_rn(nl[j], -L2E, __fmaf_rn(sv[j], L2E, bias_g)));
          ds[j] = __fmaf_rn(sa[j], pr, __half2float(__ushort_as_half(raw[j])));
          dbacc += ds[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float pr = ex2_approx(__fmaf_rn(nl[j], -L2E, __fmaf_rn(sv[j], L2E, bias_g)));
          const bool ok = gene_ok && (b0 + j < p.B);
          ds[j] = ok ? __fmaf_rn(sa[j], pr, __half2float(__ushort_as_half(raw[j]))) : 0.f;
          dbacc += ds[j];
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sts_h_sat_off<0>(base[j] + (uint32_t)(j * 128), ds[j]);
        sts_h_sat_off<8 * 128>(base[j] + (uint32_t)(j * 128), ds[j + 8]);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&acc_empty[as]);
        mbar_arrive(&o_full[s]);
      }
      if (threadIdx.x == 0) w3_stamp(p.trace, it, 9);
      if (last) {
        if (gene_ok) atomicAdd(&p.db[gene], dbacc * unscale);
        mbar_wait(d2_full, segE & 1);
        tc_fence_after();
        // dW[gene][c .. c+15] += alpha D2[r][c ..]: this warp takes every fourth 16-column chunk
        for (int c = cg * 16; c < p.h; c += 64) {
          float v[16];
          tmem_ld16(tm_lane + (uint32_t)(W3_D2_COL + c), v);
          tmem_ld_wait();
          if (gene_ok) {
            float* dst = dW + (size_t)gene * p.h + c;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              red_add_v4(dst + j, v[j] * unscale, v[j + 1] * unscale, v[j + 2] * unscale, v[j + 3] * unscale);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d2_empty);
        ++segE;
      }
      if (++ct == n_ct) { ct = 0; ++slab; }
      if (++s == W3_NST) { s = 0; sph ^= 1; }
    }
    w3_stamp_cta(p.trace, 3);
  }
  tc_fence_before();
  __syncthreads();
  w3_stamp_cta(p.trace, 4);
  if (warp == N_EPI_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// act = relu(bn(U)) materialised once (B x h fp32, L2 resident) as the TMA-fed A operand
__global__ void k_act(const float* __restrict__ U, int B, int h, BnParams bn, float* __restrict__ act,
                      __half* __restrict__ act_h) {
  pdl_prologue();
  size_t n = (size_t)B * h;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int col = (int)(i % h);
    const float v = bn_relu(U[i], bn.mean[col], bn.rstd[col], bn.gamma[col], bn.beta[col]);
    if (act_h) act_h[i] = __float2half_rn(v);
    else act[i] = rna_tf32(v);
  }
}

// fp16 copy of the last-layer weight (the MMA operand of the logits GEMM and of dAct)
__global__ void k_to_half(const float* __restrict__ src, size_t n4, __half2* __restrict__ dst) {
  pdl_prologue();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    dst[2 * i] = __floats2half2_rn(v.x, v.y);
    dst[2 * i + 1] = __floats2half2_rn(v.z, v.w);
  }
}

struct Scratch {
  float* act;      // [B][h]  relu(bn(U)) rounded to tf32
  float* Wr;       // [2G][h] last-layer weight rounded to tf32
  float* part_m;   // [slots][B]
  float* part_s;
  __half* act_h;   // [B][h]  fp16 operand copies (H16 path)
  __half* W_h;     // [2G][h]
  size_t total;
};

Scratch carve(void* base, int B, int G, int h) {
  Scratch s{};
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = (o + bytes + 255) / 256 * 256; return r; };
  size_t slots = (size_t)((G + 127) / 128) * 4;
  size_t a = take(4 * (size_t)B * h), pm = take(4 * slots * B), ps = take(4 * slots * B);
  size_t wr = take(4 * (size_t)2 * G * h);
  size_t ah = take(2 * (size_t)B * h), wh = take(2 * (size_t)2 * G * h);
  s.total = o;
  if (base) {
    char* c = (char*)base;
    s.act = (float*)(c + a);
    s.part_m = (float*)(c + pm);
    s.part_s = (float*)(c + ps);
    s.Wr = (float*)(c + wr);
    s.act_h = (__half*)(c + ah);
    s.W_h = (__half*)(c + wh);
  }
  return s;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// fp16 operands need 16-byte-aligned halves of d logits (G % 8) and a full 64-element k-block
bool use_h16(const TcDecoderArgs& a) {
  static const bool off = getenv("DRV_DEC_TF32") != nullptr;  // debug / A-B: the tf32-operand decoder
  return !off && a.G % 8 == 0 && a.h >= 64;
}

bool make_maps(const TcDecoderArgs& a, const Scratch& s, CUtensorMap* act, CUtensorMap* w1, CUtensorMap* w3, CUtensorMap* x) {
  if (use_h16(a)) {
    if (!tmap_2d_h(act, s.act_h, a.B, a.h, a.h, BM, 64)) return false;
    if (!tmap_2d_h(w1, s.W_h, a.G, a.h, a.h, 128, 64)) return false;
    uint64_t dims[3] = {(uint64_t)a.h, (uint64_t)a.G, 2};
    uint64_t strides[2] = {(uint64_t)a.h * 2, (uint64_t)a.G * a.h * 2};
    uint32_t box[3] = {64, 64, 2};
    if (!make_tensor_map(w3, s.W_h, 3, dims, strides, box, 1, true)) return false;
    return tmap_2d(x, a.x, a.B, a.G, a.G, BM, 32);
  }
  if (!tmap_2d(act, s.act, a.B, a.h, a.h, BM, BK)) return false;
  if (!tmap_2d(w1, a.W_hi, a.G, a.h, a.h, 128, BK)) return false;
  {
    uint64_t dims[3] = {(uint64_t)a.h, (uint64_t)a.G, 2};
    uint64_t strides[2] = {(uint64_t)a.h * 4, (uint64_t)a.G * a.h * 4};
    uint32_t box[3] = {BK, 64, 2};
    if (!make_tensor_map(w3, a.W_hi, 3, dims, strides, box, 1)) return false;
  }
  if (!tmap_2d(x, a.x, a.B, a.G, a.G, BM, 32)) return false;
  return true;
}

// d / r: tensor maps of the scale / log_r halves of dlogits for the TMA stores (any valid map where unused)
template <class K>
void launch_dec(K kern, int n_tiles, const CUtensorMap& act, const CUtensorMap& w, const CUtensorMap& x,
                const CUtensorMap& d, const CUtensorMap& r, const DecParams& p, cudaStream_t st,
                int smem_bytes = SMEM_SMALL) {
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  launch_k_plain(kern, grid, NTHREADS, smem_bytes, st, act, w, x, d, r, p);
  note_launch();
}

}  // namespace

size_t tc_workspace_bytes(const drv_dims& d) {
  int h = d.layers[0];  // the decoder's last hidden width is the encoder's first
  return carve(nullptr, d.n_cells, d.n_genes, h).total;
}

bool tc_decoder_supported(const drv_dims& d, const drv_hparams& hp) {
  if (hp.flags & DRV_FLAG_FORCE_SIMT) return false;
  int h = d.layers[0];
  // TMA needs 16-byte row strides; the k loop runs in 32-float blocks; the MN-major GEMMs want h % 32 == 0
  return d.n_genes % 4 == 0 && d.n_genes >= 64 && h % 32 == 0 && h >= 32 && h <= 512 && d.n_cells >= 32;
}

static void set_lse_layout(DecParams& p, int n_ct, int n_gt1) {
  p.lse_n_gt = n_gt1;
  p.lse_n_tiles = n_ct * n_gt1;
  p.lse_grid = p.lse_n_tiles < sm_count() ? p.lse_n_tiles : sm_count();
}

bool tc_decoder_prepare(const TcDecoderArgs& a, cudaStream_t st) {
  if (!use_h16(a)) return false;
  Scratch s = carve(a.scratch, a.B, a.G, a.h);
  const size_t n4 = (size_t)2 * a.G * a.h / 4;
  int wb = (int)((n4 + 255) / 256);
  if (wb > sm_count() * 8) wb = sm_count() * 8;
  launch_k(k_to_half, wb, 256, 0, st, a.W, n4, reinterpret_cast<__half2*>(s.W_h));
  note_launch();
  return true;
}

void tc_decoder_forward(const TcDecoderArgs& a, cudaStream_t st) {
  Scratch s = carve(a.scratch, a.B, a.G, a.h);
  CUtensorMap tAct, tW1, tW3, tX;
  if (!make_maps(a, s, &tAct, &tW1, &tW3, &tX)) {
    if (g_err == cudaSuccess) g_err = cudaErrorInvalidValue;
    return;
  }
  const bool h16 = use_h16(a);
  int blocks = (int)(((size_t)a.B * a.h + 255) / 256);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  launch_k(k_act, blocks, 256, 0, st, a.U, a.B, a.h, a.bn, s.act, h16 ? s.act_h : (__half*)nullptr);
  note_launch();
  if (h16 && !a.w_prepared) tc_decoder_prepare(a, st);
  DecParams p{};
  p.B = a.B; p.G = a.G; p.h = a.h; p.bias = a.bias; p.lib = a.lib; p.rshift = a.rshift; p.eps = a.eps; p.c = a.c;
  p.part_m = s.part_m; p.part_s = s.part_s; p.lse = a.lse; p.ll = a.ll; p.asum = a.asum;
  p.dlogits = a.dlogits; p.db = a.db;
  const int n_ct = (a.B + BM - 1) / BM;
  // phase 1: logsumexp over genes
  const int n_gt1 = (a.G + 127) / 128;
  set_lse_layout(p, n_ct, n_gt1);
  drv_internal_kmark(0, 0, st);
  if (h16) launch_dec(k_dec<1, false, true>, n_ct * n_gt1, tAct, tW1, tX, tX, tX, p, st);
  else launch_dec(k_dec<1, false, false>, n_ct * n_gt1, tAct, tW1, tX, tX, tX, p, st);
  drv_internal_kmark(0, 1, st);
  // phase 2: NB log-likelihood row sums (ll / asum / db were zeroed at the start of the step)
  const int n_gt = (a.G + 63) / 64;
  drv_internal_kmark(1, 0, st);
  if (a.bwd) {
    // scale / log_r halves of dlogits: [B][G] each, row stride 2G (fp16 on the H16 path, same buffer)
    CUtensorMap tD, tR;
    bool ok;
    if (h16) {
      const __half* dl = reinterpret_cast<const __half*>(a.dlogits);
      ok = tmap_2d_h(&tD, dl, a.B, a.G, 2 * (uint64_t)a.G, BM, 64) && tmap_2d_h(&tR, dl + a.G, a.B, a.G, 2 * (uint64_t)a.G, BM, 64);
    } else {
      ok = tmap_2d(&tD, a.dlogits, a.B, a.G, 2 * (uint64_t)a.G, BM, 32) &&
           tmap_2d(&tR, a.dlogits + a.G, a.B, a.G, 2 * (uint64_t)a.G, BM, 32);
    }
    if (!ok) {
      if (g_err == cudaSuccess) g_err = cudaErrorInvalidValue;
      return;
    }
    if (h16) launch_dec(k_dec<2, true, true>, n_ct * n_gt, tAct, tW3, tX, tD, tR, p, st, SMEM_BYTES);
    else launch_dec(k_dec<2, true, false>, n_ct * n_gt, tAct, tW3, tX, tD, tR, p, st, SMEM_BYTES);
  } else {
    if (h16) launch_dec(k_dec<2, false, true>, n_ct * n_gt, tAct, tW3, tX, tX, tX, p, st);
    else launch_dec(k_dec<2, false, false>, n_ct * n_gt, tAct, tW3, tX, tX, tX, p, st);
  }
  drv_internal_kmark(1, 1, st);
  // the row sums LL[b] are folded into the loss by the step's last launch (finalize_step)
}

// The fused phase 3 + dW kernel (k_dec3w) needs the fp16 operands and an h x 128-gene fp32 accumulator in TMEM.
// DRV_DEC_FUSED_DW=0 selects phase 3 + the separate dW GEMM (A/B comparisons; also the path of h > 256).
bool use_fused_dw(const TcDecoderArgs& a) {
  const char* e = getenv("DRV_DEC_FUSED_DW");  // read per call: the parity tests toggle it
  const bool off = e != nullptr && atoi(e) == 0;
  return !off && use_h16(a) && a.h <= 256;
}

void tc_decoder_backward(const TcDecoderArgs& a, cudaStream_t st) {
  Scratch s = carve(a.scratch, a.B, a.G, a.h);
  CUtensorMap tAct, tW1, tW3, tX;
  if (!make_maps(a, s, &tAct, &tW1, &tW3, &tX)) {
    if (g_err == cudaSuccess) g_err = cudaErrorInvalidValue;
    return;
  }
  DecParams p{};
  p.B = a.B; p.G = a.G; p.h = a.h; p.bias = a.bias; p.lib = a.lib; p.eps = a.eps; p.c = a.c;
  p.part_m = s.part_m; p.part_s = s.part_s; p.lse = a.lse; p.ll = a.ll; p.asum = a.asum; p.dlogits = a.dlogits; p.db = a.db;
  const int n_ct = (a.B + BM - 1) / BM, n_gt1 = (a.G + 127) / 128;
  set_lse_layout(p, n_ct, n_gt1);
  const bool h16 = use_h16(a);
  const bool fused = use_fused_dw(a);
  const float alpha = a.c * (1.f / DL_SCALE);  // d logits is stored as DL_SCALE / c times its value
  const __half* dl_h = reinterpret_cast<const __half*>(a.dlogits);
  cudaStream_t ws = st;
  int rc = 0;
  if (fused) {
    // scale half: softmax coupling + weight gradient in one kernel, gene-major walk
    CUtensorMap tA64, tD64;
    if (!tmap_2d_h(&tA64, s.act_h, a.B, a.h, a.h, W3_BM, 64) || !tmap_2d_h(&tD64, dl_h, a.B, a.G, 2 * (uint64_t)a.G, W3_BM, 64)) {
      if (g_err == cudaSuccess) g_err = cudaErrorInvalidValue;
      return;
    }
    const int n_tiles = ((a.G + W3_SLAB - 1) / W3_SLAB) * ((a.B + W3_BM - 1) / W3_BM);
    if (getenv("DRV_W3_TRACE")) {
      void* sym = nullptr;
      if (cudaGetSymbolAddress(&sym, g_w3_trace) == cudaSuccess) p.trace = (const unsigned long long*)sym;
    }
    drv_internal_kmark(2, 0, st);
    cudaFuncSetAttribute(k_dec3w, cudaFuncAttributeMaxDynamicSharedMemorySize, W3_SMEM);
    launch_k_plain(k_dec3w, n_tiles < sm_count() ? n_tiles : sm_count(), NTHREADS, W3_SMEM, st, tA64, tW1, tD64, p, a.dW);
    note_launch();
    drv_internal_kmark(2, 1, st);
  } else {
    CUtensorMap tD;  // scale half of dlogits: [B][G] with row stride 2G
    if (!(h16 ? tmap_2d_h(&tD, a.dlogits, a.B, a.G, 2 * (uint64_t)a.G, BM, 64)
              : tmap_2d(&tD, a.dlogits, a.B, a.G, 2 * (uint64_t)a.G, BM, 32))) {
      if (g_err == cudaSuccess) g_err = cudaErrorInvalidValue;
      return;
    }
    drv_internal_kmark(2, 0, st);
    if (h16) launch_dec(k_dec<3, true, true>, n_ct * n_gt1, tAct, tW1, tD, tD, tD, p, st);
    else launch_dec(k_dec<3, true, false>, n_ct * n_gt1, tAct, tW1, tD, tD, tD, p, st);
    drv_internal_kmark(2, 1, st);
    if (a.side) a.side->mark(st);
  }
  // dAct[b][i] = sum_j dlogits[b][j] W[j][i]   (A K-major, B MN-major, K = 2G): critical path, submitted first
  drv_internal_kmark(3, 0, st);
  if (rc == 0)
    rc = h16 ? gemm_f16(a.dlogits, 2 * a.G, false, s.W_h, a.h, true, a.dY, a.h, a.B, a.h, 2 * a.G, 0, true, alpha, st)
             : gemm_tf32(a.dlogits, 2 * a.G, false, a.W_hi, a.h, true, a.dY, a.h, a.B, a.h, 2 * a.G, 0, /*accumulate (cleared at step start)=*/true, st);
  drv_internal_kmark(3, 1, st);
  if (fused) {
    // the log_r half of d logits is final since phase 2: dW[G + g][i] = sum_b drho[b][g] act[b][i] (both operands
    // MN-major, K = cells) is the one weight-gradient GEMM left.  It is full-machine work off the critical chain: forked
    // BEHIND the dAct GEMM, it runs next to the small grids of the decoder-trunk backward (launched next to k_dec3w it
    // took 120 SMs for 20 us and stretched that kernel from 34 to 62 us - profiles/r02_graph_timeline_*.txt)
    if (a.side) { a.side->mark(st); a.side->attach(); ws = a.side->side; }
    drv_internal_kmark(4, 0, ws);
    if (rc == 0)
      // ... on a part of the SMs only when it shares the machine: the trunk's own tensor-core GEMMs (32-CTA grids with a
      // full SM's shared memory each) need free SMs next to it
      rc = gemm_f16(dl_h + a.G, 2 * a.G, true, s.act_h, a.h, true, a.dW + (size_t)a.G * a.h, a.h, a.G, a.h, a.B,
                    a.side ? sm_count() - 56 : 0, true, alpha, ws);
    drv_internal_kmark(4, 1, ws);
  } else {
    // dW[j][i] = sum_b dlogits[b][j] act[b][i]   (both operands MN-major, K = cells): off the critical chain
    if (a.side) { a.side->attach(); ws = a.side->side; }
    drv_internal_kmark(4, 0, ws);
    if (rc == 0)
      rc = h16 ? gemm_f16(a.dlogits, 2 * a.G, true, s.act_h, a.h, true, a.dW, a.h, 2 * a.G, a.h, a.B, 0, true, alpha, ws)
               : gemm_tf32(a.dlogits, 2 * a.G, true, s.act, a.h, true, a.dW, a.h, 2 * a.G, a.h, a.B, 0, true, ws);
    drv_internal_kmark(4, 1, ws);
  }
  if (rc != 0 && g_err == cudaSuccess) g_err = cudaErrorInvalidValue;
  // the ReLU gate of dAct is applied by the BatchNorm backward kernels (block_backward)
}

}  // namespace drv

// debug: timeline stamps of k_dec3w's CTA 0 (DRV_W3_TRACE=1); returns the number of values copied
extern "C" int drv_debug_w3_trace(unsigned long long* host, int n) {
  const int m = drv::W3_TRACE_TILES * drv::W3_TRACE_EVENTS;
  if (n > m) n = m;
  if (cudaMemcpyFromSymbol(host, drv::g_w3_trace, sizeof(unsigned long long) * (size_t)n) != cudaSuccess) return -1;
  return n;
}
