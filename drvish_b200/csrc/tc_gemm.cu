// Generic tensor-core GEMM for sm_100a:  C[M][N] (+)= alpha * sum_k A(m,k) * B(n,k)
// tcgen05.mma.kind::tf32 (fp32 operands straight from memory) or kind::f16 (fp16 operands, template flag
// H16), fp32 accumulate in TMEM, operands staged by TMA with the 128-byte swizzle through an mbarrier ring,
// warp-specialised: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
// Each operand may be K-major ([rows][K], K contiguous) or MN-major ([K][rows], rows contiguous),
// so that d logits / activations / weights are consumed in the layout they already have:
//   dW   = dlogits^T act   : A = dlogits [B][2G] MN-major, B = act [B][h] MN-major, K = cells
//   dAct = dlogits  W      : A = dlogits [B][2G] K-major,  B = W [2G][h]  MN-major, K = 2G
// Split-K partial tiles are combined with red.global.add.f32 into a zeroed C.
#include <stdlib.h>
#include "tc.cuh"
#include "tc_common.cuh"

#include <mutex>

namespace drv {
namespace tc {

bool make_tensor_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle, bool f16) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeFn)p;
  });
  if (!fn) return false;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base),
                  (const cuuint64_t*)dims, (const cuuint64_t*)strides_bytes, (const cuuint32_t*)box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// debug overrides for the MN-major descriptor fields (bytes); 0 = defaults
static int g_mn_lbo = 4096, g_mn_sbo = 512, g_mn_kstep = 1024;
static int g_force_mt = getenv("DRV_MT") ? atoi(getenv("DRV_MT")) : 0;  // debug: 1 / 2 forces the M sub-tile count

namespace {

constexpr int BM = 128;       // rows of A per tile = TMEM lanes
constexpr int BK = 32;        // fp32 elements per k-block = one 128-byte swizzle row
constexpr int STAGES = 4;
constexpr int MAXN = 256;
constexpr int A_BYTES = BM * BK * 4;
constexpr int B_BYTES = MAXN * BK * 4;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int NTHREADS = 192;

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Persistent stream-K kernel: the (tile, k-block) units are dealt out evenly to the CTAs (one per
// SM); a CTA walks its contiguous unit range as <= a few (tile, k-range) segments.  A segment that
// covers a whole tile stores it, partial segments are combined with red.global.add into a zeroed C.
// Two TMEM accumulator stages overlap the epilogue of a segment with the MMAs of the next one.
// MT = number of 128-row M sub-tiles a CTA computes against ONE staged B tile.  MT = 2 (256 x 256 output
// tile, two TMEM accumulators, 3 stages of 64 KB) moves 1.5x fewer operand bytes from L2 per MMA
// than MT = 1 (4 stages of 48 KB, accumulator double-buffered across chunks): the large GEMMs are
// bound by L2 -> shared-memory bandwidth (ncu: tensor pipe 25-44 %, ~7 TB/s of TMA traffic).
// CL = 2: thread-block clusters of two CTAs take two M tiles of the same (N tile, k-part); each CTA
// loads HALF of every staged B tile and multicasts it into both CTAs' rings, so B is read from L2
// once per pair.  The large backward GEMMs are bound by exactly that traffic (dAct: 39 k-blocks x
// 64 KB per CTA x 128 CTAs = 320 MB through L2 in ~45 us; half of it is the shared B operand).
// H16: fp16 operands (kind::f16, same 10-bit mantissa as tf32): a k-block is 64 elements (still one
// 128-byte swizzle row), an MN-major box 64 elements wide, UMMA_K = 16 - half the bytes and half the
// pipeline hand-shakes per unit of K, twice the MMA rate.  alpha scales the result (exact power-of-two
// rescaling of fp16 operands that were stored pre-scaled).
template <bool A_MN, bool B_MN, int MT, int CL = 1, bool H16 = false>
__global__ void __launch_bounds__(NTHREADS, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
          const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
          const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, int n_pairs,
          float* __restrict__ C, int ldc, int M, int N, int K, int n_tile, int splits, int force_atomic,
          const float* __restrict__ bias, long long part_stride, int mn_lbo, int mn_sbo, int mn_kstep, float alpha,
          const float* __restrict__ alpha_dev) {
  pdl_prologue();
  constexpr int BKE = H16 ? 64 : 32;                 // elements per k-block
  constexpr int MNB = H16 ? 64 : 32;                 // M/N elements per MN-major box (128 bytes)
  constexpr int MNBB = BKE * 128;                    // bytes of one MN-major box
  constexpr uint32_t MN_LAYOUT = H16 ? 2 : 1;        // SWIZZLE_128B / SWIZZLE_128B with 32-byte atoms
  constexpr int TBM = MT * BM;                       // rows of the CTA tile
  constexpr int NSTG = MT == 1 ? STAGES : 3;
  constexpr int SA_BYTES = MT * A_BYTES;
  constexpr int STG_BYTES = SA_BYTES + B_BYTES;
  constexpr int NACC = MT == 1 ? 2 : 1;              // accumulator stages (MT = 2 uses all 512 TMEM columns)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + NSTG * STG_BYTES);
  uint64_t* empty = full + NSTG;
  uint64_t* acc_full = empty + NSTG;     // [2]
  uint64_t* acc_empty = acc_full + 2;    // [2]
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  // the operand pairs are concatenated along a virtual K axis: sum_p A_p B_p^T accumulates in TMEM
  const int kb_pair = (K + BKE - 1) / BKE;
  const int kb_total = kb_pair * n_pairs;
  const int m_tiles = (M + TBM - 1) / TBM, n_tiles = (N + n_tile - 1) / n_tile;
  // work = (group of CL M tiles, N tile, k-part) chunks: every tile's k range is cut into `splits`
  // parts; cluster w owns the contiguous chunk range [c0, c1) and its CTA `rank` the M tile
  // CL * group + rank of every chunk
  const int rank = CL == 1 ? 0 : (int)cluster_ctarank();
  const int n_wk = (int)gridDim.x / CL, wk = (int)blockIdx.x / CL;
  const long long chunks = (long long)((m_tiles + CL - 1) / CL) * n_tiles * splits;
  const int c0 = (int)(chunks * wk / n_wk);
  const int c1 = (int)(chunks * (wk + 1) / n_wk);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < NSTG; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CL);  // a stage is refilled (by both CTAs of a pair) once every CTA has consumed it
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<2 * MAXN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const uint32_t tx = SA_BYTES + (uint32_t)n_tile * 128;
      int i = 0;
      for (int ch = c0; ch < c1; ++ch) {
        const int tile = ch / splits, part = ch % splits;
        const int kb_b = (int)((long long)kb_total * part / splits), kb_e = (int)((long long)kb_total * (part + 1) / splits);
        const int m0 = (CL * (tile / n_tiles) + rank) * TBM, n0 = (tile % n_tiles) * n_tile;
        for (int kb = kb_b; kb < kb_e; ++kb, ++i) {
          const int s = i % NSTG, ph = (i / NSTG) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = smem + s * STG_BYTES;
          uint8_t* sb = sa + SA_BYTES;
          mbar_expect_tx(&full[s], tx);
          const int pair = kb / kb_pair;
          const int k0 = (kb - pair * kb_pair) * BKE;
          const CUtensorMap* ma = pair == 0 ? &tmA : (pair == 1 ? &tmA1 : &tmA2);
          const CUtensorMap* mb = pair == 0 ? &tmB : (pair == 1 ? &tmB1 : &tmB2);
          if (A_MN) {
            for (int j = 0; j < TBM / MNB; ++j) tma_load_2d(sa + j * MNBB, ma, &full[s], m0 + MNB * j, k0);
          } else {
            for (int j = 0; j < MT; ++j) tma_load_2d(sa + j * A_BYTES, ma, &full[s], k0, m0 + j * BM);
          }
          if (CL == 2) {
            // my half of the B tile, delivered to both CTAs of the pair (tmB's box is half a tile here)
            if (B_MN) {
              const int nb = n_tile / (2 * MNB);
              for (int j = rank * nb; j < (rank + 1) * nb; ++j)
                tma_load_2d_mc(sb + j * MNBB, mb, &full[s], n0 + MNB * j, k0, (uint16_t)0x3);
            } else {
              const int half = n_tile / 2;
              tma_load_2d_mc(sb + rank * half * 128, mb, &full[s], k0, n0 + rank * half, (uint16_t)0x3);
            }
          } else if (B_MN) {
            for (int j = 0; j < n_tile / MNB; ++j) tma_load_2d(sb + j * MNBB, mb, &full[s], n0 + MNB * j, k0);
          } else {
            tma_load_2d(sb, mb, &full[s], k0, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      const uint32_t idesc = H16 ? idesc_f16(BM, n_tile, A_MN ? 1 : 0, B_MN ? 1 : 0) : idesc_tf32(BM, n_tile, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int i = 0, seg = 0;
      for (int ch = c0; ch < c1; ++ch, ++seg) {
        const int part = ch % splits;
        const int kb_b = (int)((long long)kb_total * part / splits), kb_e = (int)((long long)kb_total * (part + 1) / splits);
        const int as = seg % NACC, aph = (seg / NACC) & 1;
        mbar_wait(&acc_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * MAXN);
        for (int kb = kb_b; kb < kb_e; ++kb, ++i) {
          const int s = i % NSTG, ph = (i / NSTG) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STG_BYTES);
          const uint32_t sb = sa + SA_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t db = B_MN ? smem_desc(sb + kk * mn_kstep, mn_lbo, mn_sbo, MN_LAYOUT) : smem_desc(sb + kk * 32, 16, 1024);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const uint32_t sam = sa + mt * A_BYTES;  // sub-tile mt: its own 128 rows (4 MN-major boxes)
              const uint64_t da = A_MN ? smem_desc(sam + kk * mn_kstep, mn_lbo, mn_sbo, MN_LAYOUT) : smem_desc(sam + kk * 32, 16, 1024);
              if (H16) umma_f16(d_tmem + (uint32_t)(mt * MAXN), da, db, idesc, (kb != kb_b) || (kk != 0));
              else umma_tf32(d_tmem + (uint32_t)(mt * MAXN), da, db, idesc, (kb != kb_b) || (kk != 0));
            }
          }
          if (CL == 2) umma_commit_mc(&empty[s], (uint16_t)0x3);  // ... in both CTAs: the peer multicasts into this stage too
          else umma_commit(&empty[s]);  // frees the smem stage once these MMAs have read it
        }
        umma_commit(&acc_full[as]);
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global =====
    const int q = warp % 4;  // TMEM lane quarter this warp may access
    int seg = 0;
    if (H16 && alpha_dev) alpha *= __ldcg(alpha_dev);  // written by an earlier kernel of the stream
    for (int ch = c0; ch < c1; ++ch, ++seg) {
      const int tile = ch / splits, part = ch % splits;
      const int m0 = (CL * (tile / n_tiles) + rank) * TBM, n0 = (tile % n_tiles) * n_tile;
      const bool whole = !force_atomic;
      const bool add_bias = bias != nullptr && part == 0 && part_stride == 0;
      const int as = seg % NACC, aph = (seg / NACC) & 1;
      mbar_wait(&acc_full[as], aph);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
      const int m = m0 + mt * BM + q * 32 + lane;
      const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * MAXN + mt * MAXN);
      for (int c = 0; c < n_tile; c += 16) {
        float v[16];
        tmem_ld16(tb + (uint32_t)c, v);
        tmem_ld_wait();
        if (H16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= alpha;
        }
        if (m < M) {
          // part_stride != 0: deterministic split - every k-part stores into its own slab
          float* dst = C + (size_t)part * part_stride + (size_t)m * ldc + n0 + c;
          if (add_bias) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n0 + c + j < N) v[j] += __ldg(&bias[n0 + c + j]);
          }
          const bool vec_ok = (n0 + c + 16 <= N) && ((((uintptr_t)dst) & 15) == 0);
          if (vec_ok) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              if (whole) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              else red_add_v4(dst + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (n0 + c + j < N) {
                if (whole) dst[j] = v[j];
                else atomicAdd(dst + j, v[j]);
              }
            }
          }
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2 * MAXN>(tmem_base);
  }
}

}  // namespace

// C[m][n] = bias[n] + sum_p partials[p][m][n], parts added in index order (deterministic)
__global__ void k_reduce_parts(const float* __restrict__ parts, int n_parts, size_t mn, int N, const float* __restrict__ bias,
                               float* __restrict__ C, int ldc) {
  pdl_prologue();
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < mn; i += (size_t)gridDim.x * blockDim.x * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const bool vec = (N % 4 == 0) && (i + 4 <= mn);
    for (int p = 0; p < n_parts; ++p) {
      const float* src = parts + (size_t)p * mn + i;
      if (vec) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(src));
        acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
      } else {
        for (int j = 0; j < 4; ++j)
          if (i + j < mn) acc[j] += src[j];
      }
    }
    for (int j = 0; j < 4; ++j) {
      const size_t e = i + j;
      if (e >= mn) break;
      const size_t m = e / N, n = e % N;
      C[m * ldc + n] = acc[j] + (bias ? bias[n] : 0.f);
    }
  }
}

// Same sums (same part order, so C is bit-identical to k_reduce_parts) + BatchNorm statistics of C:
// block (64 column quads, 4 row lanes), grid (1, row splits); each thread folds <= 32 rows in fp32,
// the CTA folds its row lanes in fp64 and issues one double atomic per column and moment; the last
// CTA finalises mean / rstd / running buffers (stat_finalize_if_last).  Requires N % 4 == 0, ldc % 4 == 0.
__global__ void __launch_bounds__(256) k_reduce_parts_stats(const float* __restrict__ parts, int n_parts, int M, int N,
                                                            const float* __restrict__ bias, float* __restrict__ C, int ldc,
                                                            int rows_per_cta, double* __restrict__ acc, StatFinal fin) {
  pdl_prologue();
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  const size_t mn = (size_t)M * N;
  __shared__ float sh[2][4][256];
  for (int cq0 = 0; cq0 < N / 4; cq0 += 64) {
    const int cq = cq0 + tx;
    float ps[4] = {0.f, 0.f, 0.f, 0.f}, pss[4] = {0.f, 0.f, 0.f, 0.f};
    if (cq < N / 4) {
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias) bv = *reinterpret_cast<const float4*>(bias + cq * 4);
      for (int m = r0 + ty; m < r1; m += 4) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        const float* src = parts + (size_t)m * N + cq * 4;
        // the slabs come from DRAM: keep 8 independent 16-byte loads in flight per thread (the kernel
        // is latency-bound otherwise); the sums still run in part order
        for (int p0 = 0; p0 < n_parts; p0 += 8) {
          float4 v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            v[j] = (p0 + j < n_parts) ? __ldcs(reinterpret_cast<const float4*>(src + (size_t)(p0 + j) * mn))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (p0 + j < n_parts) { a[0] += v[j].x; a[1] += v[j].y; a[2] += v[j].z; a[3] += v[j].w; }
          }
        }
        a[0] += bv.x; a[1] += bv.y; a[2] += bv.z; a[3] += bv.w;
        *reinterpret_cast<float4*>(C + (size_t)m * ldc + cq * 4) = make_float4(a[0], a[1], a[2], a[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ps[j] += a[j];
          pss[j] = __fmaf_rn(a[j], a[j], pss[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sh[0][ty][tx * 4 + j] = ps[j];
      sh[1][ty][tx * 4 + j] = pss[j];
    }
    __syncthreads();
    const int tid = ty * 64 + tx;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const int col = cq0 * 4 + tid;
      if (col < N) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) t += (double)sh[which][i][tid];
        atomicAdd(&acc[(size_t)which * N + col], t);
      }
    }
    __syncthreads();
  }
  stat_finalize_if_last(fin, acc, N, 0, N);
}

static int n_sm_tc() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// A: K-major [M][K] (lda = row stride in elements) or MN-major [K][M]; same for B.
// Requirements: row strides multiples of 4 floats, bases 16-byte aligned, N tile multiple of 16
// (32 for an MN-major B).  n_pairs operand pairs (same shapes/strides) are summed: C = sum_p A_p B_p^T.
// accumulate: add into the existing contents of C; otherwise C is overwritten (it is zeroed here
// first when stream-K splits a tile over several CTAs).  max_ctas <= 0: one CTA per SM.
static int gemm_impl(bool h16, float alpha, const float* alpha_dev, int n_pairs, const void* const* As, const void* const* Bs, int lda, bool a_mn,
                     int ldb, bool b_mn, float* C, int ldc, int M, int N, int K, int max_ctas, bool accumulate,
                     const float* bias, cudaStream_t st, int max_chain_kb, float* partials, int max_parts,
                     const ColStatsOut* stats) {
  if (n_pairs < 1 || n_pairs > 3) return -3;
  CUtensorMap ta[3], tb[3];
  const int bke = h16 ? 64 : 32;   // elements per k-block
  const int mnb = h16 ? 64 : 32;   // an MN-major operand is staged in boxes of 128 bytes of M/N elements
  const int nq = b_mn ? mnb : 16;
  int n_tile = N >= MAXN ? MAXN : ((N + nq - 1) / nq) * nq;
  for (int p = 0; p < 3; ++p) {
    int q = p < n_pairs ? p : 0;
    bool ok;
    if (h16) {
      if (a_mn) ok = tmap_2d_h(&ta[p], As[q], (uint64_t)K, (uint64_t)M, (uint64_t)lda, bke, mnb);
      else ok = tmap_2d_h(&ta[p], As[q], (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, bke);
      if (!ok) return -3;
      if (b_mn) ok = tmap_2d_h(&tb[p], Bs[q], (uint64_t)K, (uint64_t)N, (uint64_t)ldb, bke, mnb);
      else ok = tmap_2d_h(&tb[p], Bs[q], (uint64_t)N, (uint64_t)K, (uint64_t)ldb, (uint32_t)n_tile, bke);
      if (!ok) return -3;
      continue;
    }
    if (a_mn) ok = tmap_2d(&ta[p], (const float*)As[q], (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, 32, 2);
    else ok = tmap_2d(&ta[p], (const float*)As[q], (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK);
    if (!ok) return -3;
    if (b_mn) ok = tmap_2d(&tb[p], (const float*)Bs[q], (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, 32, 2);
    else ok = tmap_2d(&tb[p], (const float*)Bs[q], (uint64_t)N, (uint64_t)K, (uint64_t)ldb, (uint32_t)n_tile, BK);
    if (!ok) return -3;
  }
  const int kb_total = ((K + bke - 1) / bke) * n_pairs;
  // MN-major descriptor fields (bytes): stride between M/N boxes, between 8-row k groups, per UMMA_K
  const int mn_lbo = h16 ? 8192 : g_mn_lbo, mn_sbo = h16 ? 1024 : g_mn_sbo, mn_kstep = h16 ? 2048 : g_mn_kstep;
  const int sms = max_ctas > 0 && max_ctas < n_sm_tc() ? max_ctas : n_sm_tc();
  // Candidate schedules: MT = 1 (128-row tiles) or MT = 2 (256-row tiles, 1.5x fewer operand bytes
  // per MMA but no accumulator double-buffering, so only for long k ranges), each with the k range
  // of every tile split into `splits` parts so that the (tile, part) chunks fill the SMs.  Cost
  // model per schedule: waves x (k-blocks per chunk x relative k-block time + epilogue), where a
  // k-block costs MT x 1.0 MMA units but only (MT + 2) / 3 of the L2 traffic time of MT = 1; the
  // epilogue is ~6 k-blocks (store) or ~12 (red.add of split tiles) per 128 rows.
  int splits = 1, mt_sel = 1;
  {
    const int cand[] = {1, 2, 3, 4, 6, 8, 12, 16};
    double best = -1;
    for (int mt = 1; mt <= 2; ++mt) {
      if (mt == 2 && (M <= BM || g_force_mt == 1)) continue;
      if (mt == 1 && g_force_mt == 2 && M > BM) continue;
      const long long tiles = (long long)((M + mt * BM - 1) / (mt * BM)) * ((N + n_tile - 1) / n_tile);
      for (int ci = 0; ci < 8; ++ci) {
        const int sp = cand[ci];
        if (sp > 1 && kb_total / sp < 8) break;
        const long long waves = (tiles * sp + sms - 1) / sms;
        const double per = (double)((kb_total + sp - 1) / sp);
        // time per k-block relative to MT = 1: bound by max(MMA time, L2 traffic time); with the
        // measured ~40 % tensor utilisation of MT = 1 the traffic term dominates
        const double kb_cost = mt == 1 ? 1.0 : 1.35;
        const double epi = (sp > 1 ? 12.0 : 6.0) * mt * (mt == 2 ? 1.5 : 1.0);
        const double cost = (double)waves * (per * kb_cost + epi);
        if (best < 0 || cost < best) { best = cost; splits = sp; mt_sel = mt; }
      }
    }
  }
  // Precision knob: the TMEM accumulator is updated with truncation, a bias that grows with the
  // length of the accumulation chain (measured ~5e-9 relative per MMA step).  GEMMs that feed a
  // BatchNorm -> ReLU directly cap the chain; the partial tiles are combined by fp32 red.add (RN).
  if (max_chain_kb > 0 && (kb_total + splits - 1) / splits > max_chain_kb) splits = (kb_total + max_chain_kb - 1) / max_chain_kb;
  // Deterministic mode (forward GEMMs: their results decide ReLU gates, so they must not depend on
  // the order of atomics): every k-part is stored to its own slab of `partials` and the slabs are
  // summed in a fixed order by k_reduce_parts.
  const bool deterministic = partials != nullptr && !accumulate;
  if (deterministic && splits > max_parts) splits = max_parts;
  const long long tiles = (long long)((M + mt_sel * BM - 1) / (mt_sel * BM)) * ((N + n_tile - 1) / n_tile);
  const long long chunks = tiles * splits;
  int grid = (int)(chunks < sms ? chunks : sms);
  const bool split = splits > 1;
  float* outp = C;
  int out_ld = ldc, force_atomic = (split || accumulate) ? 1 : 0;
  long long part_stride = 0;
  const float* kbias = bias;
  if (deterministic && split) {
    outp = partials; out_ld = N; force_atomic = 0; part_stride = (long long)M * N; kbias = nullptr;
  } else if (split && !accumulate) {
    if (ldc == N) cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, st);
    else cudaMemset2DAsync(C, sizeof(float) * (size_t)ldc, 0, sizeof(float) * (size_t)N, (size_t)M, st);
  }
  auto launch = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    launch_k_plain(kern, grid, NTHREADS, SMEM_BYTES, st, ta[0], tb[0], ta[1], tb[1], ta[2], tb[2], n_pairs, outp, out_ld, M, N, K, n_tile,
                                             splits, force_atomic, kbias, part_stride, mn_lbo, mn_sbo, mn_kstep, alpha, alpha_dev);
  };
  // Cluster-of-2 schedule for the large accumulate-mode GEMMs (256-row tiles, one operand pair): the
  // B tile is multicast across two M tiles.
  // Measured (profiles/r01_notes.md): neutral at cluster size 2 - dW 55.7 -> 51.7 us, dAct 55.8 -> 57.7 us -
  // consistent with the L2's own de-duplication of near-simultaneous unicast requests; opt-in until a
  // larger cluster / 2-CTA MMA variant exists.  DRV_CLUSTER_GEMM=1 enables it (parity-tested).
  static const bool no_cluster = getenv("DRV_CLUSTER_GEMM") == nullptr;
  const int m_tiles2 = (M + 2 * BM - 1) / (2 * BM);
  if (!h16 && !no_cluster && mt_sel == 2 && n_pairs == 1 && !deterministic && n_tile % 64 == 0 && m_tiles2 >= 2 && sms >= 2) {
    if (!b_mn) {  // K-major B: the pair's halves are two boxes of n_tile / 2 rows
      if (!tmap_2d(&tb[0], (const float*)Bs[0], (uint64_t)N, (uint64_t)K, (uint64_t)ldb, (uint32_t)(n_tile / 2), BK)) return -3;
    }
    const long long units = (long long)((m_tiles2 + 1) / 2) * ((N + n_tile - 1) / n_tile) * splits;
    const int n_cl = (int)(units < sms / 2 ? units : sms / 2);
    auto launch_cl = [&](auto kern) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(2 * n_cl);
      cfg.blockDim = dim3(NTHREADS);
      cfg.dynamicSmemBytes = SMEM_BYTES;
      cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, kern, ta[0], tb[0], ta[1], tb[1], ta[2], tb[2], n_pairs, outp, out_ld, M, N, K, n_tile, splits,
                         force_atomic, kbias, part_stride, mn_lbo, mn_sbo, mn_kstep, alpha, alpha_dev);
    };
    if (a_mn && b_mn) launch_cl(k_tc_gemm<true, true, 2, 2>);
    else if (a_mn) launch_cl(k_tc_gemm<true, false, 2, 2>);
    else if (b_mn) launch_cl(k_tc_gemm<false, true, 2, 2>);
    else launch_cl(k_tc_gemm<false, false, 2, 2>);
    note_launch();
    return 0;
  }
  if (h16) {
    if (mt_sel == 1) {
      if (a_mn && b_mn) launch(k_tc_gemm<true, true, 1, 1, true>);
      else if (a_mn) launch(k_tc_gemm<true, false, 1, 1, true>);
      else if (b_mn) launch(k_tc_gemm<false, true, 1, 1, true>);
      else launch(k_tc_gemm<false, false, 1, 1, true>);
    } else {
      if (a_mn && b_mn) launch(k_tc_gemm<true, true, 2, 1, true>);
      else if (a_mn) launch(k_tc_gemm<true, false, 2, 1, true>);
      else if (b_mn) launch(k_tc_gemm<false, true, 2, 1, true>);
      else launch(k_tc_gemm<false, false, 2, 1, true>);
    }
  } else if (mt_sel == 1) {
    if (a_mn && b_mn) launch(k_tc_gemm<true, true, 1>);
    else if (a_mn) launch(k_tc_gemm<true, false, 1>);
    else if (b_mn) launch(k_tc_gemm<false, true, 1>);
    else launch(k_tc_gemm<false, false, 1>);
  } else {
    if (a_mn && b_mn) launch(k_tc_gemm<true, true, 2>);
    else if (a_mn) launch(k_tc_gemm<true, false, 2>);
    else if (b_mn) launch(k_tc_gemm<false, true, 2>);
    else launch(k_tc_gemm<false, false, 2>);
  }
  note_launch();
  if (deterministic && split) return reduce_parts(partials, splits, M, N, bias, C, ldc, st, stats) ? 1 : 0;
  return 0;
}

int gemm_tf32_multi(int n_pairs, const float* const* As, const float* const* Bs, int lda, bool a_mn, int ldb, bool b_mn,
                    float* C, int ldc, int M, int N, int K, int max_ctas, bool accumulate, const float* bias,
                    cudaStream_t st, int max_chain_kb, float* partials, int max_parts, const ColStatsOut* stats) {
  const void* a[3] = {As[0], n_pairs > 1 ? As[1] : nullptr, n_pairs > 2 ? As[2] : nullptr};
  const void* b[3] = {Bs[0], n_pairs > 1 ? Bs[1] : nullptr, n_pairs > 2 ? Bs[2] : nullptr};
  return gemm_impl(false, 1.f, nullptr, n_pairs, a, b, lda, a_mn, ldb, b_mn, C, ldc, M, N, K, max_ctas, accumulate, bias, st,
                   max_chain_kb, partials, max_parts, stats);
}

// fp16 operands (raw 16-bit storage; lda / ldb in elements, row strides multiples of 8), C += or = alpha A B^T
int gemm_f16(const void* A, int lda, bool a_mn, const void* Bm, int ldb, bool b_mn, float* C, int ldc, int M, int N, int K,
             int max_ctas, bool accumulate, float alpha, cudaStream_t st, const float* alpha_dev) {
  return gemm_impl(true, alpha, alpha_dev, 1, &A, &Bm, lda, a_mn, ldb, b_mn, C, ldc, M, N, K, max_ctas, accumulate, nullptr, st, 0,
                   nullptr, 0, nullptr);
}

bool reduce_parts(const float* parts, int n_parts, int M, int N, const float* bias, float* C, int ldc, cudaStream_t st,
                  const ColStatsOut* stats) {
  if (stats && N % 4 == 0 && ldc % 4 == 0 && M > 1) {
    static const int rows_env = getenv("DRV_RP_ROWS") ? atoi(getenv("DRV_RP_ROWS")) : 16;
    int R = (M + rows_env - 1) / rows_env;
    if (R > n_sm_tc() * 4) R = n_sm_tc() * 4;
    if (R < (M + 127) / 128) R = (M + 127) / 128;  // <= 32 rows per thread in the fp32 partial sums
    const int rows_per_cta = (M + R - 1) / R;
    R = (M + rows_per_cta - 1) / rows_per_cta;
    launch_k(k_reduce_parts_stats, dim3(1, R), dim3(64, 4), 0, st, parts, n_parts, M, N, bias, C, ldc, rows_per_cta, stats->acc,
                                                             stats->fin);
    note_launch();
    return true;
  }
  const size_t n4 = ((size_t)M * N + 3) / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > n_sm_tc() * 8) blocks = n_sm_tc() * 8;
  launch_k(k_reduce_parts, blocks, 256, 0, st, parts, n_parts, (size_t)M * N, N, bias, C, ldc);
  note_launch();
  return false;
}

int gemm_tf32(const float* A, int lda, bool a_mn, const float* Bm, int ldb, bool b_mn, float* C, int ldc, int M, int N,
              int K, int max_ctas, bool accumulate, cudaStream_t st) {
  return gemm_tf32_multi(1, &A, &Bm, lda, a_mn, ldb, b_mn, C, ldc, M, N, K, max_ctas, accumulate, nullptr, st, 0, nullptr,
                         0, nullptr);
}

}  // namespace tc
}  // namespace drv

extern "C" void drv_test_tc_gemm_cfg(int lbo, int sbo, int kstep) {
  drv::tc::g_mn_lbo = lbo; drv::tc::g_mn_sbo = sbo; drv::tc::g_mn_kstep = kstep;
}

// Test hook (tests/test_tc_gemm.py): C = A B^T on the tensor cores; C must be zeroed by the
// caller when splits > 1.
extern "C" int drv_test_tc_gemm(const float* A, int lda, int a_mn, const float* B, int ldb, int b_mn, float* C, int ldc,
                                int M, int N, int K, int splits, void* stream) {
  drv::g_launches = 0;
  drv::g_err = cudaSuccess;
  int r = drv::tc::gemm_tf32(A, lda, a_mn != 0, B, ldb, b_mn != 0, C, ldc, M, N, K, splits, false, (cudaStream_t)stream);
  if (r) return r;
  cudaError_t e = drv::g_err;
  drv::g_err = cudaSuccess;
  return (int)e;
}

// Test hook: same with fp16 operands (tests/test_tc_gemm.py)
extern "C" int drv_test_tc_gemm_h(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, float* C, int ldc,
                                  int M, int N, int K, int splits, float alpha, void* stream) {
  drv::g_launches = 0;
  drv::g_err = cudaSuccess;
  int r = drv::tc::gemm_f16(A, lda, a_mn != 0, B, ldb, b_mn != 0, C, ldc, M, N, K, splits, false, alpha, (cudaStream_t)stream);
  if (r) return r;
  cudaError_t e = drv::g_err;
  drv::g_err = cudaSuccess;
  return (int)e;
}
