// Generic tf32 tensor-core GEMM for sm_100a:  C[M][N] (+)= sum_k A(m,k) * B(n,k)
// tcgen05.mma.kind::tf32 (fp32 operands straight from memory, fp32 accumulate in TMEM), operands
// staged by TMA with the 128-byte swizzle through a 4-stage mbarrier ring, warp-specialised:
// warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
// Each operand may be K-major ([rows][K], K contiguous) or MN-major ([K][rows], rows contiguous),
// so that d logits / activations / weights are consumed in the layout they already have:
//   dW   = dlogits^T act   : A = dlogits [B][2G] MN-major, B = act [B][h] MN-major, K = cells
//   dAct = dlogits  W      : A = dlogits [B][2G] K-major,  B = W [2G][h]  MN-major, K = 2G
// Split-K partial tiles are combined with red.global.add.f32 into a zeroed C.
#include "tc.cuh"
#include "tc_common.cuh"

#include <mutex>

namespace drv {
namespace tc {

bool make_tensor_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle) {
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
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base),
                  (const cuuint64_t*)dims, (const cuuint64_t*)strides_bytes, (const cuuint32_t*)box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// debug overrides for the MN-major descriptor fields (bytes); 0 = defaults
static int g_mn_lbo = 4096, g_mn_sbo = 512, g_mn_kstep = 1024;

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

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NTHREADS, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
          const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
          const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, int n_pairs,
          float* __restrict__ C, int ldc, int M, int N, int K, int kb_per_split, int n_tile, int use_atomic,
          const float* __restrict__ bias, int mn_lbo, int mn_sbo, int mn_kstep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * n_tile;
  // the operand pairs are concatenated along a virtual K axis: sum_p A_p B_p^T accumulates in TMEM
  const int kb_pair = (K + BK - 1) / BK;
  const int kb_total = kb_pair * n_pairs;
  const int kb0 = blockIdx.z * kb_per_split;
  const int kb1 = min(kb_total, kb0 + kb_per_split);
  const int nkb = kb1 - kb0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<MAXN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0 && nkb > 0) {
      const uint32_t tx = A_BYTES + (uint32_t)n_tile * BK * 4;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES, ph = (i / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem + s * STAGE_BYTES;
        uint8_t* sb = sa + A_BYTES;
        mbar_expect_tx(&full[s], tx);
        const int pair = (kb0 + i) / kb_pair;
        const int k0 = ((kb0 + i) - pair * kb_pair) * BK;
        const CUtensorMap* ma = pair == 0 ? &tmA : (pair == 1 ? &tmA1 : &tmA2);
        const CUtensorMap* mb = pair == 0 ? &tmB : (pair == 1 ? &tmB1 : &tmB2);
        if (A_MN) {
          for (int j = 0; j < BM / 32; ++j) tma_load_2d(sa + j * 4096, ma, &full[s], m0 + 32 * j, k0);
        } else {
          tma_load_2d(sa, ma, &full[s], k0, m0);
        }
        if (B_MN) {
          for (int j = 0; j < n_tile / 32; ++j) tma_load_2d(sb + j * 4096, mb, &full[s], n0 + 32 * j, k0);
        } else {
          tma_load_2d(sb, mb, &full[s], k0, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0 && nkb > 0) {
      const uint32_t idesc = idesc_tf32(BM, n_tile, A_MN ? 1 : 0, B_MN ? 1 : 0);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES, ph = (i / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk) {
          const uint64_t da = A_MN ? smem_desc(sa + kk * mn_kstep, mn_lbo, mn_sbo, 1) : smem_desc(sa + kk * 32, 16, 1024);
          const uint64_t db = B_MN ? smem_desc(sb + kk * mn_kstep, mn_lbo, mn_sbo, 1) : smem_desc(sb + kk * 32, 16, 1024);
          umma_tf32(tmem_base, da, db, idesc, (i | kk) != 0);
        }
        umma_commit(&empty[s]);  // frees the smem stage once these MMAs have read it
      }
      umma_commit(acc_full);
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global =====
    const int q = warp % 4;  // TMEM lane quarter this warp may access
    const int m = m0 + q * 32 + lane;
    if (nkb > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
      for (int c = 0; c < n_tile; c += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
        tmem_ld_wait();
        if (m < M) {
          float* dst = C + (size_t)m * ldc + n0 + c;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (n0 + c + j < N) {
              float o = v[j];
              if (bias != nullptr && blockIdx.z == 0) o += __ldg(&bias[n0 + c + j]);
              if (use_atomic) atomicAdd(dst + j, o);
              else dst[j] = o;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<MAXN>(tmem_base);
  }
}

}  // namespace

// A: K-major [M][K] (lda = row stride in elements) or MN-major [K][M]; same for B.
// Requirements: row strides multiples of 4 floats, bases 16-byte aligned, N tile multiple of 16
// (32 for an MN-major B).  n_pairs operand pairs (same shapes/strides) are summed: C = sum_p A_p B_p^T.
int gemm_tf32_multi(int n_pairs, const float* const* As, const float* const* Bs, int lda, bool a_mn, int ldb, bool b_mn,
                    float* C, int ldc, int M, int N, int K, int splits, bool accumulate_atomic, const float* bias,
                    cudaStream_t st) {
  if (n_pairs < 1 || n_pairs > 3) return -3;
  CUtensorMap ta[3], tb[3];
  int n_tile = N >= MAXN ? MAXN : ((N + 15) / 16) * 16;
  if (b_mn && n_tile % 32) return -3;
  for (int p = 0; p < 3; ++p) {
    int q = p < n_pairs ? p : 0;
    bool ok;
    if (a_mn) ok = tmap_2d(&ta[p], As[q], (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, 32, 2);
    else ok = tmap_2d(&ta[p], As[q], (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK);
    if (!ok) return -3;
    if (b_mn) ok = tmap_2d(&tb[p], Bs[q], (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, 32, 2);
    else ok = tmap_2d(&tb[p], Bs[q], (uint64_t)N, (uint64_t)K, (uint64_t)ldb, (uint32_t)n_tile, BK);
    if (!ok) return -3;
  }
  int kb_total = ((K + BK - 1) / BK) * n_pairs;
  if (splits < 1) splits = 1;
  if (splits > kb_total) splits = kb_total;
  int kps = (kb_total + splits - 1) / splits;
  splits = (kb_total + kps - 1) / kps;
  dim3 grid((M + BM - 1) / BM, (N + n_tile - 1) / n_tile, splits);
  int atomic = (splits > 1 || accumulate_atomic) ? 1 : 0;
  auto launch = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    kern<<<grid, NTHREADS, SMEM_BYTES, st>>>(ta[0], tb[0], ta[1], tb[1], ta[2], tb[2], n_pairs, C, ldc, M, N, K, kps,
                                             n_tile, atomic, bias, g_mn_lbo, g_mn_sbo, g_mn_kstep);
  };
  if (a_mn && b_mn) launch(k_tc_gemm<true, true>);
  else if (a_mn) launch(k_tc_gemm<true, false>);
  else if (b_mn) launch(k_tc_gemm<false, true>);
  else launch(k_tc_gemm<false, false>);
  note_launch();
  return 0;
}

int gemm_tf32(const float* A, int lda, bool a_mn, const float* Bm, int ldb, bool b_mn, float* C, int ldc, int M, int N,
              int K, int splits, bool accumulate_atomic, cudaStream_t st) {
  return gemm_tf32_multi(1, &A, &Bm, lda, a_mn, ldb, b_mn, C, ldc, M, N, K, splits, accumulate_atomic, nullptr, st);
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
