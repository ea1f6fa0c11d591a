// Lossless 16-bit staging of count tiles for the host->device copy.
// The reference ships every minibatch as a dense fp32 tensor (`xs = (x.cuda() for x in xs)`,
// src/drvish/train/__init__.py:42-43): 82 MB per 4096 x 5000 tile, which at PCIe Gen5 rates
// (~53 GB/s measured) takes 1.5 ms - longer than the whole ELBO step on a B200.  scRNA counts are
// non-negative integers, so a tile whose entries all fit in 16 bits crosses the bus as uint16 and
// is widened back to fp32 on the device, bit-exactly.  drv_host_pack_counts_u16 is the host side
// (multi-threaded, verifies representability while it packs); drv_unpack_counts_u16 the device side.
#include <atomic>
#include <thread>
#include <vector>

#include "kernels.cuh"

namespace drv {
namespace {

// returns the index of the first entry that is not an integer in [0, 65535], or n
int64_t pack_range(const float* __restrict__ src, uint16_t* __restrict__ dst, int64_t n) {
  constexpr int64_t BLK = 4096;
  for (int64_t i0 = 0; i0 < n; i0 += BLK) {
    const int64_t i1 = i0 + BLK < n ? i0 + BLK : n;
    unsigned bad = 0;
    for (int64_t i = i0; i < i1; ++i) {  // branch-free body: vectorised by the host compiler
      const float v = src[i];
      const int iv = (int)v;
      bad |= (unsigned)((float)iv != v) | (unsigned)((unsigned)iv > 65535u);
      dst[i] = (uint16_t)iv;
    }
    if (bad) {
      for (int64_t i = i0; i < i1; ++i) {
        const float v = src[i];
        if (!(v >= 0.f && v <= 65535.f) || (float)(int)v != v) return i;
      }
    }
  }
  return n;
}

__global__ void __launch_bounds__(256) k_unpack_u16(const uint16_t* __restrict__ src, float* __restrict__ dst, int64_t n) {
  pdl_prologue();
  const int64_t n8 = n / 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(src) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float4 a, b;
    a.x = (float)(w[0] & 0xffffu); a.y = (float)(w[0] >> 16); a.z = (float)(w[1] & 0xffffu); a.w = (float)(w[1] >> 16);
    b.x = (float)(w[2] & 0xffffu); b.y = (float)(w[2] >> 16); b.z = (float)(w[3] & 0xffffu); b.w = (float)(w[3] >> 16);
    reinterpret_cast<float4*>(dst)[2 * i] = a;
    reinterpret_cast<float4*>(dst)[2 * i + 1] = b;
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - n8 * 8)) dst[n8 * 8 + threadIdx.x] = (float)src[n8 * 8 + threadIdx.x];
}

// 8-bit transport (host side: host_pack8.cpp): bytes 0..254 are the counts themselves, 255 marks an escape whose fp32
// value follows in the escape list.  Widening is one 128-bit load -> four 128-bit stores per thread.
__global__ void __launch_bounds__(256) k_unpack_u8(const uint8_t* __restrict__ src, float* __restrict__ dst, int64_t n) {
  pdl_prologue();
  const int64_t n16 = n / 16;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(src) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 a;
      a.x = (float)(w[k] & 0xffu); a.y = (float)((w[k] >> 8) & 0xffu); a.z = (float)((w[k] >> 16) & 0xffu); a.w = (float)(w[k] >> 24);
      reinterpret_cast<float4*>(dst)[4 * i + k] = a;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - n16 * 16)) dst[n16 * 16 + threadIdx.x] = (float)src[n16 * 16 + threadIdx.x];
}
__global__ void __launch_bounds__(256) k_patch_escapes(const uint32_t* __restrict__ idx, const float* __restrict__ val, int n_esc,
                                                       float* __restrict__ dst, int64_t n) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_esc && (int64_t)idx[i] < n) dst[idx[i]] = val[i];
}

// Dense minibatch tile from a device-resident CSR count matrix (reference CupySparseDataLoader,
// src/drvish/data/cupy.py:36-43: `t[indices].todense()` per batch).  One warp per output row: the
// row is cleared with 128-bit stores, then its nonzeros are scattered.  HBM-bound on the dense
// write (B x G x 4 bytes); the CSR side reads 8 bytes per nonzero.
template <typename IdxT>
__global__ void __launch_bounds__(256) k_csr_gather(const IdxT* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                    const float* __restrict__ data, const int64_t* __restrict__ rows,
                                                    int B, int G, float* __restrict__ out) {
  pdl_prologue();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (warp >= B) return;
  float* o = out + (size_t)warp * G;
  if ((G & 3) == 0) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = lane; i < G / 4; i += 32) reinterpret_cast<float4*>(o)[i] = z;
  } else {
    for (int i = lane; i < G; i += 32) o[i] = 0.f;
  }
  __syncwarp();
  const int64_t r = rows ? rows[warp] : warp;
  const long long a = (long long)indptr[r], b = (long long)indptr[r + 1];
  for (long long k = a + lane; k < b; k += 32) {
    const int c = indices[k];
    if (c >= 0 && c < G) o[c] = data[k];
  }
}

}  // namespace
}  // namespace drv

extern "C" int drv_csr_gather_rows(const void* indptr, int indptr_is_int64, const int32_t* indices, const float* data,
                                   const int64_t* rows, int n_rows, int n_genes, float* out, void* stream) {
  using namespace drv;
  g_launches = 0;
  g_err = cudaSuccess;
  if (!indptr || !indices || !data || !out || n_rows < 0 || n_genes < 1) return DRV_EINVAL;
  if (n_rows == 0) return 0;
  const unsigned blocks = (unsigned)((n_rows * 32 + 255) / 256);
  if (indptr_is_int64)
    launch_k(k_csr_gather<long long>, blocks, 256, 0, (cudaStream_t)stream, (const long long*)indptr, indices, data, rows, n_rows,
             n_genes, out);
  else
    launch_k(k_csr_gather<int>, blocks, 256, 0, (cudaStream_t)stream, (const int*)indptr, indices, data, rows, n_rows, n_genes,
             out);
  note_launch();
  cudaError_t e = g_err;
  g_err = cudaSuccess;
  return (int)e;
}

extern "C" int64_t drv_host_pack_counts_u16(const float* h_src, uint16_t* h_dst, int64_t n, int n_threads) {
  if (!h_src || !h_dst || n < 0) return -1;
  if (n_threads < 1) n_threads = 1;
  const int64_t min_chunk = 1 << 16;
  if ((int64_t)n_threads > (n + min_chunk - 1) / min_chunk) n_threads = (int)((n + min_chunk - 1) / min_chunk);
  if (n_threads <= 1) return drv::pack_range(h_src, h_dst, n);
  std::atomic<int64_t> first_bad{n};
  std::vector<std::thread> pool;
  const int64_t chunk = ((n + n_threads - 1) / n_threads + 63) / 64 * 64;
  for (int t = 0; t < n_threads; ++t) {
    const int64_t a = (int64_t)t * chunk, b = a + chunk < n ? a + chunk : n;
    if (a >= b) break;
    pool.emplace_back([=, &first_bad] {
      const int64_t r = drv::pack_range(h_src + a, h_dst + a, b - a);
      if (r < b - a) {
        int64_t cur = first_bad.load();
        while (a + r < cur && !first_bad.compare_exchange_weak(cur, a + r)) {}
      }
    });
  }
  for (auto& th : pool) th.join();
  return first_bad.load();
}

extern "C" int drv_unpack_counts_u8(const uint8_t* src, const uint32_t* esc_idx, const float* esc_val, int n_escapes, float* dst,
                                    int64_t n, void* stream) {
  using namespace drv;
  g_launches = 0;
  g_err = cudaSuccess;
  if (!src || !dst || n < 0 || n_escapes < 0 || (n_escapes > 0 && (!esc_idx || !esc_val))) return DRV_EINVAL;
  if (n == 0) return 0;
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15)) return DRV_EINVAL;
  int64_t blocks = (n / 16 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  launch_k(k_unpack_u8, (unsigned)blocks, 256, 0, (cudaStream_t)stream, src, dst, n);
  note_launch();
  if (n_escapes > 0) {
    launch_k(k_patch_escapes, (unsigned)((n_escapes + 255) / 256), 256, 0, (cudaStream_t)stream, esc_idx, esc_val, n_escapes, dst, n);
    note_launch();
  }
  cudaError_t e = g_err;
  g_err = cudaSuccess;
  return (int)e;
}

extern "C" int drv_unpack_counts_u16(const uint16_t* src, float* dst, int64_t n, void* stream) {
  using namespace drv;
  g_launches = 0;
  g_err = cudaSuccess;
  if (!src || !dst || n < 0) return DRV_EINVAL;
  if (n == 0) return 0;
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15)) return DRV_EINVAL;
  int64_t blocks = (n / 8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  launch_k(k_unpack_u16, (unsigned)blocks, 256, 0, (cudaStream_t)stream, src, dst, n);
  note_launch();
  cudaError_t e = g_err;
  g_err = cudaSuccess;
  return (int)e;
}
