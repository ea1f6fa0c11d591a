// Gradient all-reduce over NVLink peer memory (the one collective of a data-parallel step).
// Every rank's flat gradient buffer lives in a cudaMalloc'd region that the other ranks of the node map
// through CUDA IPC.  One kernel per rank does a two-shot all-reduce with direct peer loads / stores:
//   barrier 1 (all gradients final) -> rank r sums chunk r of all ranks in rank order (so every rank
//   ends with bit-identical sums) and writes the result into chunk r of EVERY rank's buffer
//   -> barrier 2 (all chunks delivered).
// Per GPU and step: (R-1)/R of the buffer in and out over NVLink, no staging copies, no SM-hungry
// protocol kernels.  Measured against the torch-bundled NCCL 2.28 (which picks RING_LL for these 16 MB
// messages and cannot use NVLS on these boxes): see DESIGN.md section 5.
//
// Region layout: [CommHeader (1024 B)] [data].  flags[s] of rank d is written only by rank s
// (st.release.sys over NVLink) and polled only by rank d (ld.acquire.sys, local memory); the values are
// the monotonically increasing barrier numbers 2e+1 / 2e+2 of all-reduce e, so a fast neighbour that is
// already in the next all-reduce never confuses a slow one.
#include "kernels.cuh"

namespace drv {
namespace {

constexpr int COMM_MAX_RANKS = 16;
constexpr size_t COMM_HEADER_BYTES = 1024;

struct CommHeader {
  unsigned flags[64];
  unsigned epoch;   // all-reduces completed on this rank
  unsigned ticket;  // CTAs of the running all-reduce that have delivered their part
};

struct Peers {
  void* base[COMM_MAX_RANKS];
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_sys_v4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_v4(float4* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// [off4, off4 + n4): the float4 range of the buffer this launch reduces (the whole buffer, or one of the two slices of
// a step that overlaps the decoder slice with the encoder backward)
template <int WORLD>
__global__ void __launch_bounds__(512) k_peer_allreduce(Peers p, int rank, long long off4, long long n4) {
  CommHeader* me = reinterpret_cast<CommHeader*>(p.base[rank]);
  const unsigned e = ld_acquire_sys(&me->epoch);  // read by every CTA before the last one advances it
  const unsigned b1 = 2 * e + 1, b2 = 2 * e + 2;
  const int t = threadIdx.x;
  // barrier 1: my gradients are final (stream order); tell every rank, then wait for every rank
  if (blockIdx.x == 0 && t < WORLD) {
    __threadfence_system();
    st_release_sys(&reinterpret_cast<CommHeader*>(p.base[t])->flags[rank], b1);
  }
  if (t < WORLD) {
    while (ld_acquire_sys(&me->flags[t]) < b1) {
    }
  }
  __syncthreads();
  // my chunk of the buffer, in float4 units
  const long long per = (n4 + WORLD - 1) / WORLD;
  const long long c0 = off4 + min((long long)rank * per, n4), c1 = min(c0 + per, off4 + n4);
  float4* data[WORLD];
#pragma unroll
  for (int j = 0; j < WORLD; ++j) data[j] = reinterpret_cast<float4*>(reinterpret_cast<char*>(p.base[j]) + COMM_HEADER_BYTES);
  // U independent elements per thread and trip: U x WORLD 16-byte loads in flight (the loop is bound by the
  // NVLink round trip, not by bytes, when only one remote load per thread is outstanding)
  constexpr int U = WORLD >= 8 ? 2 : 8 / WORLD;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = c0 + (long long)blockIdx.x * blockDim.x + t; i < c1; i += U * stride) {
    float4 v[U][WORLD];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < WORLD; ++j)
        if (i + u * stride < c1) v[u][j] = ld_sys_v4(data[j] + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u * stride >= c1) break;
      float4 acc = v[u][0];
#pragma unroll
      for (int j = 1; j < WORLD; ++j) { acc.x += v[u][j].x; acc.y += v[u][j].y; acc.z += v[u][j].z; acc.w += v[u][j].w; }
#pragma unroll
      for (int j = 0; j < WORLD; ++j) st_sys_v4(data[j] + i + u * stride, acc);
    }
  }
  // barrier 2: the last CTA to finish announces that this rank has delivered its chunk everywhere and
  // keeps the kernel alive until every other rank has done the same
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (t == 0) last = atomicAdd(&me->ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  if (t < WORLD) {
    st_release_sys(&reinterpret_cast<CommHeader*>(p.base[t])->flags[rank], b2);
    while (ld_acquire_sys(&me->flags[t]) < b2) {
    }
  }
  __syncthreads();
  if (t == 0) {
    me->ticket = 0;
    st_release_sys(&me->epoch, e + 1);
  }
}

}  // namespace
}  // namespace drv

extern "C" size_t drv_comm_header_bytes(void) { return drv::COMM_HEADER_BYTES; }

extern "C" int drv_comm_alloc(size_t data_bytes, void** base_out, void* h_handle64) {
  if (!base_out || !h_handle64) return DRV_EINVAL;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, data_bytes + drv::COMM_HEADER_BYTES);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(p, 0, data_bytes + drv::COMM_HEADER_BYTES);
  if (e != cudaSuccess) return (int)e;
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return (int)e; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(h_handle64, &h, 64);
  *base_out = p;
  return 0;
}

extern "C" int drv_comm_open(const void* h_handle64, void** base_out) {
  if (!h_handle64 || !base_out) return DRV_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, h_handle64, 64);
  return (int)cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess);
}

extern "C" int drv_comm_close(void* base) { return base ? (int)cudaIpcCloseMemHandle(base) : 0; }
extern "C" int drv_comm_free(void* base) { return base ? (int)cudaFree(base) : 0; }

namespace drv {

// launch of one (slice of an) all-reduce; max_ctas > 0 caps the grid (a collective that runs NEXT TO compute kernels)
// threads: 512, or 256 for a slice that has to fit NEXT TO a persistent compute kernel (registers of the SM)
int peer_allreduce_launch(const Peers& p, int rank, int world, long long off4, long long n4, int max_ctas, cudaStream_t st,
                          int threads = 512) {
  const long long per = (n4 + world - 1) / world;
  int grid = (int)((per + threads - 1) / threads);
  if (grid > 148) grid = 148;  // one wave: every CTA must be able to poll the flags
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  if (grid < 1) grid = 1;
  switch (world) {
    case 2: launch_k_plain(k_peer_allreduce<2>, grid, threads, 0, st, p, rank, off4, n4); break;
    case 4: launch_k_plain(k_peer_allreduce<4>, grid, threads, 0, st, p, rank, off4, n4); break;
    case 8: launch_k_plain(k_peer_allreduce<8>, grid, threads, 0, st, p, rank, off4, n4); break;
    default: return DRV_EINVAL;
  }
  note_launch();
  return 0;
}

// Binding of the calling thread's step calls to a peer-mapped gradient buffer (drv_comm_bind): with
// DRV_FLAG_FUSED_ALLREDUCE the step launches the collective itself.
struct CommBinding {
  bool on = false;
  Peers p{};
  int rank = 0, world = 0;
  long long n_floats = 0;
};
static thread_local CommBinding g_comm;

bool comm_bound(const float* grads, long long n_floats_needed) {
  return g_comm.on && grads == reinterpret_cast<const float*>(reinterpret_cast<const char*>(g_comm.p.base[g_comm.rank]) + COMM_HEADER_BYTES) &&
         g_comm.n_floats >= n_floats_needed;
}

int comm_world() { return g_comm.on ? g_comm.world : 1; }

int comm_allreduce_slice(long long off_floats, long long n_floats, int max_ctas, cudaStream_t st, int threads) {
  if (!g_comm.on || (off_floats & 3) || (threads != 256 && threads != 512)) return DRV_EINVAL;
  return peer_allreduce_launch(g_comm.p, g_comm.rank, g_comm.world, off_floats / 4, (n_floats + 3) / 4, max_ctas, st, threads);
}

}  // namespace drv

extern "C" int drv_comm_bind(void* const* h_bases, int rank, int world, int64_t n_floats) {
  using namespace drv;
  if (!h_bases) {
    g_comm.on = false;
    return 0;
  }
  if (world < 2 || world > COMM_MAX_RANKS || rank < 0 || rank >= world || n_floats < 1) return DRV_EINVAL;
  if (world != 2 && world != 4 && world != 8) return DRV_EINVAL;
  for (int j = 0; j < world; ++j) {
    if (!h_bases[j]) return DRV_EINVAL;
    g_comm.p.base[j] = h_bases[j];
  }
  g_comm.rank = rank;
  g_comm.world = world;
  g_comm.n_floats = n_floats;
  g_comm.on = true;
  return 0;
}

extern "C" int drv_peer_allreduce(void* const* h_bases, int rank, int world, int64_t n_floats, void* stream) {
  using namespace drv;
  g_launches = 0;
  g_err = cudaSuccess;
  if (!h_bases || world < 2 || world > COMM_MAX_RANKS || rank < 0 || rank >= world || n_floats < 1) return DRV_EINVAL;
  Peers p{};
  for (int j = 0; j < world; ++j) {
    if (!h_bases[j]) return DRV_EINVAL;
    p.base[j] = h_bases[j];
  }
  const int rc = peer_allreduce_launch(p, rank, world, 0, (n_floats + 3) / 4, 0, (cudaStream_t)stream);
  if (rc != 0) return rc;  // the caller falls back to NCCL
  cudaError_t e = g_err;
  g_err = cudaSuccess;
  return (int)e;
}
