// Shared device helpers for the drvish_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/drvish_b200.h"

#define DRV_HALF_LOG_2PI 0.91893853320467274178f

namespace drv {

// launch bookkeeping (host side, per thread) -------------------------------------------------
extern thread_local int g_launches;
extern thread_local cudaError_t g_err;

inline void note_launch() {
  ++g_launches;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess && g_err == cudaSuccess) g_err = e;
}

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Programmatic dependent launch: every kernel of the step is launched with the "programmatic stream
// serialization" attribute and starts with pdl_prologue().  launch_dependents lets the NEXT kernel of
// the stream be scheduled as soon as all CTAs of this one are resident (its launch latency and CTA
// start-up overlap this kernel's execution); griddepcontrol.wait then blocks until the PREVIOUS
// kernel has completed and flushed its memory, so data dependencies are exactly those of plain
// stream order.  The step is a chain of ~60 short kernels: this removes most of the per-boundary
// bubble.  DRV_PDL=0 in the environment turns the attribute off (the instructions are then no-ops).
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args);

// Plain stream-ordered launch.  For persistent kernels that fill an SM's shared memory: launched as
// a programmatic dependent such a kernel becomes RESIDENT while it waits for its predecessor and
// blocks the SMs for the kernels of the other stream (seen in the graph timeline: a 32-CTA weight-
// gradient GEMM idling on 32 SMs for 48 us behind the previous side-stream GEMM).
template <typename... KArgs, typename... Args>
inline void launch_k_plain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = nullptr;
  cfg.numAttrs = 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<Args&&>(args)...);
}

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<Args&&>(args)...);
}

// Side stream for work that is off the critical chain of a step (weight-gradient GEMMs, the
// dose-response head): forked from / joined to the caller's stream with events, so the whole step
// still looks like one stream to the caller (and is CUDA-graph capturable).
struct SideStream {
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_aux = nullptr;
  bool used = false;
  // work enqueued on `side` after this call sees everything enqueued on `main` so far
  void fork(cudaStream_t main) {
    cudaEventRecord(ev_fork, main);
    cudaStreamWaitEvent(side, ev_fork, 0);
    used = true;
  }
  // split form: mark the fork point now, keep submitting critical-path work to `main`, and attach
  // the side stream later - kernels that become ready together are dispatched in submission order,
  // so the critical-path kernel gets the SMs first
  void mark(cudaStream_t main) { cudaEventRecord(ev_fork, main); }
  void attach() {
    cudaStreamWaitEvent(side, ev_fork, 0);
    used = true;
  }
  // a second, named join point: `main` waits for the side-stream work submitted before record_aux()
  // only (join() would also wait for whatever was queued on the side stream afterwards)
  void record_aux() { cudaEventRecord(ev_aux, side); }
  void wait_aux(cudaStream_t main) { cudaStreamWaitEvent(main, ev_aux, 0); }
  // work enqueued on `main` after this call sees everything enqueued on `side` so far
  void join(cudaStream_t main) {
    if (!used) return;
    cudaEventRecord(ev_join, side);
    cudaStreamWaitEvent(main, ev_join, 0);
    used = false;
  }
};
SideStream* side_stream();  // per-thread, created on first use (nullptr if creation failed)

// Finalisation folded into the statistics kernel: the last CTA to finish (ticket counter) turns the
// fp64 sums into mean / rstd and updates the running buffers (momentum, unbiased variance) - one
// launch less per BatchNorm.
struct StatFinal {
  unsigned int* ticket;  // zeroed counter
  int B;
  float eps, momentum;
  float *mean, *rstd, *rmean, *rvar;  // rmean/rvar may be null
};

// One ticket per column block (blockIdx.x): the last of its gridDim.y row-split CTAs finalises the
// block's own columns [col0, col0 + ncols), so the finalisation is spread over gridDim.x CTAs.
__device__ __forceinline__ void stat_finalize_if_last(const StatFinal& f, const double* acc, int D, int col0, int ncols) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid == 0) is_last = atomicAdd(f.ticket + blockIdx.x, 1u) == gridDim.y - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int nthr = blockDim.x * blockDim.y;
  const int cend = min(D, col0 + ncols);
  for (int i = col0 + tid; i < cend; i += nthr) {
    const double m = __ldcg(&acc[i]) / f.B;
    double var = __ldcg(&acc[D + i]) / f.B - m * m;
    if (var < 0.0) var = 0.0;
    f.mean[i] = (float)m;
    // the fp64 part ends here (sum of squares minus squared mean); sqrt / divide in fp32 keep the
    // single finalising CTA short
    f.rstd[i] = 1.0f / sqrtf((float)(var + (double)f.eps));
    if (f.rmean) {
      const float unbiased = (float)var * ((float)f.B / (float)(f.B - 1));
      f.rmean[i] = (1.0f - f.momentum) * f.rmean[i] + f.momentum * (float)m;
      f.rvar[i] = (1.0f - f.momentum) * f.rvar[i] + f.momentum * unbiased;
    }
  }
}

// Where a producer kernel (the split-K reduction of a Linear) leaves the BatchNorm statistics of
// its OUTPUT for the next block: acc = [2][D] zeroed fp64 sums.
struct ColStatsOut {
  double* acc;
  StatFinal fin;
};

// BatchNorm -> ReLU applied to one element (modules.py:19-20).  Explicit fma so that the
// forward value and the recomputed ReLU gate of the backward are bit-identical.
struct BnParams {
  const float* mean;   // batch mean per column
  const float* rstd;   // 1/sqrt(biased var + eps)
  const float* gamma;  // BatchNorm weight
  const float* beta;   // BatchNorm bias
};

// sqrt of a count (encoder input transform, modules.py:56).  MUFU.SQRT (<= 1 ulp): the IEEE sqrtf
// takes a slow fix-up path for x = 0 - about half of all scRNA counts - which ncu showed to dominate
// otherwise memory-bound kernels.  Used everywhere so that forward values and recomputed gates agree.
__device__ __forceinline__ float sqrt_count(float v) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

__device__ __forceinline__ float bn_xhat(float u, float mean, float rstd) { return (u - mean) * rstd; }
__device__ __forceinline__ float bn_y(float xhat, float gamma, float beta) { return __fmaf_rn(xhat, gamma, beta); }
__device__ __forceinline__ float bn_relu(float u, float mean, float rstd, float gamma, float beta) {
  return fmaxf(bn_y(bn_xhat(u, mean, rstd), gamma, beta), 0.f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// special functions
// ---------------------------------------------------------------------------------------------
// digamma for y > 0: recurrence up to y >= 6, then the asymptotic series (next term 1/(240 y^8)).
__device__ __forceinline__ float digamma_pos(float y) {
  float acc = 0.f;
  while (y < 6.f) {
    acc -= 1.f / y;
    y += 1.f;
  }
  float inv = 1.f / y, inv2 = inv * inv;
  float series = inv2 * (0.083333333333f - inv2 * (0.0083333333333f - inv2 * 0.003968253968f));
  return acc + logf(y) - 0.5f * inv - series;
}

// lgamma(1 + x) for a count x >= 0.  Integer counts below 32 come from a table.
static __constant__ float c_lgamma1p[32] = {
    0.f, 0.f, 0.69314718055994531f, 1.79175946922805500f, 3.17805383034794562f, 4.78749174278204599f,
    6.57925121201010100f, 8.52516136106541430f, 10.6046029027452502f, 12.8018274800814696f,
    15.1044125730755153f, 17.5023078458738858f, 19.9872144956618861f, 22.5521638531234229f,
    25.1912211827386815f, 27.8992713838408916f, 30.6718601060806728f, 33.5050734501368889f,
    36.3954452080330536f, 39.3398841871994940f, 42.3356164607534850f, 45.3801388984769080f,
    48.4711813518352239f, 51.6066755677643736f, 54.7847293981123192f, 58.0036052229805199f,
    61.2617017610020020f, 64.5575386270063311f, 67.8897431371815350f, 71.2570389671680090f,
    74.6582363488301644f, 78.0922235533153106f};

// Result of one (cell, gene) element of the NB block.
struct NbElem {
  float ll;    // log NB(x | r, logits=ell)
  float a;     // dLL/d ell = x - (x + r) sigmoid(ell)
  float drho;  // dLL/d rho = e^rho (log sigmoid(-ell) + psi(r+x) - psi(r)) - a
};

__device__ __forceinline__ float ex2_approx(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float lg2_approx(float v) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// Copies log(x!) for x < 32 into a 32-float shared-memory table (one bank per entry: divergent
// count lookups are conflict-free, unlike a constant-memory lookup).
__device__ __forceinline__ void load_lgamma_table(float* tab_smem, int tid) {
  if (tid < 32) tab_smem[tid] = c_lgamma1p[tid];
}

// Rare path (non-integer count, count >= 32 or r >= 1e4): generic lgamma / digamma.  Kept out of
// line so that the hot epilogue loop stays small enough for the instruction cache.
static __device__ __noinline__ float2 nb_count_slow(float r, float x, bool bwd) {
  float2 o;
  o.x = lgammaf(r + x) - lgammaf(r) - lgammaf(1.f + x);
  o.y = bwd ? digamma_pos(r + x) - digamma_pos(r) : 0.f;
  return o;
}

// NV elements of the NB block per thread (straight-line code; the NV independent elements give
// the instruction-level parallelism).
//   lsl = library + s - logsumexp(s)  (so ell = lsl - rho),  rho = log_r,  x = count >= 0.
// Follows torch.distributions.NegativeBinomial.log_prob (negative_binomial.py:130-150):
//   r logsig(-ell) + x logsig(ell) + lgamma(r+x) - lgamma(1+x) - lgamma(r),   r = e^rho + eps
// written as  x ell - (x + r) softplus(ell) + [lgamma(r+x) - lgamma(r)] - log(x!).
// The bracket and psi(r+x) - psi(r) are formed as DIFFERENCES of shifted Stirling series
//   lgamma(y) = F(y+4) - log(y(y+1)(y+2)(y+3)),  F(z) = (z-1/2) log z - z + 1/(12z) - 1/(360z^3)
//   psi(y)    = G(y+4) - Q'(y)/Q(y),             G(z) = log z - 1/(2z) - 1/(12z^2) + 1/(120z^4)
// (truncation error < 1e-6 for z >= 4; the constant log(2 pi)/2 cancels in the difference), with
// log(x!) from a 32-entry shared-memory table; x = 0 returns exactly 0.  Counts >= 32 or
// non-integer take the generic lgamma / digamma route.  No int<->float conversions (they share
// the SFU pipe with the transcendentals).
template <bool BWD, int NV>
__device__ __forceinline__ void nb_elems(const float (&lsl)[NV], const float (&rho)[NV], const float (&x)[NV], float eps,
                                         const float* __restrict__ lgtab, NbElem (&o)[NV]) {
  const float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f, MAGIC = 8388608.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float rexp = ex2_approx(rho[v] * LOG2E);
    const float r = rexp + eps;
    const float ell = lsl[v] - rho[v];
    const float e = ex2_approx(-fabsf(ell) * LOG2E);
    const float t = 1.f + e;
    const float sp = __fmaf_rn(lg2_approx(t), LN2, fmaxf(ell, 0.f));  // softplus(ell)
    const float xr = x[v] + MAGIC;                                     // integer test without F2I/I2F
    const bool fast = ((xr - MAGIC) == x[v]) && (x[v] < 32.f) && (r < 1.0e8f);
    // shifted Stirling at y0 = r + 4 and y1 = r + x + 4
    const float rx = r + x[v];
    const float y0 = r + 4.f, y1 = rx + 4.f;
    const float u0 = r * (r + 3.f), w0 = u0 + 2.f, q0 = u0 * w0;      // r(r+1)(r+2)(r+3)
    const float u1 = rx * (rx + 3.f), w1 = u1 + 2.f, q1 = u1 * w1;
    const float i0 = __fdividef(1.f, y0), i1 = __fdividef(1.f, y1);
    const float l0 = lg2_approx(y0) * LN2, l1 = lg2_approx(y1) * LN2;
    const float i0s = i0 * i0, i1s = i1 * i1;
    const float iq0 = __fdividef(1.f, q0);
    const float c0 = i0 * __fmaf_rn(i0s, -2.7777777778e-3f, 8.3333333333e-2f);
    const float c1 = i1 * __fmaf_rn(i1s, -2.7777777778e-3f, 8.3333333333e-2f);
    float lgd = __fmaf_rn(y1 - 0.5f, l1, -(y0 - 0.5f) * l0) - x[v] + (c1 - c0) - lg2_approx(q1 * iq0) * LN2;
    float cnt = lgd - lgtab[__float_as_int(xr) & 31];
    float psd = 0.f;
    if (BWD) {
      const float iq1 = __fdividef(1.f, q1);
      const float g0 = l0 - 0.5f * i0 - i0s * __fmaf_rn(i0s, -8.3333333333e-3f, 8.3333333333e-2f);
      const float g1 = l1 - 0.5f * i1 - i1s * __fmaf_rn(i1s, -8.3333333333e-3f, 8.3333333333e-2f);
      const float qd0 = (u0 + w0) * __fmaf_rn(2.f, r, 3.f);          // Q'(r)
      const float qd1 = (u1 + w1) * __fmaf_rn(2.f, rx, 3.f);
      psd = (g1 - g0) - (qd1 * iq1 - qd0 * iq0);
    }
    if (x[v] == 0.f) {
      cnt = 0.f;
      psd = 0.f;
    }
    if (!fast) {
      const float2 sl = nb_count_slow(r, x[v], BWD);
      cnt = sl.x;
      psd = sl.y;
    }
    o[v].ll = __fmaf_rn(x[v], ell, cnt) - rx * sp;
    if (BWD) {
      const float inv = __fdividef(1.f, t);
      const float sig = ell >= 0.f ? inv : e * inv;
      o[v].a = x[v] - rx * sig;
      o[v].drho = rexp * (psd - sp) - o[v].a;
    } else {
      o[v].a = 0.f;
      o[v].drho = 0.f;
    }
  }
}

template <bool BWD>
__device__ __forceinline__ NbElem nb_elem(float lsl, float rho, float x, float eps, const float* __restrict__ lgtab) {
  const float a[1] = {lsl}, b[1] = {rho}, c[1] = {x};
  NbElem o[1];
  nb_elems<BWD, 1>(a, b, c, eps, lgtab, o);
  return o[0];
}

// softplus / its derivative exactly as torch.nn.functional.softplus (beta=1, threshold=20)
__device__ __forceinline__ float softplus_t(float v) { return v > 20.f ? v : log1pf(expf(v)); }
__device__ __forceinline__ float softplus_grad_t(float v) {
  if (v > 20.f) return 1.f;
  float z = expf(v);
  return z / (z + 1.f);
}

}  // namespace drv
