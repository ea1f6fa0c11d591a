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

// lgamma(1 + x) for a count x >= 0.  Integer counts below LGTAB_N come from a table.
constexpr int LGTAB_N = 128;
static __constant__ float c_lgamma1p[LGTAB_N] = {
    0.f, 0.f, 0.69314718055994495f, 1.7917594692280554f,
    3.1780538303479449f, 4.7874917427820467f, 6.5792512120101021f, 8.5251613610654147f,
    10.604602902745249f, 12.801827480081467f, 15.104412573075514f, 17.502307845873887f,
    19.987214495661885f, 22.552163853123421f, 25.191221182738683f, 27.89927138384089f,
    30.671860106080672f, 33.505073450136891f, 36.395445208033053f, 39.339884187199495f,
    42.335616460753485f, 45.380138898476908f, 48.47118135183522f, 51.606675567764377f,
    54.784729398112319f, 58.003605222980518f, 61.261701761002008f, 64.557538627006338f,
    67.889743137181526f, 71.257038967168f, 74.658236348830172f, 78.092223553315307f,
    81.557959456115029f, 85.054467017581516f, 88.580827542197682f, 92.136175603687093f,
    95.719694542143202f, 99.330612454787428f, 102.96819861451381f, 106.63176026064346f,
    110.32063971475738f, 114.03421178146169f, 117.77188139974507f, 121.53308151543864f,
    125.3172711493569f, 129.12393363912722f, 132.95257503561632f, 136.80272263732635f,
    140.67392364823425f, 144.5657439463449f, 148.47776695177305f, 152.40959258449732f,
    156.3608363030788f, 160.3311282166309f, 164.32011226319514f, 168.32744544842765f,
    172.35279713916282f, 176.39584840699737f, 180.45629141754375f, 184.53382886144948f,
    188.62817342367163f, 192.7390472878449f, 196.86618167288998f, 201.00931639928152f,
    205.16819948264123f, 209.34258675253685f, 213.53224149456327f, 217.73693411395419f,
    221.95644181913036f, 226.1905483237276f, 230.43904356577693f, 234.70172344281826f,
    238.97838956183432f, 243.26884900298276f, 247.57291409618691f, 251.89040220972316f,
    256.22113555000954f, 260.56494097186322f, 264.92164979855283f, 269.29109765101981f,
    273.67312428569369f, 278.06757344036612f, 282.47429268763034f, 286.89313329542699f,
    291.32395009427034f, 295.7666013507606f, 300.22094864701415f, 304.68685676566867f,
    309.16419358014696f, 313.65282994987905f, 318.15263962020936f, 322.66349912672626f,
    327.1852877037752f, 331.71788719692847f, 336.26118197919851f, 340.81505887079896f,
    345.37940706226686f, 349.95411804077025f, 354.53908551944085f, 359.1342053695754f,
    363.73937555556347f, 368.35449607240474f, 372.97946888568902f, 377.61419787391861f,
    382.25858877306007f, 386.91254912321756f, 391.57598821732967f, 396.24881705179149f,
    400.93094827891576f, 405.62229616114485f, 410.32277652693728f, 415.03230672824964f,
    419.75080559954478f, 424.47819341825709f, 429.21439186665162f, 433.95932399501481f,
    438.71291418612117f, 443.47508812091894f, 448.24577274538456f, 453.02489623849618f,
    457.81238798127811f, 462.60817852687495f, 467.41219957160814f, 472.22438392698058f,
    477.04466549258569f, 481.87297922988796f, 486.70926113683942f, 491.55344822329801f};

// Result of one (cell, gene) element of the NB block.
struct NbElem {
  float ll;    // log NB(x | r, logits=ell)
  float a;     // dLL/d ell = x - (x + r) sigmoid(ell)
  float drho;  // dLL/d rho = e^rho (log sigmoid(-ell) + psi(r+x) - psi(r)) - a
};

// MUFU.RCP alone: __fdividef(1, y) wraps it in range fix-ups (6 instructions) that the NB terms never need
__device__ __forceinline__ float rcp_approx(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
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

// Copies log(x!) for x < LGTAB_N into a shared-memory table (divergent count lookups are cheap there,
// unlike a constant-memory lookup).  Needs >= LGTAB_N threads.
__device__ __forceinline__ void load_lgamma_table(float* tab_smem, int tid) {
  if (tid < LGTAB_N) tab_smem[tid] = c_lgamma1p[tid];
}

// Rare path (non-integer count, count >= LGTAB_N or r >= 1e4): generic lgamma / digamma.  Kept out of
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
// rshift_l2e: log2(e) times the per-cell shift added to log_r for the COUNT parameter only
// (r = exp(rho + shift) + eps; ell keeps the unshifted rho) - MCVNBVAE's data-split adjustment,
// mcv_nbvae.py:78-79; 0 for NBVAE / DRNBVAE.  It rides in the FFMA that feeds ex2: no extra instruction.
__device__ __forceinline__ void nb_elems(const float (&lsl)[NV], const float (&rho)[NV], const float (&x)[NV], float eps,
                                         const float* __restrict__ lgtab, NbElem (&o)[NV], float rshift_l2e = 0.f) {
  const float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f, MAGIC = 8388608.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float rexp = ex2_approx(__fmaf_rn(rho[v], LOG2E, rshift_l2e));
    const float r = rexp + eps;
    const float ell = lsl[v] - rho[v];
    const float e = ex2_approx(-fabsf(ell) * LOG2E);
    const float t = 1.f + e;
    const float sp = __fmaf_rn(lg2_approx(t), LN2, fmaxf(ell, 0.f));  // softplus(ell)
    const float xr = x[v] + MAGIC;                                     // integer test without F2I/I2F
    const bool fast = ((xr - MAGIC) == x[v]) && (x[v] < (float)LGTAB_N) && (r < 1.0e8f);
    // shifted Stirling at y0 = r + 4 and y1 = r + x + 4
    const float rx = r + x[v];
    const float y0 = r + 4.f, y1 = rx + 4.f;
    const float u0 = r * (r + 3.f), w0 = u0 + 2.f, q0 = u0 * w0;      // r(r+1)(r+2)(r+3)
    const float u1 = rx * (rx + 3.f), w1 = u1 + 2.f, q1 = u1 * w1;
    const float i0 = rcp_approx(y0), i1 = rcp_approx(y1);
    const float l0 = lg2_approx(y0) * LN2, l1 = lg2_approx(y1) * LN2;
    const float i0s = i0 * i0, i1s = i1 * i1;
    const float iq0 = rcp_approx(q0);
    const float c0 = i0 * __fmaf_rn(i0s, -2.7777777778e-3f, 8.3333333333e-2f);
    const float c1 = i1 * __fmaf_rn(i1s, -2.7777777778e-3f, 8.3333333333e-2f);
    float lgd = __fmaf_rn(y1 - 0.5f, l1, -(y0 - 0.5f) * l0) - x[v] + (c1 - c0) - lg2_approx(q1 * iq0) * LN2;
    float cnt = lgd - lgtab[__float_as_int(xr) & (LGTAB_N - 1)];
    float psd = 0.f;
    if (BWD) {
      const float iq1 = rcp_approx(q1);
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
      const float inv = rcp_approx(t);
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
