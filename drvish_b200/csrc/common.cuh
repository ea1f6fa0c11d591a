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
  // third stream for a collective that runs next to the step (data parallel, DRV_FLAG_FUSED_ALLREDUCE): it waits for
  // BOTH streams (comm_fork) and is joined to `main` before anything that must follow it (comm_join)
  cudaStream_t comm = nullptr;
  cudaEvent_t ev_cm = nullptr, ev_cs = nullptr, ev_cdone = nullptr;
  bool comm_used = false;
  void comm_fork(cudaStream_t main) {
    cudaEventRecord(ev_cm, main);
    cudaEventRecord(ev_cs, side);
    cudaStreamWaitEvent(comm, ev_cm, 0);
    cudaStreamWaitEvent(comm, ev_cs, 0);
    comm_used = true;
  }
  void comm_join(cudaStream_t main) {
    if (!comm_used) return;
    cudaEventRecord(ev_cdone, comm);
    cudaStreamWaitEvent(main, ev_cdone, 0);
    comm_used = false;
  }
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

// ---------------------------------------------------------------------------------------------
// packed fp32 pairs: sm_100 has FFMA2 / FMUL2 / FADD2 (PTX fma/mul/add/sub .f32x2 on 64-bit
// registers).  One issue slot does the same operation on two elements: the NB epilogue is bound by
// instruction issue, so its arithmetic runs on element PAIRS.
// ---------------------------------------------------------------------------------------------
struct f2 {
  unsigned long long v;
};
__device__ __forceinline__ f2 mk2(float a, float b) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ f2 bc2(float a) { return mk2(a, a); }
__device__ __forceinline__ void un2(f2 p, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  f2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}

// The fast route of the NB element math covers integer counts 0 <= x < LGTAB_N and r < NB_R_BIG; everything
// else (a warp-rare event with real data) goes through nb_elem_generic.  NB_R_BIG is where the fast route's
// error takes off: it forms (r + 7/2) log((r+x+4) / (r+4)) from two MUFU.LG2 results whose relative error
// (2^-22) is multiplied by r log2 r - ~5e-3 absolute per element at r = 2048, twice that at 4096
// (tools/nb_math_probe.py; the reference's own fp32 arithmetic, lgamma(r+x) - lgamma(r), loses as much there).
// The generic route uses log1p / the Stirling difference in a form without large cancelling terms.
constexpr float NB_R_BIG = 2048.f;

// Generic route: any count x >= 0 (non-integer, >= LGTAB_N), any r.  Out of line so that the hot epilogue loop
// stays small (instruction cache); precision first.
//   lgamma(r+x) - lgamma(r) and psi(r+x) - psi(r):
//     r >= 32: Stirling without shift, differences formed analytically (u = x / r):
//        (r - 1/2) log1p(u) + x (log(r+x) - 1) + [S(r+x) - S(r)],   S(y) = 1/(12y) - 1/(360y^3)
//        log1p(u) - (i1 - i0)/2 - (i1^2 - i0^2)/12 + (i1^4 - i0^4)/120,   i0 = 1/r, i1 = 1/(r+x)
//     r <  32: lgammaf / digamma differences (magnitudes <= ~100: no harmful cancellation)
// returns (ll, a, ndrho) by value: reference outputs of an out-of-line function would live on the stack
template <bool BWD>
static __device__ __noinline__ float3 nb_elem_generic(float lsl, float rho, float x, float eps, float rshift_l2e) {
  float ll, a, ndrho;
  const float rexp = exp2f(__fmaf_rn(rho, 1.4426950408889634f, rshift_l2e));
  const float r = rexp + eps;
  const float ell = lsl - rho;
  const float e = expf(-fabsf(ell));
  const float L = log1pf(e);
  const float spp = L + fmaxf(ell, 0.f), spn = L + fmaxf(-ell, 0.f);  // softplus(ell), softplus(-ell)
  float cnt = 0.f, psd = 0.f;
  if (x != 0.f) {
    const float rx = r + x;
    if (r >= 32.f) {
      const float i0 = 1.f / r, i1 = 1.f / rx;
      const float l1p = log1pf(x * i0);
      const float d = -x * i0 * i1;  // i1 - i0
      const float s = i0 + i1, p = i0 * i1, s2 = s * s;
      cnt = (r - 0.5f) * l1p + x * (logf(rx) - 1.f) + d * (8.3333333333e-2f - (s2 - p) * 2.7777777778e-3f);
      if (BWD) psd = l1p - d * (0.5f + s * (8.3333333333e-2f - (s2 - 2.f * p) * 8.3333333333e-3f));
    } else {
      cnt = lgammaf(rx) - lgammaf(r);
      if (BWD) psd = digamma_pos(rx) - digamma_pos(r);
    }
    cnt -= lgammaf(1.f + x);
  }
  ll = cnt - x * spn - r * spp;
  if (BWD) {
    const float inv = 1.f / (1.f + e);
    const float sg = ell >= 0.f ? inv : e * inv, sn = ell >= 0.f ? e * inv : inv;  // sigmoid(ell), sigmoid(-ell)
    a = x * sn - r * sg;
    ndrho = rexp * (spp - psd) + a;
  } else {
    a = 0.f;
    ndrho = 0.f;
  }
  return make_float3(ll, a, ndrho);
}

// Two elements of the NB block on packed registers (straight-line code).
//   lsl = library + s - logsumexp(s)  (so ell = lsl - rho),  rho = log_r,  x = count >= 0.
// Follows torch.distributions.NegativeBinomial.log_prob (negative_binomial.py:130-150):
//   ll = r logsig(-ell) + x logsig(ell) + lgamma(r+x) - lgamma(1+x) - lgamma(r),   r = e^rho + eps
//      = cnt - x softplus(-ell) - r softplus(ell),          cnt = lgamma(r+x) - lgamma(r) - log(x!)
//   a     = dLL/d ell = x sigmoid(-ell) - r sigmoid(ell)                     
//   ndrho = -dLL/d rho = e^rho (softplus(ell) - [psi(r+x) - psi(r)]) + a
// lgamma(r+x) - lgamma(r) and psi(r+x) - psi(r) are DIFFERENCES of shifted Stirling series
//   lgamma(y) = F(y+4) - log Q(y),  F(z) = (z-1/2) log z - z + 1/(12z) - 1/(360z^3),  Q(y) = y(y+1)(y+2)(y+3)
//   psi(y)    = G(y+4) - Q'(y)/Q(y),  G(z) = log z - 1/(2z) - 1/(12z^2) + 1/(120z^4)
// (truncation < 1e-6 for z >= 4) with every difference taken analytically, in terms of
//   y0 = r+4, y1 = r+x+4, i01 = 1/(y0 y1), dpos = x i01 = 1/y0 - 1/y1, s = (y0+y1) i01 = 1/y0 + 1/y1:
//   F(y1)-F(y0) = x log y1 + (y0-1/2) log(y1/y0) - x - dpos (1/12 - (s^2 - i01)/360)
//   G(y1)-G(y0) = log(y1/y0) + dpos (1/2 + s (1/12 - (s^2 - 2 i01)/120))
//   Q'/Q(r+x) - Q'/Q(r) = (Q'(r+x) Q(r) - Q'(r) Q(r+x)) / (Q(r) Q(r+x))
// so x = 0 gives exactly 0 and nothing of size r log r is subtracted from anything.  8 MUFU per element
// (2 ex2, 4 lg2, 2 rcp: the reciprocals of y0 y1 (1 + e^-|ell|) and of Q(r) Q(r+x) serve five quotients);
// log(x!) from the shared-memory table.  No int<->float conversions (they share the SFU pipe).
// rshift_l2e: log2(e) times the per-cell shift added to log_r for the COUNT parameter only
// (r = exp(rho + shift) + eps; ell keeps the unshifted rho) - MCVNBVAE's data-split adjustment,
// mcv_nbvae.py:78-79; 0 for NBVAE / DRNBVAE.  It rides in the FFMA that feeds ex2: no extra instruction.
// Returns false when one of the two elements is outside the fast route's domain (the caller then replaces
// that element's results with nb_elem_generic's).
// xb = (x + 2^23) - 2^23: the nearest integer for 0 <= x < 2^22
__device__ __forceinline__ bool nb_fast_domain(float x, float xb, float r) {
  return (__float_as_uint(x) < __float_as_uint((float)LGTAB_N)) && (xb == x) && (r < NB_R_BIG);
}
__device__ __forceinline__ bool nb_fast_domain(float x, float r) { return nb_fast_domain(x, (x + 8388608.f) - 8388608.f, r); }

template <bool BWD>
__device__ __forceinline__ bool nb_pair(f2 lsl, f2 rho, f2 x, float eps, const float* lgtab, float rshift_l2e, f2& ll,
                                        f2& a, f2& ndrho) {
  const float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f, MAGIC = 8388608.f;
  float s0, s1;
  un2(fma2(rho, bc2(LOG2E), bc2(rshift_l2e)), s0, s1);
  const f2 rexp = mk2(ex2_approx(s0), ex2_approx(s1));
  const f2 r = add2(rexp, bc2(eps));
  const f2 nr = fma2(rexp, bc2(-1.f), bc2(-eps));
  const f2 ell = sub2(lsl, rho);
  float el0, el1;
  un2(ell, el0, el1);
  const f2 e = mk2(ex2_approx(-fabsf(el0) * LOG2E), ex2_approx(-fabsf(el1) * LOG2E));
  const f2 t = add2(e, bc2(1.f));
  un2(t, s0, s1);
  const f2 spp = fma2(mk2(lg2_approx(s0), lg2_approx(s1)), bc2(LN2), mk2(fmaxf(el0, 0.f), fmaxf(el1, 0.f)));  // softplus(ell)
  const f2 nspn = sub2(ell, spp);                                                                             // -softplus(-ell)
  // Q(r), Q(r+x), shifted arguments
  const f2 rx = add2(r, x);
  const f2 y0 = add2(r, bc2(4.f)), y1 = add2(rx, bc2(4.f));
  const f2 u0 = mul2(r, add2(r, bc2(3.f))), q0 = mul2(u0, add2(u0, bc2(2.f)));
  const f2 u1 = mul2(rx, add2(rx, bc2(3.f))), q1 = mul2(u1, add2(u1, bc2(2.f)));
  // two reciprocals for five quotients
  const f2 P = mul2(y0, y1);
  un2(mul2(P, t), s0, s1);
  const f2 rP = mk2(rcp_approx(s0), rcp_approx(s1));
  const f2 i01 = mul2(rP, t), inv = mul2(rP, P);  // 1/(y0 y1), 1/(1 + e)
  un2(mul2(q0, q1), s0, s1);
  const f2 rQ = mk2(rcp_approx(s0), rcp_approx(s1));
  un2(y0, s0, s1);
  const f2 l0 = mk2(lg2_approx(s0), lg2_approx(s1));
  un2(y1, s0, s1);
  const f2 l1 = mk2(lg2_approx(s0), lg2_approx(s1));
  un2(mul2(mul2(q1, rQ), q1), s0, s1);
  const f2 lgr = mk2(lg2_approx(s0), lg2_approx(s1));  // log2(Q(r+x) / Q(r))
  const f2 dl = sub2(l1, l0);                           // log2(y1 / y0)
  const f2 main = fma2(x, l1, mul2(add2(y0, bc2(-0.5f)), dl));
  const f2 dpos = mul2(x, i01);
  const f2 s = mul2(add2(y0, y1), i01);
  const f2 S2 = mul2(s, s);
  const f2 cc = mul2(dpos, fma2(sub2(S2, i01), bc2(2.7777777778e-3f), bc2(-8.3333333333e-2f)));
  // log(x!): table lookup on the integer part (entries outside the fast domain are replaced by the caller)
  const f2 xr = add2(x, bc2(MAGIC));
  un2(xr, s0, s1);
  const f2 lf = mk2(lgtab[__float_as_int(s0) & (LGTAB_N - 1)], lgtab[__float_as_int(s1) & (LGTAB_N - 1)]);
  const f2 cnt = sub2(fma2(sub2(main, lgr), bc2(LN2), cc), add2(x, lf));
  ll = fma2(x, nspn, fma2(nr, spp, cnt));
  float x0, x1, r0, r1, xb0, xb1;
  un2(x, x0, x1);
  un2(r, r0, r1);
  un2(sub2(xr, bc2(MAGIC)), xb0, xb1);
  const bool fast = nb_fast_domain(x0, xb0, r0) && nb_fast_domain(x1, xb1, r1);
  if (BWD) {
    const f2 gd = fma2(dl, bc2(LN2), mul2(dpos, fma2(s, fma2(fma2(i01, bc2(-2.f), S2), bc2(-8.3333333333e-3f), bc2(8.3333333333e-2f)), bc2(0.5f))));
    // (Q'(r+x) Q(r) - Q'(r) Q(r+x)) / (Q(r) Q(r+x)),  Q'(y) = 2 (u+1) (2y+3),  u = y (y+3)
    const f2 qd0n = mul2(add2(u0, bc2(1.f)), fma2(r, bc2(-2.f), bc2(-3.f)));
    const f2 qd1 = mul2(add2(u1, bc2(1.f)), fma2(rx, bc2(2.f), bc2(3.f)));
    const f2 zq = mul2(fma2(qd1, q0, mul2(qd0n, q1)), rQ);
    const f2 psd = fma2(zq, bc2(-2.f), gd);  // psi(r+x) - psi(r)
    const f2 einv = mul2(e, inv);
    float i0, i1, ei0, ei1;
    un2(inv, i0, i1);
    un2(einv, ei0, ei1);
    const f2 sg = mk2(el0 >= 0.f ? i0 : ei0, el1 >= 0.f ? i1 : ei1);  // sigmoid(ell)
    const f2 sn = mk2(el0 >= 0.f ? ei0 : i0, el1 >= 0.f ? ei1 : i1);  // sigmoid(-ell)
    a = fma2(x, sn, mul2(nr, sg));
    ndrho = fma2(rexp, sub2(spp, psd), a);
  }
  return fast;
}

// Result of one (cell, gene) element of the NB block (scalar interface of the CUDA-core path).
struct NbElem {
  float ll;    // log NB(x | r, logits=ell)
  float a;     // dLL/d ell = x sigmoid(-ell) - r sigmoid(ell)
  float drho;  // dLL/d rho = e^rho (psi(r+x) - psi(r) - softplus(ell)) - a
};

// NV (even) elements per thread through nb_pair; elements outside the fast domain are redone by the generic route.
template <bool BWD, int NV>
__device__ __forceinline__ void nb_elems(const float (&lsl)[NV], const float (&rho)[NV], const float (&x)[NV], float eps,
                                         const float* __restrict__ lgtab, NbElem (&o)[NV], float rshift_l2e = 0.f) {
  static_assert(NV % 2 == 0, "pairs");
  bool fast = true;
#pragma unroll
  for (int v = 0; v < NV; v += 2) {
    f2 ll, a = bc2(0.f), nd = bc2(0.f);
    fast &= nb_pair<BWD>(mk2(lsl[v], lsl[v + 1]), mk2(rho[v], rho[v + 1]), mk2(x[v], x[v + 1]), eps, lgtab, rshift_l2e, ll, a, nd);
    un2(ll, o[v].ll, o[v + 1].ll);
    un2(a, o[v].a, o[v + 1].a);
    float n0, n1;
    un2(nd, n0, n1);
    o[v].drho = -n0;
    o[v + 1].drho = -n1;
  }
  if (!fast) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float r = exp2f(__fmaf_rn(rho[v], 1.4426950408889634f, rshift_l2e)) + eps;
      if (!nb_fast_domain(x[v], r)) {
        const float3 g = nb_elem_generic<BWD>(lsl[v], rho[v], x[v], eps, rshift_l2e);
        o[v].ll = g.x;
        o[v].a = g.y;
        o[v].drho = -g.z;
      }
    }
  }
}

// softplus / its derivative exactly as torch.nn.functional.softplus (beta=1, threshold=20)
__device__ __forceinline__ float softplus_t(float v) { return v > 20.f ? v : log1pf(expf(v)); }
__device__ __forceinline__ float softplus_grad_t(float v) {
  if (v > 20.f) return 1.f;
  float z = expf(v);
  return z / (z + 1.f);
}

}  // namespace drv
