// Host-side orchestration of one SVI step and the C ABI (include/drvish_b200.h).
// Replaces, for one minibatch, everything between svi.step(*xs) being entered and the
// per-parameter .grad tensors being populated: guide trace (encoder, rsample),
// model trace (priors, decoder, NB likelihood, dose-response head), ELBO assembly and autograd.
// Everything is enqueued on the caller's stream; no allocation, no host synchronisation.
#include <string.h>

#include <stdlib.h>

#include "kernels.cuh"
#include "tc.cuh"

namespace drv {

thread_local int g_launches = 0;
thread_local cudaError_t g_err = cudaSuccess;

// Optional section timing with CUDA events recorded on the step's own stream (bench.py's live
// roofline measurement).  Sections: see DRV_SEC_* in the public header.
namespace prof {
constexpr int MAX_STEPS = 64;
bool enabled = false;
int n_steps = 0;
cudaEvent_t ev[MAX_STEPS][DRV_N_SECTIONS + 1];
cudaEvent_t kev[MAX_STEPS][DRV_N_PROFILED_KERNELS][2];  // start/stop pairs around single kernels
bool created = false;
inline void mark(int i, cudaStream_t st) {
  if (enabled && n_steps < MAX_STEPS) cudaEventRecord(ev[n_steps][i], st);
}
inline void kmark(int k, int which, cudaStream_t st) {
  if (enabled && n_steps < MAX_STEPS) cudaEventRecord(kev[n_steps][k][which], st);
}
}  // namespace prof

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DRV_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

SideStream* side_stream() {
  // one set of streams / events per (thread, device): a stream belongs to the device that was current when it was
  // created, and a thread may drive models on several GPUs
  constexpr int MAX_DEV = 16;
  static thread_local SideStream pool[MAX_DEV];
  static thread_local bool tried[MAX_DEV] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return nullptr;
  SideStream& ss = pool[dev];
  if (!tried[dev]) {
    tried[dev] = true;
    // high priority: the side stream's kernels are short (operand conversions at the start of the step, the dose-response
    // head) or small grids (trunk weight gradients); DRV_SIDE_PRIORITY=0 restores the default priority.  (No effect was
    // seen inside a captured graph - profiles/r02_graph_timeline_cfg3.txt -; kept for the eager path.)
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    const char* pe = getenv("DRV_SIDE_PRIORITY");
    const int prio = (pe && atoi(pe) == 0) ? prio_lo : prio_hi;
    if (cudaStreamCreateWithPriority(&ss.side, cudaStreamNonBlocking, prio) != cudaSuccess ||
        cudaEventCreateWithFlags(&ss.ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ss.ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ss.ev_aux, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ss.comm, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ss.ev_cm, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ss.ev_cs, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ss.ev_cdone, cudaEventDisableTiming) != cudaSuccess) {
      ss.side = nullptr;
      cudaGetLastError();
    }
  }
  return ss.side ? &ss : nullptr;
}

namespace {

constexpr int64_t ALIGN_F = 32;  // tensors start on 128-byte boundaries inside the flat buffers

inline int64_t up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct Block {      // one BatchNorm1d -> ReLU -> Linear triple of make_fc (modules.py:15-23)
  int din, dout;
  int64_t gamma, beta, W, bias;   // float offsets in the flat parameter buffer
  int64_t rmean, rvar;            // float offsets in the flat BN-buffer
};

struct Net {
  int n_blocks;
  Block enc[DRV_MAX_LAYERS + 1], dec[DRV_MAX_LAYERS + 1];
  int64_t w_loc, w_scale, b_loc, b_scale;  // dose-response guide parameters (D > 0)
  int64_t n_param_floats, n_bn_floats;
  int n_bn;
};

bool valid(const drv_dims* d) {
  if (!d) return false;
  if (d->n_cells < 2 || d->n_genes < 1 || d->n_latent < 1) return false;
  if (d->n_layers < 1 || d->n_layers > DRV_MAX_LAYERS) return false;
  for (int i = 0; i < d->n_layers; ++i)
    if (d->layers[i] < 1) return false;
  if (d->n_drugs < 0) return false;
  if (d->n_drugs > 0 && (d->n_classes < 1 || d->n_doses_total < d->n_drugs)) return false;
  return true;
}

Net make_net(const drv_dims& d) {
  Net n{};
  n.n_blocks = d.n_layers + 1;
  int64_t p = 0, q = 0;
  auto stack = [&](Block* blk, const int* dims) {
    for (int k = 0; k < n.n_blocks; ++k) {
      Block& b = blk[k];
      b.din = dims[k];
      b.dout = dims[k + 1];
      b.gamma = p; p = up(p + b.din, ALIGN_F);
      b.beta = p;  p = up(p + b.din, ALIGN_F);
      b.W = p;     p = up(p + (int64_t)b.dout * b.din, ALIGN_F);
      b.bias = p;  p = up(p + b.dout, ALIGN_F);
      b.rmean = q; q = up(q + b.din, ALIGN_F);
      b.rvar = q;  q = up(q + b.din, ALIGN_F);
    }
  };
  int ed[DRV_MAX_LAYERS + 2], dd[DRV_MAX_LAYERS + 2];
  ed[0] = d.n_genes;
  dd[0] = d.n_latent;
  for (int i = 0; i < d.n_layers; ++i) {
    ed[i + 1] = d.layers[i];
    dd[i + 1] = d.layers[d.n_layers - 1 - i];
  }
  ed[d.n_layers + 1] = 2 * d.n_latent + 2;
  dd[d.n_layers + 1] = 2 * d.n_genes;
  stack(n.enc, ed);
  stack(n.dec, dd);
  if (d.n_drugs > 0) {
    n.w_loc = p;   p = up(p + (int64_t)d.n_drugs * d.n_latent, ALIGN_F);
    n.w_scale = p; p = up(p + (int64_t)d.n_drugs * d.n_latent, ALIGN_F);
    n.b_loc = p;   p = up(p + d.n_doses_total, ALIGN_F);
    n.b_scale = p; p = up(p + d.n_doses_total, ALIGN_F);
  }
  n.n_param_floats = p;
  n.n_bn_floats = q;
  n.n_bn = 2 * n.n_blocks;
  return n;
}

// workspace carve-up (byte offsets, 256-byte aligned)
struct Plan {
  size_t zero_begin, zero_bytes;            // region cleared at the start of every step
  size_t stat_acc[2][DRV_MAX_LAYERS + 1];   // [enc/dec][block]: double[2*din] forward statistics
  size_t grad_acc[2][DRV_MAX_LAYERS + 1];   // double[2*din] backward sums
  size_t status;                             // uint32
  size_t ticket[2][DRV_MAX_LAYERS + 1];      // uint32 'last CTA' counters of the statistics kernels
  size_t mean[2][DRV_MAX_LAYERS + 1], rstd[2][DRV_MAX_LAYERS + 1];
  size_t V[2][DRV_MAX_LAYERS + 1];          // Linear outputs (B x dout); dec last = logits B x 2G
  size_t Gb[2][DRV_MAX_LAYERS + 1];         // dY / dU of each block input (B x din), block 0 of enc unused
  size_t z, l, lse, ll, asum, dO, rshift;
  size_t u, m, gfac, du, dz_drs;
  size_t tc;                                 // scratch of the fused tcgen05 decoder path
  size_t p_hi, p_lo;                         // tf32 hi / lo split of the whole flat parameter buffer
  size_t tcb[2][DRV_MAX_LAYERS + 1];         // scratch of the tensor-core blocks (0 = block runs on CUDA cores)
  size_t total;
};

Plan make_plan(const drv_dims& d, const Net& n) {
  Plan p{};
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = (o + bytes + 255) / 256 * 256; return r; };
  const size_t B = d.n_cells;
  p.zero_begin = o;
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < n.n_blocks; ++k) {
      const Block& b = s == 0 ? n.enc[k] : n.dec[k];
      p.stat_acc[s][k] = take(sizeof(double) * 2 * b.din);
      p.grad_acc[s][k] = take(sizeof(double) * 2 * b.din);
      p.ticket[s][k] = take(4 * (size_t)(b.din / 32 + 2));  // one counter per column block of the statistics kernels
    }
  p.ll = take(4 * B);    // per-cell log-likelihood / dLL/dl sums: atomic accumulation
  p.asum = take(4 * B);
  p.status = take(256);
  p.zero_bytes = o - p.zero_begin;
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < n.n_blocks; ++k) {
      const Block& b = s == 0 ? n.enc[k] : n.dec[k];
      p.mean[s][k] = take(4 * (size_t)b.din);
      p.rstd[s][k] = take(4 * (size_t)b.din);
      p.V[s][k] = take(4 * B * b.dout);
      p.Gb[s][k] = (s == 0 && k == 0) ? 0 : take(4 * B * b.din);
    }
  p.z = take(4 * B * d.n_latent);
  p.l = take(4 * B);
  p.lse = take(4 * B);
  p.rshift = take(4 * B);  // MCVNBVAE: log_split_complement - log_split per cell
  p.dO = take(4 * B * (2 * d.n_latent + 2));
  if (d.n_drugs > 0) {
    p.u = take(4 * B * d.n_drugs);
    p.du = take(4 * B * d.n_drugs);
    p.dz_drs = take(4 * B * d.n_latent);
    p.m = take(4 * (size_t)d.n_classes * d.n_doses_total);
    p.gfac = take(4 * (size_t)d.n_classes * d.n_doses_total);
  }
  p.p_hi = take(4 * (size_t)n.n_param_floats);
  p.p_lo = take(4 * (size_t)n.n_param_floats);
  p.tc = o;
  o += tc_workspace_bytes(d);
  o = (o + 255) / 256 * 256;
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < n.n_blocks; ++k) {
      const Block& b = s == 0 ? n.enc[k] : n.dec[k];
      p.tcb[s][k] = 0;
      if (s == 1 && k == n.n_blocks - 1) continue;  // fused decoder path
      drv_hparams hp0{};
      if (tc_block_forward_supported(d.n_cells, b.din, b.dout, hp0)) {
        p.tcb[s][k] = o;
        o += tc_block_workspace_bytes(d.n_cells, b.din, b.dout);
        o = (o + 255) / 256 * 256;
      }
    }
  p.total = o;
  return p;
}


// ---- shape padding (k_pad.cu): the next aligned problem for shapes the tensor-core kernels cannot take ----------
inline int up_to(int v, int a) { return (v + a - 1) / a * a; }

drv_dims padded_dims(const drv_dims& d) {
  drv_dims p = d;
  p.n_genes = up_to(d.n_genes, 8);
  for (int i = 0; i < d.n_layers; ++i) p.layers[i] = up_to(d.layers[i], 32);
  return p;
}

// pad when it turns a CUDA-core shape into a tensor-core one: genes not a multiple of 8 (16-byte-aligned halves of the
// fp16 d logits, TMA row strides) or a hidden width not a multiple of 32 (k-blocks, MN-major GEMM tiles)
bool pad_wanted(const drv_dims& d) {
  bool ragged = d.n_genes % 8 != 0;
  for (int i = 0; i < d.n_layers; ++i) ragged |= d.layers[i] % 32 != 0;
  // minibatches below one 128-cell tile stay on the fp32 CUDA-core path: nothing to gain there, and the tensor-core
  // operands' 11-bit rounding noise is averaged over too few cells (measured: 3e-3 on a 7-entry BatchNorm gradient at 96 cells)
  if (!ragged || d.n_genes < 64 || d.n_cells < 128) return false;
  static const bool off = getenv("DRV_NO_PAD") != nullptr;  // debug / A-B
  return !off && up_to(d.layers[0], 32) <= 512;  // the fused decoder takes hidden widths up to 512
}

struct PadPlan {
  drv_dims dp;
  Net nr, np;
  PadSegs params, bn;
  size_t off_params, off_grads, off_bn, off_x, off_x2, total;  // byte offsets in the workspace, behind the padded plan
};

void add_seg(PadSegs& S, long long src_off, long long dst_off, int src_rows, int src_cols, int dst_rows, int dst_cols, float pad0,
             int n_blocks = 1, int blk_stride = 0, float pad1 = 0.f) {
  PadSeg& g = S.seg[S.n++];
  g.src_off = src_off; g.dst_off = dst_off; g.src_rows = src_rows; g.src_cols = src_cols; g.dst_rows = dst_rows; g.dst_cols = dst_cols;
  g.n_blocks = n_blocks; g.blk_stride = blk_stride; g.pad[0] = pad0; g.pad[1] = pad1;
}

PadPlan make_pad_plan(const drv_dims& d, size_t plan_total) {
  PadPlan P{};
  P.dp = padded_dims(d);
  P.nr = make_net(d);
  P.np = make_net(P.dp);
  const int G = d.n_genes, Gp = P.dp.n_genes;
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < P.nr.n_blocks; ++k) {
      const Block& r = s == 0 ? P.nr.enc[k] : P.nr.dec[k];
      const Block& q = s == 0 ? P.np.enc[k] : P.np.dec[k];
      const bool last_dec = s == 1 && k == P.nr.n_blocks - 1;  // [scale half | log_r half] rows: each half is padded on its own
      add_seg(P.params, r.gamma, q.gamma, r.din, 1, q.din, 1, 1.f);
      add_seg(P.params, r.beta, q.beta, r.din, 1, q.din, 1, 0.f);
      if (last_dec) {
        add_seg(P.params, r.W, q.W, G, r.din, 2 * Gp, q.din, 0.f, 2, Gp, 0.f);
        // pad genes: scale-half bias -1e30 (softmax probability exactly 0, NB terms exactly 0), log_r-half bias 0
        add_seg(P.params, r.bias, q.bias, G, 1, 2 * Gp, 1, -1.0e30f, 2, Gp, 0.f);
      } else {
        add_seg(P.params, r.W, q.W, r.dout, r.din, q.dout, q.din, 0.f);
        add_seg(P.params, r.bias, q.bias, r.dout, 1, q.dout, 1, 0.f);
      }
      add_seg(P.bn, r.rmean, q.rmean, r.din, 1, q.din, 1, 0.f);
      add_seg(P.bn, r.rvar, q.rvar, r.din, 1, q.din, 1, 1.f);
    }
  if (d.n_drugs > 0) {
    const int DL = d.n_drugs * d.n_latent, T = d.n_doses_total;
    add_seg(P.params, P.nr.w_loc, P.np.w_loc, DL, 1, DL, 1, 0.f);
    add_seg(P.params, P.nr.w_scale, P.np.w_scale, DL, 1, DL, 1, 0.f);
    add_seg(P.params, P.nr.b_loc, P.np.b_loc, T, 1, T, 1, 0.f);
    add_seg(P.params, P.nr.b_scale, P.np.b_scale, T, 1, T, 1, 0.f);
  }
  size_t o = (plan_total + 255) / 256 * 256;
  auto take = [&](size_t bytes) { size_t r = o; o = (o + bytes + 255) / 256 * 256; return r; };
  P.off_params = take(4 * (size_t)P.np.n_param_floats);
  P.off_grads = take(4 * ((size_t)P.np.n_param_floats + 32));
  P.off_bn = take(4 * (size_t)P.np.n_bn_floats);
  P.off_x = take(4 * (size_t)d.n_cells * Gp);
  P.off_x2 = take(4 * (size_t)d.n_cells * Gp);
  P.total = o;
  return P;
}

struct Ctx {
  const drv_dims& d;
  const drv_hparams& hp;
  const Net& n;
  const Plan& p;
  char* ws;
  const float* params;
  float* grads;
  float* bnbuf;
  cudaStream_t st;
  SideStream* side;  // may be null: everything then runs on st
  bool enc0_w_prepared = false;  // fp16 weight parts of the encoder input block were produced at the start of the step
  bool w_prepared[2][DRV_MAX_LAYERS + 1] = {};  // ... and of the trunk blocks
  template <class T> T* at(size_t off) const { return reinterpret_cast<T*>(ws + off); }
  BnParams bn(int s, int k) const {
    const Block& b = s == 0 ? n.enc[k] : n.dec[k];
    return BnParams{at<float>(p.mean[s][k]), at<float>(p.rstd[s][k]), params + b.gamma, params + b.beta};
  }
};

// One BatchNorm1d -> ReLU -> Linear block, forward: batch statistics (+ running buffers), then
// optionally the Linear with the BN/ReLU (and sqrt for the encoder input) applied in its prologue.
// stats_done: the previous block's Linear already left this block's BatchNorm statistics (fused into
// its split-K reduction).  Returns whether THIS block's Linear did the same for block k + 1.
bool block_forward(const Ctx& c, int s, int k, const float* U0, bool do_linear, bool stats_done = false) {
  const int B = c.d.n_cells;
  const bool update = !(c.hp.flags & DRV_FLAG_NO_BN_UPDATE) && c.bnbuf;
  const Block& b = s == 0 ? c.n.enc[k] : c.n.dec[k];
  const float* U = k == 0 ? U0 : c.at<float>(c.p.V[s][k - 1]);
  const bool sq = (s == 0 && k == 0);
  double* acc = c.at<double>(c.p.stat_acc[s][k]);
  if (!stats_done)
    colstats(U, B, b.din, sq, acc, c.at<unsigned int>(c.p.ticket[s][k]), c.hp.bn_eps, c.hp.bn_momentum,
             c.at<float>(c.p.mean[s][k]), c.at<float>(c.p.rstd[s][k]), update ? c.bnbuf + b.rmean : nullptr,
             update ? c.bnbuf + b.rvar : nullptr, c.st);
  if (!do_linear) return false;
  if (c.p.tcb[s][k] && tc_block_forward_supported(B, b.din, b.dout, c.hp) && (sq ? !(c.hp.flags & 32) : !(c.hp.flags & 8))) {
    TcBlockArgs a{};
    a.U = U; a.sqrt_in = sq; a.bn = c.bn(s, k); a.W = c.params + b.W; a.bias = c.params + b.bias;
    a.W_hi = c.at<float>(c.p.p_hi) + b.W; a.W_lo = c.at<float>(c.p.p_lo) + b.W;
    a.B = B; a.din = b.din; a.dout = b.dout; a.V = c.at<float>(c.p.V[s][k]); a.scratch = c.ws + c.p.tcb[s][k];
    a.no_fused_prologue = (c.hp.flags & 512) != 0;
    a.w_prepared = (sq && c.enc0_w_prepared) || c.w_prepared[s][k];
    ColStatsOut ns{};
    const int n_blocks = c.n.n_blocks;
    if (k + 1 < n_blocks && !(c.hp.flags & 1024)) {  // 1024 (debug): separate statistics kernels
      const Block& nb = s == 0 ? c.n.enc[k + 1] : c.n.dec[k + 1];
      ns.acc = c.at<double>(c.p.stat_acc[s][k + 1]);
      ns.fin = StatFinal{c.at<unsigned int>(c.p.ticket[s][k + 1]), B, c.hp.bn_eps, c.hp.bn_momentum,
                         c.at<float>(c.p.mean[s][k + 1]), c.at<float>(c.p.rstd[s][k + 1]),
                         update ? c.bnbuf + nb.rmean : nullptr, update ? c.bnbuf + nb.rvar : nullptr};
      a.next_stats = &ns;
    }
    return tc_block_forward(a, c.st);
  } else {
    const int n_blocks = c.n.n_blocks;
    static const bool narrow_off = getenv("DRV_NARROW_FUSED") != nullptr && atoi(getenv("DRV_NARROW_FUSED")) == 0;
    if (!sq && b.din <= 64 && k + 1 < n_blocks && !narrow_off && !(c.hp.flags & (1024 | DRV_FLAG_FORCE_SIMT))) {
      // narrow input (latent -> hidden): Linear + the next block's BatchNorm statistics in one launch
      const Block& nb = s == 0 ? c.n.enc[k + 1] : c.n.dec[k + 1];
      ColStatsOut ns{};
      ns.acc = c.at<double>(c.p.stat_acc[s][k + 1]);
      ns.fin = StatFinal{c.at<unsigned int>(c.p.ticket[s][k + 1]), B, c.hp.bn_eps, c.hp.bn_momentum,
                         c.at<float>(c.p.mean[s][k + 1]), c.at<float>(c.p.rstd[s][k + 1]),
                         update ? c.bnbuf + nb.rmean : nullptr, update ? c.bnbuf + nb.rvar : nullptr};
      if (simt_narrow_linear_forward_stats(U, B, b.din, c.bn(s, k), c.params + b.W, c.params + b.bias, b.dout,
                                           c.at<float>(c.p.V[s][k]), ns, c.st))
        return true;
    }
    simt_linear_forward(U, B, b.din, sq ? 2 : 1, c.bn(s, k), c.params + b.W, c.params + b.bias, b.dout,
                        c.at<float>(c.p.V[s][k]), c.st);
  }
  return false;
}

// Backward of block k given dV of its Linear: dW, db, then dY = (dV W) * relu-gate, then the
// BatchNorm backward (d gamma, d beta, dU in place).  `linear_done`: dW/db/dY were already
// produced (fused tcgen05 path).  The first encoder block needs no dU (x is data): its dY is
// reduced to d gamma / d beta on the fly.
// dv_prepared: the block above already left this block's tf32-rounded dV and bias gradient
// (bn_bwd_apply_round).  Returns whether this block did the same for block k - 1.
// tc_part (tensor-core block only): 0 = everything, 1 = weight gradient only, 2 = the rest after a part-1 call.
bool block_backward(const Ctx& c, int s, int k, const float* U0, const float* dV, bool linear_done,
                    bool dv_prepared = false, int tc_part = 0) {
  bool gated = false;  // Gb holds the ungated dAct: the BatchNorm backward kernels apply the ReLU gate
  const int B = c.d.n_cells;
  const Block& b = s == 0 ? c.n.enc[k] : c.n.dec[k];
  const float* U = k == 0 ? U0 : c.at<float>(c.p.V[s][k - 1]);
  const bool sq = (s == 0 && k == 0);
  const int pro = sq ? 2 : 1;
  double* gacc = c.at<double>(c.p.grad_acc[s][k]);
  if (!linear_done) {
    if (c.p.tcb[s][k] && tc_block_backward_supported(B, b.din, b.dout, c.hp) && (sq ? !(c.hp.flags & 32) : !(c.hp.flags & 4))) {
      // the forward of this block ran on the tensor cores too: its A_hi / W_hi operands are reused
      TcBlockArgs a{};
      a.U = U; a.sqrt_in = sq; a.bn = c.bn(s, k); a.W = c.params + b.W; a.B = B; a.din = b.din; a.dout = b.dout;
      a.W_hi = c.at<float>(c.p.p_hi) + b.W; a.W_lo = c.at<float>(c.p.p_lo) + b.W;
      a.scratch = c.ws + c.p.tcb[s][k]; a.dV = dV; a.dW = c.grads + b.W; a.db = c.grads + b.bias;
      a.dY = sq ? nullptr : c.at<float>(c.p.Gb[s][k]); a.gacc = gacc; a.side = c.side;
      a.no_fused_prologue = (c.hp.flags & 512) != 0;
      a.dv_prepared = dv_prepared;
      a.part = tc_part;
      tc_block_backward(a, c.st);
      gated = false;
      if (tc_part == 1) return false;
    } else {
      // the weight gradient is off the critical chain: side stream
      cudaStream_t ws = c.st;
      if (c.side) { c.side->fork(c.st); ws = c.side->side; }
      simt_linear_wgrad(dV, B, b.dout, U, b.din, pro, c.bn(s, k), c.grads + b.W, c.grads + b.bias, ws);
      if (sq)
        simt_linear_dgrad_bngrad(dV, B, b.dout, c.params + b.W, b.din, U, pro, c.bn(s, k), gacc, c.st);
      else
        simt_linear_dgrad_gate(dV, B, b.dout, c.params + b.W, b.din, U, pro, c.bn(s, k), c.at<float>(c.p.Gb[s][k]), c.st);
    }
  }
  if (sq) {
    bn_grad_finalize(gacc, b.din, c.grads + b.gamma, c.grads + b.beta, c.st);
    return false;
  }
  float* dY = c.at<float>(c.p.Gb[s][k]);
  if (linear_done) gated = false;  // fused decoder path: dAct arrives ungated
  bn_bwd_reduce(dY, U, B, b.din, c.bn(s, k), gacc, !gated, c.st);
  if (k >= 1 && !(c.hp.flags & 2048)) {  // 2048 (debug): separate round / column-sum kernel
    const Block& nb = s == 0 ? c.n.enc[k - 1] : c.n.dec[k - 1];
    const bool nsq = (s == 0 && k - 1 == 0);
    if (c.p.tcb[s][k - 1] && tc_block_backward_supported(B, nb.din, nb.dout, c.hp) &&
        (nsq ? !(c.hp.flags & 32) : !(c.hp.flags & 4))) {
      bn_bwd_apply_round(dY, U, B, b.din, c.bn(s, k), gacc, c.grads + b.gamma, c.grads + b.beta, !gated,
                         tc_block_rounded_dv(c.ws + c.p.tcb[s][k - 1], B, nb.din, nb.dout), c.grads + nb.bias, c.st);
      return true;
    }
  }
  bn_bwd_apply(dY, U, B, b.din, c.bn(s, k), gacc, c.grads + b.gamma, c.grads + b.beta, !gated, c.st);
  return false;
}

int finish() {
  cudaError_t e = g_err;
  g_err = cudaSuccess;
  return (int)e;
}

}  // namespace
}  // namespace drv

using namespace drv;

extern "C" {

const char* drv_version(void) { return "drvish_b200 0.2.0 sm_100a"; }

int drv_shape_padded(const drv_dims* dims) { return valid(dims) && pad_wanted(*dims) ? 1 : 0; }

int drv_last_launch_count(void) { return g_launches; }

int drv_param_offsets(const drv_dims* dims, int64_t* offsets, int64_t* sizes, int max_n, int64_t* total_floats) {
  if (!valid(dims)) return DRV_EINVAL;
  Net n = make_net(*dims);
  int cnt = 0;
  auto put = [&](int64_t off, int64_t sz) {
    if (cnt < max_n) {
      if (offsets) offsets[cnt] = off;
      if (sizes) sizes[cnt] = sz;
    }
    ++cnt;
  };
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < n.n_blocks; ++k) {
      const Block& b = s == 0 ? n.enc[k] : n.dec[k];
      put(b.gamma, b.din);
      put(b.beta, b.din);
      put(b.W, (int64_t)b.dout * b.din);
      put(b.bias, b.dout);
    }
  if (dims->n_drugs > 0) {
    put(n.w_loc, (int64_t)dims->n_drugs * dims->n_latent);
    put(n.w_scale, (int64_t)dims->n_drugs * dims->n_latent);
    put(n.b_loc, dims->n_doses_total);
    put(n.b_scale, dims->n_doses_total);
  }
  if (total_floats) *total_floats = n.n_param_floats;
  return cnt;
}

int drv_bn_offsets(const drv_dims* dims, int64_t* offsets, int64_t* sizes, int max_n, int64_t* total_floats) {
  if (!valid(dims)) return DRV_EINVAL;
  Net n = make_net(*dims);
  int cnt = 0;
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < n.n_blocks; ++k) {
      const Block& b = s == 0 ? n.enc[k] : n.dec[k];
      if (cnt + 1 < max_n) {
        if (offsets) { offsets[cnt] = b.rmean; offsets[cnt + 1] = b.rvar; }
        if (sizes) { sizes[cnt] = b.din; sizes[cnt + 1] = b.din; }
      }
      cnt += 2;
    }
  if (total_floats) *total_floats = n.n_bn_floats;
  return cnt;
}

size_t drv_workspace_bytes(const drv_dims* dims) {
  if (!valid(dims)) return 0;
  if (pad_wanted(*dims)) {  // the step runs as the next aligned problem: its plan + the padded copies
    const drv_dims dp = padded_dims(*dims);
    return make_pad_plan(*dims, make_plan(dp, make_net(dp)).total).total;
  }
  Net n = make_net(*dims);
  return make_plan(*dims, n).total;
}

// x feeds the encoder, x_lik is scored by the likelihood (the same tile except for MCVNBVAE: x0 / x1);
// log_split / log_split_c (both or neither): per-cell data-split terms of MCVNBVAE, log_r += lsc - ls
// for the NB count parameter (mcv_nbvae.py:78-79).
static int elbo_step_core(const drv_dims* dims, const drv_hparams* hp, const float* x, const float* x_lik,
                          const float* log_split, const float* log_split_c, const int64_t* labels, const float* y,
                          const float* eps_z, const float* eps_l, const float* w_prior, const float* b_prior,
                          const float* eps_bias, const int32_t* dose_offsets, const float* params, float* grads,
                          float* bn_running, int64_t* bn_batches, drv_step_out* out, void* workspace,
                          size_t workspace_bytes, void* stream) {
  if (!valid(dims) || !hp || !x || !x_lik || !eps_z || !eps_l || !params || !out || !workspace) return DRV_EINVAL;
  if ((log_split == nullptr) != (log_split_c == nullptr)) return DRV_EINVAL;
  const drv_dims& d = *dims;
  if (d.n_drugs > 0 && (!labels || !y || !w_prior || !b_prior || !eps_bias || !dose_offsets)) return DRV_EINVAL;
  Net n = make_net(d);
  Plan p = make_plan(d, n);
  if (workspace_bytes < p.total) return DRV_ENOSPC;
  cudaStream_t st = (cudaStream_t)stream;
  SideStream* side = (hp->flags & 64) ? nullptr : side_stream();
  Ctx c{d, *hp, n, p, (char*)workspace, params, grads, bn_running, st, side};
  const int B = d.n_cells, G = d.n_genes, L = d.n_latent;
  const bool bwd = grads != nullptr;

  uint32_t* status = c.at<uint32_t>(p.status);
  const int NB = n.n_blocks;
  // Two-part execution for data parallel (DRV_FLAG_STOP_AFTER_DECODER / DRV_FLAG_ENCODER_BACKWARD_ONLY):
  // the caller starts the all-reduce of the decoder + dose-response gradient slice between the
  // parts, so that it overlaps the encoder backward.
  const bool part2_only = bwd && (hp->flags & DRV_FLAG_ENCODER_BACKWARD_ONLY);
  const bool stop_after_decoder = bwd && (hp->flags & DRV_FLAG_STOP_AFTER_DECODER);
  // Data parallel with the collective inside the step: the decoder / dose-response slice of the flat gradient buffer
  // (and the loss / status tail slots) is final before the encoder backward starts - it is all-reduced on the side
  // stream by a few CTAs next to the encoder backward; the encoder slice follows at the end of the step.
  const bool fused_ar = bwd && (hp->flags & DRV_FLAG_FUSED_ALLREDUCE);
  long long ar_tail_floats = 0;  // what the collective at the end of the step still has to reduce: [0, ar_tail_floats)
  if (fused_ar && (part2_only || stop_after_decoder || !side || !comm_bound(grads, n.n_param_floats + DRV_GRAD_TAIL_FLOATS)))
    return DRV_EINVAL;
  if (part2_only) {
    for (int k = NB - 1, prep = 0; k >= 0; --k)
      prep = block_backward(c, 0, k, x, k == NB - 1 ? c.at<float>(p.dO) : c.at<float>(p.Gb[0][k + 1]), false, prep);
    if (side) side->join(st);
    if (bn_batches && !(hp->flags & DRV_FLAG_NO_BN_UPDATE)) bump_counters(bn_batches, n.n_bn, st);
    return finish();
  }

  cudaMemsetAsync(c.ws + p.zero_begin, 0, p.zero_bytes, st);
  prof::mark(0, st);
  // Start-of-step work that is off the critical chain runs on the side stream, next to the
  // statistics of x: the result record and the whole gradient buffer are cleared ONCE (every
  // gradient kernel then accumulates - no memset nodes inside the backward chain), and the tf32
  // hi / lo parts of every weight are produced in one pass over the flat parameter buffer (the
  // tensor-core kernels take pre-rounded operands: kind::tf32 truncates).
  const bool tc_any = !(hp->flags & DRV_FLAG_FORCE_SIMT);
  const bool dec_tc = tc_decoder_supported(d, *hp) && !(hp->flags & 16);
  bool enc_sd0 = false, dec_w_prepared = false;
  {
    cudaStream_t ss = st;
    if (side) { side->fork(st); ss = side->side; }
    cudaMemsetAsync(out, 0, sizeof(drv_step_out), ss);
    // encoder weights first: they are all the encoder forward waits for
    const size_t enc_floats = (size_t)n.dec[0].gamma;
    if (tc_any) split_params_tf32(params, enc_floats, c.at<float>(p.p_hi), c.at<float>(p.p_lo), ss);
    if (side && p.tcb[0][0] && tc_block_forward_supported(B, n.enc[0].din, n.enc[0].dout, *hp) && !(hp->flags & 32)) {
      TcBlockArgs a{};  // fp16 hi / lo parts of the encoder input weight, next to the tf32 split
      a.sqrt_in = true; a.no_fused_prologue = (hp->flags & 512) != 0; a.W = params + n.enc[0].W;
      a.B = B; a.din = n.enc[0].din; a.dout = n.enc[0].dout; a.scratch = c.ws + p.tcb[0][0];
      c.enc0_w_prepared = tc_block_prepare(a, ss);
    }
    if (side && tc_any)
      for (int s2 = 0; s2 < 2; ++s2)
        for (int k = (s2 == 0 ? 1 : 0); k < NB - (s2 == 1 ? 1 : 0); ++k) {  // trunk blocks (not the input layer, not the fused decoder)
          const Block& b = s2 == 0 ? n.enc[k] : n.dec[k];
          if (!p.tcb[s2][k] || !tc_block_forward_supported(B, b.din, b.dout, *hp) || (hp->flags & 8)) continue;
          TcBlockArgs a{};
          a.sqrt_in = false; a.no_fused_prologue = (hp->flags & 512) != 0; a.W = params + b.W;
          a.B = B; a.din = b.din; a.dout = b.dout; a.scratch = c.ws + p.tcb[s2][k];
          c.w_prepared[s2][k] = tc_block_prepare(a, ss);
        }
    bool enc_sd_ = false;
    if (side) {
      block_forward(c, 0, 0, x, false);  // statistics of x (main stream) overlap the side-stream work
      enc_sd_ = true;
      side->join(st);
      side->used = true;  // the rest of the start-up work stays behind on the side stream (joined after the sampling)
    }
    enc_sd0 = enc_sd_;
  }
  // Start-of-step work the DECODER needs (tf32 parts / fp16 copy of its weights, the cleared gradient buffers): queued on
  // the side stream BEHIND the encoder input layer, so that it runs next to the small grids of the encoder trunk.
  // Submitted at the start it delayed the encoder forward by 8 us: the block scheduler drains the CTAs of earlier
  // submitted kernels first, and the gradient memset competed with the x tiles for HBM (profiles/r02_graph_timeline_*.txt).
  auto late_prep = [&]() {
    cudaStream_t ss = st;
    if (side) { side->fork(st); ss = side->side; }
    const size_t enc_floats = (size_t)n.dec[0].gamma;
    if (tc_any) {
      // the fp16-operand decoder reads its own fp16 copy of the last Linear's weight (2G x h, by far the largest
      // tensor): its tf32 parts are not needed then
      static const bool dec_tf32_env = getenv("DRV_DEC_TF32") != nullptr;
      const Block& lastb = n.dec[NB - 1];
      const bool skip_last = dec_tc && !dec_tf32_env && G % 8 == 0 && lastb.din >= 64;
      const size_t w0 = (size_t)lastb.W, w1 = w0 + (size_t)lastb.dout * lastb.din;
      auto split = [&](size_t a, size_t b) {
        if (b > a) split_params_tf32(params + a, b - a, c.at<float>(p.p_hi) + a, c.at<float>(p.p_lo) + a, ss);
      };
      if (skip_last) {
        split(enc_floats, w0);
        split(w1, (size_t)n.n_param_floats);
      } else {
        split(enc_floats, (size_t)n.n_param_floats);
      }
    }
    if (side && dec_tc) {
      TcDecoderArgs a{};  // fp16 copy of the last decoder weight (joined before the decoder starts)
      a.W = params + n.dec[NB - 1].W; a.B = B; a.G = G; a.h = n.dec[NB - 1].din; a.scratch = c.ws + p.tc;
      dec_w_prepared = tc_decoder_prepare(a, ss);
    }
  };
  // ... and the cleared gradient buffers (same place: measured better than behind the sampling, where the memsets sat in
  // front of the dose-response head on the side stream)
  auto clear_grads = [&]() {
    if (!bwd) return;
    cudaStream_t ss = st;
    if (side) { side->fork(st); ss = side->side; }
    cudaMemsetAsync(grads, 0, sizeof(float) * (size_t)n.n_param_floats, ss);
    // d loss / d (last decoder activation): split-K accumulation target of the fused decoder backward
    if (dec_tc) cudaMemsetAsync(c.at<float>(p.Gb[1][NB - 1]), 0, sizeof(float) * (size_t)B * n.dec[NB - 1].din, ss);
  };
  bool enc_sd = enc_sd0;
  // ---- guide: encoder + reparameterised draws (nbvae.py:81-90) --------------------------------
  for (int k = 0; k < NB; ++k) {
    enc_sd = block_forward(c, 0, k, x, true, enc_sd);
    if (k == 0) {
      late_prep();
      clear_grads();
    }
  }
  const float* O = c.at<float>(p.V[0][NB - 1]);
  float* z = c.at<float>(p.z);
  float* l = c.at<float>(p.l);
  sample_forward(O, eps_z, eps_l, B, L, *hp, z, l, out, st);
  prof::mark(1, st);

  if (side) side->join(st);  // decoder weights' tf32 parts / fp16 copy, cleared gradients
  // ---- dose-response head (drnbvae.py:110-116, 131-132): needs only z, runs on the side stream
  // concurrently with the decoder; joined before the sampling backward consumes dz_drs
  DrShape ds{B, L, d.n_drugs, d.n_classes, d.n_doses_total};
  auto dose_response_head = [&]() {
  if (d.n_drugs > 0) {
    cudaStream_t dst = st;
    if (side) { side->fork(st); dst = side->side; }
    dr_project(z, w_prior, ds, c.at<float>(p.u), dst);
    dr_means(c.at<float>(p.u), b_prior, labels, dose_offsets, ds, y, hp->sigma_scale, c.at<float>(p.m),
             c.at<float>(p.gfac), out, status, dst);
    if (bwd) {
      // weight_loc / weight_scale never enter the loss: their gradient stays zero
    }
    dr_prior_factor(w_prior, b_prior, eps_bias, params + n.b_loc, params + n.b_scale, dose_offsets, ds, *hp,
                    bwd ? grads + n.b_loc : nullptr, bwd ? grads + n.b_scale : nullptr, out, dst);
    if (bwd)
      dr_backward(c.at<float>(p.u), b_prior, w_prior, labels, dose_offsets, c.at<float>(p.gfac), ds,
                  c.at<float>(p.du), c.at<float>(p.dz_drs), dst);
    if (side) side->record_aux();  // the sampling backward waits for exactly this point of the side stream
  }
  };
  // forward-only calls and stream-less calls run it here; in training it is deferred to the decoder
  // backward, where it overlaps the bandwidth-bound GEMMs instead of the latency-bound trunk
  static const bool dr_late_env = getenv("DRV_DR_LATE") != nullptr;
  const bool dr_late = dr_late_env && bwd && side != nullptr && dec_tc;
  if (!dr_late) dose_response_head();

  // ---- model: decoder trunk (nbvae.py:70 -> modules.py:93) ------------------------------------
  const Block& last = n.dec[NB - 1];
  ar_tail_floats = last.W;
  const bool use_tc = tc_decoder_supported(d, *hp) && !(hp->flags & 16);
  // drv_step_out.path: bit 0 = tcgen05 decoder kernels, bit 1 = tcgen05 encoder input layer
  const uint32_t path = (use_tc ? 1u : 0u) |
                        ((p.tcb[0][0] && tc_block_forward_supported(B, n.enc[0].din, n.enc[0].dout, *hp) && !(hp->flags & 32)) ? 2u : 0u);
  bool dec_sd = false;
  for (int k = 0; k < NB - 1; ++k) dec_sd = block_forward(c, 1, k, z, true, dec_sd);
  block_forward(c, 1, NB - 1, z, false, dec_sd);  // statistics of the last BatchNorm only
  prof::mark(2, st);

  // MCVNBVAE: per-cell shift of log_r (count parameter only), mcv_nbvae.py:78-79
  const float* rshift = nullptr;
  if (log_split) {
    float* rs = c.at<float>(p.rshift);
    axpby(log_split_c, log_split, 1.f, -1.f, B, rs, st);
    rshift = rs;
  }
  // ---- last Linear + gene softmax + NB likelihood (modules.py:93-96, nbvae.py:72-78) ----------
  float* dlogits = c.at<float>(p.V[1][NB - 1]);
  TcDecoderArgs ta{};
  if (use_tc) {
    ta.U = c.at<float>(p.V[1][NB - 2]); ta.bn = c.bn(1, NB - 1); ta.W = params + last.W; ta.W_hi = c.at<float>(p.p_hi) + last.W; ta.bias = params + last.bias;
    ta.x = x_lik; ta.rshift = rshift; ta.lib = l; ta.B = B; ta.G = G; ta.h = last.din; ta.eps = hp->epsilon; ta.c = hp->scale_factor;
    ta.lse = c.at<float>(p.lse); ta.ll = c.at<float>(p.ll); ta.asum = c.at<float>(p.asum); ta.out = out;
    ta.scratch = c.ws + p.tc; ta.dlogits = dlogits; ta.bwd = bwd; ta.w_prepared = dec_w_prepared;
    ta.dW = bwd ? grads + last.W : nullptr; ta.db = bwd ? grads + last.bias : nullptr;
    ta.dY = c.at<float>(p.Gb[1][NB - 1]);
    tc_decoder_forward(ta, st);
  } else {
    simt_linear_forward(c.at<float>(p.V[1][NB - 2]), B, last.din, 1, c.bn(1, NB - 1), params + last.W,
                        params + last.bias, last.dout, dlogits, st);
    nb_rows(dlogits, x_lik, l, rshift, B, G, hp->epsilon, hp->scale_factor, bwd, c.at<float>(p.lse), c.at<float>(p.ll),
            c.at<float>(p.asum), out, st);
  }
  prof::mark(3, st);

  prof::mark(4, st);

  if (bwd) {
    // ---- backward of the last decoder Linear ---------------------------------------------------
    if (use_tc) {
      ta.side = side;
      if (dr_late) dose_response_head();  // side stream: overlaps phase 3 / the dAct GEMM, ahead of the dW GEMM
      tc_decoder_backward(ta, st);
    } else {
      cudaStream_t ws = st;
      if (side) { side->fork(st); ws = side->side; }
      simt_linear_wgrad(dlogits, B, last.dout, c.at<float>(p.V[1][NB - 2]), last.din, 1, c.bn(1, NB - 1),
                        grads + last.W, grads + last.bias, ws);
      simt_linear_dgrad_gate(dlogits, B, last.dout, params + last.W, last.din, c.at<float>(p.V[1][NB - 2]), 1,
                             c.bn(1, NB - 1), c.at<float>(p.Gb[1][NB - 1]), st);
    }
    prof::mark(5, st);
    if (fused_ar) {
      // The bulk of the gradient bytes - the last decoder Linear (2G x h), the dose-response head - and the loss / status
      // tail slots are final once the weight-gradient GEMM just queued on the side stream has run: write the tail slots,
      // then all-reduce [last decoder weight, tail] on the comm stream.  It runs next to the decoder-trunk backward, whose
      // kernels are small grids, and is over before the full-machine kernels of the encoder backward start.
      finalize_step(out, status, (bn_batches && !(hp->flags & DRV_FLAG_NO_BN_UPDATE)) ? bn_batches : nullptr, n.n_bn,
                    use_tc ? c.at<float>(p.ll) : nullptr, B, hp->scale_factor, st, grads + n.n_param_floats, path);
      side->comm_fork(st);
      // Few CTAs: the slice has ~100 us of small-grid kernels to hide behind, and every CTA it takes slows those down
      // (measured at cfg 3, ms per step: 8 GPUs 64 CTAs 0.733, 48 x 256 threads 0.716, 32 x 256 threads 0.749, 32 CTAs
      // 0.686-0.694; 4 GPUs 32 CTAs 0.674, 32 x 256 threads 0.664; 2 GPUs 64 CTAs 0.649, 32 CTAs 0.641, 32 x 256 threads
      // 0.634, 16 CTAs 0.680 - too slow for the window).
      static const int early_ctas = getenv("DRV_AR_EARLY_CTAS") ? atoi(getenv("DRV_AR_EARLY_CTAS")) : 32;
      static const int early_threads_env = getenv("DRV_AR_EARLY_THREADS") ? atoi(getenv("DRV_AR_EARLY_THREADS")) : 0;
      const int early_threads = early_threads_env ? early_threads_env : (comm_world() <= 4 ? 256 : 512);
      const int rc = comm_allreduce_slice(last.W, n.n_param_floats + DRV_GRAD_TAIL_FLOATS - last.W, early_ctas, side->comm, early_threads);
      if (rc != 0) return rc;
    }
    // ---- decoder trunk backward, sampling backward ---------------------------------------------
    bool prep = block_backward(c, 1, NB - 1, z, nullptr, true);
    for (int k = NB - 2; k >= 0; --k) prep = block_backward(c, 1, k, z, c.at<float>(p.Gb[1][k + 1]), false, prep);
    if (side && d.n_drugs > 0) side->wait_aux(st);  // dz_drs is ready; weight-gradient GEMMs queued behind it are not waited for
    sample_backward(O, eps_z, eps_l, z, l, c.at<float>(p.Gb[1][0]), d.n_drugs > 0 ? c.at<float>(p.dz_drs) : nullptr,
                    c.at<float>(p.asum), B, L, *hp, c.at<float>(p.dO), st);
    prof::mark(6, st);
    if (stop_after_decoder) {
      // decoder / dose-response gradients and the loss are final: hand back to the caller
      if (side) side->join(st);
      finalize_step(out, status, nullptr, 0, use_tc ? c.at<float>(p.ll) : nullptr, B, hp->scale_factor, st,
                    grads + n.n_param_floats, path);  // loss -> tail slot of the gradient buffer (one collective carries both)
      return finish();
    }
    // ---- encoder backward ----------------------------------------------------------------------
    // Data parallel, collective inside the step - opt-in tail split (DRV_AR_TAIL_SPLIT=1): the input layer's WEIGHT gradient
    // (G x h: nearly all of the encoder's gradient bytes) first, on the main stream; [encoder input weight, last decoder
    // weight) is then all-reduced on the comm stream by 256-thread CTAs that fit next to the persistent CTAs of the input
    // BatchNorm's gradient kernel, which follows; only that kernel's 2 G floats are left for the collective at the end.
    // Measured at 2 GPUs (profiles/r02_dp_timeline_2gpu.txt): the mid slice does overlap, but the last collective costs
    // 25 us however small it is (two flag round trips + the wait for the slower rank; 39 us with the 6 MB encoder slice),
    // and the weight-gradient path no longer runs next to the BatchNorm-gradient kernel (+17 us): 0.705 vs 0.650 ms.
    const bool enc0_tc = p.tcb[0][0] && tc_block_backward_supported(B, n.enc[0].din, n.enc[0].dout, *hp) && !(hp->flags & 32);
    static const bool tail_split_on = getenv("DRV_AR_TAIL_SPLIT") != nullptr && atoi(getenv("DRV_AR_TAIL_SPLIT")) == 1;
    const bool tail_split = fused_ar && enc0_tc && NB >= 2 && tail_split_on;
    for (int k = NB - 1, eprep = 0; k >= 0; --k) {
      const float* dVk = k == NB - 1 ? c.at<float>(p.dO) : c.at<float>(p.Gb[0][k + 1]);
      if (k == 0 && tail_split) {
        Ctx cm = c;
        cm.side = nullptr;  // weight gradient on the main stream
        block_backward(cm, 0, 0, x, dVk, false, eprep, 1);
        side->join(st);       // the trunk weight gradients queued on the side stream
        side->comm_fork(st);  // (the comm stream runs its collectives in order: this one behind the decoder slice)
        static const int mid_ctas = getenv("DRV_AR_MID_CTAS") ? atoi(getenv("DRV_AR_MID_CTAS")) : 96;
        const int rc = comm_allreduce_slice(n.enc[0].W, last.W - n.enc[0].W, mid_ctas, side->comm, 256);
        if (rc != 0) return rc;
        block_backward(c, 0, 0, x, dVk, false, true, 2);
        ar_tail_floats = n.enc[0].W;
        break;
      }
      eprep = block_backward(c, 0, k, x, dVk, false, eprep);
    }
    prof::mark(7, st);
  }
  if (fused_ar) {
    side->join(st);
    side->comm_join(st);  // the two collectives share the flag words: strictly one after the other on every rank
    const int rc = comm_allreduce_slice(0, ar_tail_floats, 0, st);
    if (rc != 0) return rc;
    if (prof::enabled && prof::n_steps < prof::MAX_STEPS) ++prof::n_steps;
    return finish();
  }
  // BatchNorm batch counters + status word folded into the result record: one launch.  In a training step everything
  // that feeds the loss and the status was joined before the sampling backward (wait_aux), so the record is finished
  // NEXT TO the weight-gradient GEMMs still running on the side stream instead of behind them.
  if (side && !bwd) side->join(st);
  finalize_step(out, status, (bn_batches && !(hp->flags & DRV_FLAG_NO_BN_UPDATE)) ? bn_batches : nullptr, n.n_bn,
                use_tc ? c.at<float>(p.ll) : nullptr, B, hp->scale_factor, st,
                (bwd && (hp->flags & DRV_FLAG_LOSS_IN_GRADS)) ? grads + n.n_param_floats : nullptr, path);
  if (side && bwd) side->join(st);
  if (prof::enabled && prof::n_steps < prof::MAX_STEPS) {
    if (!bwd) for (int i = 5; i <= DRV_N_SECTIONS; ++i) prof::mark(i, st);
    ++prof::n_steps;
  }
  return finish();
}

// Entry of both step flavours: shapes the tensor-core kernels cannot take are run as the next aligned problem (k_pad.cu).
static int elbo_step_impl(const drv_dims* dims, const drv_hparams* hp, const float* x, const float* x_lik,
                          const float* log_split, const float* log_split_c, const int64_t* labels, const float* y,
                          const float* eps_z, const float* eps_l, const float* w_prior, const float* b_prior,
                          const float* eps_bias, const int32_t* dose_offsets, const float* params, float* grads,
                          float* bn_running, int64_t* bn_batches, drv_step_out* out, void* workspace,
                          size_t workspace_bytes, void* stream) {
  g_launches = 0;
  g_err = cudaSuccess;
  if (!valid(dims) || !hp || !workspace) return DRV_EINVAL;
  if (!pad_wanted(*dims) || (hp->flags & DRV_FLAG_FORCE_SIMT))
    return elbo_step_core(dims, hp, x, x_lik, log_split, log_split_c, labels, y, eps_z, eps_l, w_prior, b_prior, eps_bias,
                          dose_offsets, params, grads, bn_running, bn_batches, out, workspace, workspace_bytes, stream);
  if (!x || !x_lik || !params) return DRV_EINVAL;
  if (hp->flags & DRV_FLAG_FUSED_ALLREDUCE) return DRV_EINVAL;  // the padded gradient buffer is internal: reduce after the step
  const drv_dims dp = padded_dims(*dims);
  const size_t plan_total = make_plan(dp, make_net(dp)).total;
  const PadPlan P = make_pad_plan(*dims, plan_total);
  if (workspace_bytes < P.total) return DRV_ENOSPC;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  float* pp = (float*)(ws + P.off_params);
  float* gp = grads ? (float*)(ws + P.off_grads) : nullptr;
  float* bp = bn_running ? (float*)(ws + P.off_bn) : nullptr;
  float* xp = (float*)(ws + P.off_x);
  float* xp2 = x_lik != x ? (float*)(ws + P.off_x2) : xp;
  const int B = dims->n_cells, G = dims->n_genes, Gp = dp.n_genes;
  const bool part2 = grads && (hp->flags & DRV_FLAG_ENCODER_BACKWARD_ONLY);  // continues part 1: the padded copies are in place
  if (!part2) {
    pad_params(params, pp, P.params, P.np.n_param_floats, st);
    if (bp) pad_params(bn_running, bp, P.bn, P.np.n_bn_floats, st);
    if (Gp != G) {
      pad_rows(x, B, G, Gp, xp, st);
      if (x_lik != x) pad_rows(x_lik, B, G, Gp, xp2, st);
    }
  }
  const float* xin = Gp != G ? xp : x;
  const float* xlin = Gp != G ? xp2 : x_lik;
  const int rc = elbo_step_core(&dp, hp, xin, xlin, log_split, log_split_c, labels, y, eps_z, eps_l, w_prior, b_prior, eps_bias,
                                dose_offsets, pp, gp, bp, bn_batches, out, workspace, plan_total, stream);
  if (rc != 0) return rc;
  const bool tail = grads && (hp->flags & (DRV_FLAG_LOSS_IN_GRADS | DRV_FLAG_STOP_AFTER_DECODER));
  if (grads)
    unpad(gp, grads, P.params, P.nr.n_param_floats, tail ? DRV_GRAD_TAIL_FLOATS : 0, P.np.n_param_floats, P.nr.n_param_floats, st);
  if (bp && !(hp->flags & DRV_FLAG_NO_BN_UPDATE)) unpad(bp, bn_running, P.bn, P.nr.n_bn_floats, 0, 0, 0, st);
  return finish();
}

int drv_elbo_step(const drv_dims* dims, const drv_hparams* hp, const float* x, const int64_t* labels, const float* y,
                  const float* eps_z, const float* eps_l, const float* w_prior, const float* b_prior,
                  const float* eps_bias, const int32_t* dose_offsets, const float* params, float* grads,
                  float* bn_running, int64_t* bn_batches, drv_step_out* out, void* workspace, size_t workspace_bytes,
                  void* stream) {
  return elbo_step_impl(dims, hp, x, x, nullptr, nullptr, labels, y, eps_z, eps_l, w_prior, b_prior, eps_bias,
                        dose_offsets, params, grads, bn_running, bn_batches, out, workspace, workspace_bytes, stream);
}

int drv_elbo_step_mcv(const drv_dims* dims, const drv_hparams* hp, const float* x0, const float* x1,
                      const float* log_split, const float* log_split_complement, const float* eps_z,
                      const float* eps_l, const float* params, float* grads, float* bn_running, int64_t* bn_batches,
                      drv_step_out* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!dims || dims->n_drugs != 0 || !log_split || !log_split_complement) return DRV_EINVAL;
  return elbo_step_impl(dims, hp, x0, x1, log_split, log_split_complement, nullptr, nullptr, eps_z, eps_l, nullptr,
                        nullptr, nullptr, nullptr, params, grads, bn_running, bn_batches, out, workspace,
                        workspace_bytes, stream);
}

int drv_aggmo_step(const drv_dims* dims, const drv_aggmo* hp, float* params, const float* grads, float* momentum,
                   const int32_t* dose_offsets, void* scratch, void* stream) {
  g_launches = 0;
  g_err = cudaSuccess;
  if (!valid(dims) || !hp || !params || !grads || !momentum || !scratch) return DRV_EINVAL;
  if (dims->n_drugs > 0 && !dose_offsets) return DRV_EINVAL;
  if (hp->n_betas < 1 || hp->n_betas > DRV_MAX_BETAS) return DRV_EINVAL;
  Net n = make_net(*dims);
  long long offs[DRV_MAX_TENSORS], lens[DRV_MAX_TENSORS];
  int nt = 0;
  auto add = [&](int64_t off, int64_t len) {
    if (len > 0 && nt < DRV_MAX_TENSORS) { offs[nt] = off; lens[nt] = len; ++nt; }
  };
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < n.n_blocks; ++k) {
      const Block& b = s == 0 ? n.enc[k] : n.dec[k];
      add(b.gamma, b.din);
      add(b.beta, b.din);
      add(b.W, (int64_t)b.dout * b.din);
      add(b.bias, b.dout);
    }
  int rc = aggmo_step(offs, lens, nt, *hp, params, grads, momentum, (long long)n.n_param_floats, (double*)scratch,
                      (cudaStream_t)stream);
  if (rc != 0) return rc;
  if (dims->n_drugs > 0) {
    // per-drug tensors of the dose-response head (each one its own pyro.param in the reference)
    const long long bases[4] = {n.w_loc, n.w_scale, n.b_loc, n.b_scale};
    aggmo_step_dose_response(bases, dims->n_latent, dims->n_drugs, dose_offsets, *hp, params, grads, momentum,
                             (long long)n.n_param_floats, (cudaStream_t)stream);
  }
  return finish();
}

int drv_profile_enable(int on) {
  if (on && !prof::created) {
    for (int i = 0; i < prof::MAX_STEPS; ++i)
      for (int j = 0; j <= DRV_N_SECTIONS; ++j)
        if (cudaEventCreate(&prof::ev[i][j]) != cudaSuccess) return (int)cudaGetLastError();
    for (int i = 0; i < prof::MAX_STEPS; ++i)
      for (int j = 0; j < DRV_N_PROFILED_KERNELS; ++j)
        for (int w = 0; w < 2; ++w)
          if (cudaEventCreate(&prof::kev[i][j][w]) != cudaSuccess) return (int)cudaGetLastError();
    prof::created = true;
  }
  prof::enabled = on != 0;
  prof::n_steps = 0;
  return 0;
}

int drv_profile_read(float* h_ms_per_section, int* h_n_steps) {
  if (!prof::created) return DRV_EINVAL;
  int ns = prof::n_steps;
  for (int j = 0; j < DRV_N_SECTIONS; ++j) h_ms_per_section[j] = 0.f;
  for (int i = 0; i < ns; ++i) {
    cudaError_t e = cudaEventSynchronize(prof::ev[i][DRV_N_SECTIONS]);
    if (e != cudaSuccess) return (int)e;
    for (int j = 0; j < DRV_N_SECTIONS; ++j) {
      float ms = 0.f;
      e = cudaEventElapsedTime(&ms, prof::ev[i][j], prof::ev[i][j + 1]);
      if (e != cudaSuccess) return (int)e;
      h_ms_per_section[j] += ms / ns;
    }
  }
  if (h_n_steps) *h_n_steps = ns;
  return 0;
}

int drv_profile_read_kernels(float* h_ms_per_kernel) {
  if (!prof::created) return DRV_EINVAL;
  int ns = prof::n_steps;
  for (int j = 0; j < DRV_N_PROFILED_KERNELS; ++j) h_ms_per_kernel[j] = 0.f;
  for (int i = 0; i < ns; ++i)
    for (int j = 0; j < DRV_N_PROFILED_KERNELS; ++j) {
      float ms = 0.f;
      cudaError_t e = cudaEventSynchronize(prof::kev[i][j][1]);
      if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, prof::kev[i][j][0], prof::kev[i][j][1]);
      if (e != cudaSuccess) { cudaGetLastError(); ms = 0.f; }  // kernel not launched on this path
      h_ms_per_kernel[j] += ms / ns;
    }
  return 0;
}

void drv_internal_kmark(int k, int which, void* stream) { prof::kmark(k, which, (cudaStream_t)stream); }

int drv_encoder_forward(const drv_dims* dims, const drv_hparams* hp, const float* x, const float* params,
                        float* bn_running, int64_t* bn_batches, float* z_loc, float* z_scale, float* l_loc,
                        float* l_scale, void* workspace, size_t workspace_bytes, void* stream) {
  g_launches = 0;
  g_err = cudaSuccess;
  if (!valid(dims) || !hp || !x || !params || !workspace) return DRV_EINVAL;
  Net n = make_net(*dims);
  Plan p = make_plan(*dims, n);
  if (workspace_bytes < p.total) return DRV_ENOSPC;
  cudaStream_t st = (cudaStream_t)stream;
  Ctx c{*dims, *hp, n, p, (char*)workspace, params, nullptr, bn_running, st, nullptr};
  cudaMemsetAsync(c.ws + p.zero_begin, 0, p.zero_bytes, st);
  if (!(hp->flags & DRV_FLAG_FORCE_SIMT))
    split_params_tf32(params, (size_t)n.n_param_floats, c.at<float>(p.p_hi), c.at<float>(p.p_lo), st);
  for (int k = 0, sd = 0; k < n.n_blocks; ++k) sd = block_forward(c, 0, k, x, true, sd);
  encoder_outputs(c.at<float>(p.V[0][n.n_blocks - 1]), dims->n_cells, dims->n_latent, z_loc, z_scale, l_loc, l_scale, st);
  if (bn_batches && !(hp->flags & DRV_FLAG_NO_BN_UPDATE)) bump_counters(bn_batches, n.n_blocks, st);
  return finish();
}

int drv_decoder_forward(const drv_dims* dims, const drv_hparams* hp, const float* z, const float* library,
                        const float* params, float* bn_running, int64_t* bn_batches, float* log_r, float* logit,
                        void* workspace, size_t workspace_bytes, void* stream) {
  g_launches = 0;
  g_err = cudaSuccess;
  if (!valid(dims) || !hp || !z || !library || !params || !workspace) return DRV_EINVAL;
  Net n = make_net(*dims);
  Plan p = make_plan(*dims, n);
  if (workspace_bytes < p.total) return DRV_ENOSPC;
  cudaStream_t st = (cudaStream_t)stream;
  Ctx c{*dims, *hp, n, p, (char*)workspace, params, nullptr, bn_running, st, nullptr};
  cudaMemsetAsync(c.ws + p.zero_begin, 0, p.zero_bytes, st);
  if (!(hp->flags & DRV_FLAG_FORCE_SIMT))
    split_params_tf32(params, (size_t)n.n_param_floats, c.at<float>(p.p_hi), c.at<float>(p.p_lo), st);
  for (int k = 0, sd = 0; k < n.n_blocks; ++k) sd = block_forward(c, 1, k, z, true, sd);
  nb_decoder_outputs(c.at<float>(p.V[1][n.n_blocks - 1]), library, dims->n_cells, dims->n_genes, log_r, logit, st);
  if (bn_batches && !(hp->flags & DRV_FLAG_NO_BN_UPDATE)) bump_counters(bn_batches + n.n_blocks, n.n_blocks, st);
  return finish();
}

int drv_logit_mean_sigmoid(const drv_dims* dims, const float* z, const float* w, const float* b, const int64_t* labels,
                           const int32_t* dose_offsets, float* out, uint32_t* status, void* workspace,
                           size_t workspace_bytes, void* stream) {
  g_launches = 0;
  g_err = cudaSuccess;
  if (!valid(dims) || dims->n_drugs < 1 || !z || !w || !b || !labels || !dose_offsets || !out || !status || !workspace)
    return DRV_EINVAL;
  if (workspace_bytes < sizeof(float) * (size_t)dims->n_cells * dims->n_drugs) return DRV_ENOSPC;
  cudaStream_t st = (cudaStream_t)stream;
  DrShape ds{dims->n_cells, dims->n_latent, dims->n_drugs, dims->n_classes, dims->n_doses_total};
  cudaMemsetAsync(status, 0, sizeof(uint32_t), st);
  dr_project(z, w, ds, (float*)workspace, st);
  dr_means((const float*)workspace, b, labels, dose_offsets, ds, nullptr, 1.f, out, nullptr, nullptr, status, st);
  return finish();
}

int drv_fill_normal(float* out, int64_t n, uint64_t seed, uint64_t offset, const uint64_t* step_counter, uint64_t stride,
                    void* stream) {
  g_launches = 0;
  g_err = cudaSuccess;
  if (!out || n < 0) return DRV_EINVAL;
  fill_normal(out, n, seed, offset, (const unsigned long long*)step_counter, stride, (cudaStream_t)stream);
  return finish();
}

int drv_draw_step_noise(const drv_dims* dims, const drv_hparams* hp, uint64_t seed, uint64_t offset, uint64_t* state,
                        uint64_t stride, float* eps_z, float* eps_l, float* w_prior, float* b_prior, float* eps_bias,
                        void* stream) {
  g_launches = 0;
  g_err = cudaSuccess;
  if (!dims || !hp || !state || !eps_z || !eps_l) return DRV_EINVAL;
  const bool dr = dims->n_drugs > 0;
  if (dr && (!w_prior || !b_prior || !eps_bias)) return DRV_EINVAL;
  float* outs[5] = {eps_z, eps_l, dr ? w_prior : nullptr, dr ? b_prior : nullptr, dr ? eps_bias : nullptr};
  int64_t ns[5] = {(int64_t)dims->n_cells * dims->n_latent, dims->n_cells, (int64_t)dims->n_drugs * dims->n_latent,
                   dims->n_doses_total, dims->n_doses_total};
  int lap[5] = {0, 0, 1, 0, 0};
  float sc[5] = {1.f, 1.f, hp->lam_scale, hp->bias_scale, 1.f};
  draw_step_noise(outs, ns, lap, sc, seed, offset, (unsigned long long*)state, stride, (cudaStream_t)stream);
  return finish();
}

int drv_fill_laplace(float* out, int64_t n, uint64_t seed, uint64_t offset, const uint64_t* step_counter, uint64_t stride,
                     void* stream) {
  g_launches = 0;
  g_err = cudaSuccess;
  if (!out || n < 0) return DRV_EINVAL;
  fill_laplace(out, n, seed, offset, (const unsigned long long*)step_counter, stride, (cudaStream_t)stream);
  return finish();
}

int drv_counter_add(uint64_t* counter, uint64_t v, void* stream) {
  g_launches = 0;
  g_err = cudaSuccess;
  if (!counter) return DRV_EINVAL;
  counter_add((unsigned long long*)counter, v, (cudaStream_t)stream);
  return finish();
}

}  // extern "C"
