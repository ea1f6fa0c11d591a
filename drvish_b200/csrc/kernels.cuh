// Host-callable launchers of the drvish_b200 kernels (internal; the public surface is
// include/drvish_b200.h).
#pragma once
#include "common.cuh"

namespace drv {

// ---- k_stats.cu: BatchNorm statistics and their backward ------------------------------------
// acc: [2][D] doubles and *ticket, all zero on entry.  Sum / sum of squares of (sqrt?)U over the
// batch; the last CTA turns them into mean / rstd and updates the running buffers like torch
// BatchNorm1d in training mode.
void colstats(const float* U, int B, int D, bool sqrt_in, double* acc, unsigned int* ticket, float eps, float momentum,
              float* mean, float* rstd, float* running_mean, float* running_var, cudaStream_t st);
// acc[0][i] += sum_b dY, acc[1][i] += sum_b dY * xhat   (gate: dY is ungated, apply [bn(U) > 0] here)
void bn_bwd_reduce(const float* dY, const float* U, int B, int D, BnParams bn, double* acc, bool gate, cudaStream_t st);
// dU = gamma rstd (dY - s1/B - xhat s2/B) in place over dY; also dgamma = s2, dbeta = s1.
void bn_bwd_apply(float* dY_dU, const float* U, int B, int D, BnParams bn, const double* acc, float* dgamma,
                  float* dbeta, bool gate, cudaStream_t st);
// same, but dU is written tf32-rounded to dU_r (not in place) and its column sums are added to db_next
void bn_bwd_apply_round(const float* dY, const float* U, int B, int D, BnParams bn, const double* acc, float* dgamma,
                        float* dbeta, bool gate, float* dU_r, float* db_next, cudaStream_t st);
void bn_grad_finalize(const double* acc, int D, float* dgamma, float* dbeta, cudaStream_t st);
void bump_counters(int64_t* counters, int n, cudaStream_t st);

// ---- k_gemm_simt.cu ---------------------------------------------------------------------------
void simt_linear_forward(const float* U, int B, int din, int pro_mode, BnParams bn, const float* W, const float* bias,
                         int dout, float* V, cudaStream_t st);
// narrow input (din <= 64), BatchNorm + ReLU prologue, also the batch statistics of V for the next block;
// false (nothing launched) when the shape does not qualify
bool simt_narrow_linear_forward_stats(const float* U, int B, int din, BnParams bn, const float* W, const float* bias,
                                      int dout, float* V, const ColStatsOut& ns, cudaStream_t st);
void simt_linear_wgrad(const float* dV, int B, int dout, const float* U, int din, int pro_mode, BnParams bn, float* dW,
                       float* db, cudaStream_t st);
void simt_linear_dgrad_gate(const float* dV, int B, int dout, const float* W, int din, const float* U, int pro_mode,
                            BnParams bn, float* dY, cudaStream_t st);
void simt_linear_dgrad_bngrad(const float* dV, int B, int dout, const float* W, int din, const float* U, int pro_mode,
                              BnParams bn, double* acc, cudaStream_t st);

// ---- k_nb.cu: gene-softmax + library scaling + NB log-likelihood on materialised logits -------
// logits [B][2G] = [scale | log_r]; bwd: overwritten with d loss / d logits.
// rshift (may be null): per-cell shift of log_r for the count parameter (MCVNBVAE, mcv_nbvae.py:78-79)
void nb_rows(float* logits, const float* x, const float* lib, const float* rshift, int B, int G, float eps,
             float scale_factor, bool bwd, float* lse, float* ll, float* asum, drv_step_out* out, cudaStream_t st);
// NBDecoder.forward outputs from materialised logits
void nb_decoder_outputs(const float* logits, const float* lib, int B, int G, float* log_r, float* logit, cudaStream_t st);

// ---- k_sample.cu --------------------------------------------------------------------------------
void sample_forward(const float* O, const float* eps_z, const float* eps_l, int B, int L, const drv_hparams& hp,
                    float* z, float* l, drv_step_out* out, cudaStream_t st);
void sample_backward(const float* O, const float* eps_z, const float* eps_l, const float* z, const float* l,
                     const float* dz_dec, const float* dz_drs, const float* asum, int B, int L, const drv_hparams& hp,
                     float* dO, cudaStream_t st);
void encoder_outputs(const float* O, int B, int L, float* z_loc, float* z_scale, float* l_loc, float* l_scale,
                     cudaStream_t st);
void fill_normal(float* out, int64_t n, uint64_t seed, uint64_t offset, const unsigned long long* ctr, uint64_t stride,
                 cudaStream_t st);
void fill_laplace(float* out, int64_t n, uint64_t seed, uint64_t offset, const unsigned long long* ctr, uint64_t stride,
                  cudaStream_t st);
void counter_add(unsigned long long* ctr, unsigned long long v, cudaStream_t st);
void axpby(const float* a, const float* b, float ca, float cb, int n, float* out, cudaStream_t st);  // out = ca a + cb b
// all five noise buffers of one step in one launch; state = {step counter, ticket}
void draw_step_noise(float* const* outs, const int64_t* ns, const int* laplace, const float* scales, uint64_t seed,
                     uint64_t offset, unsigned long long* state, uint64_t stride, cudaStream_t st);

// ---- k_dr.cu: dose-response head ----------------------------------------------------------------
struct DrShape {
  int B, L, D, K, T;  // T = total doses
};
// u[b][i] = z[b,:] . w[i,:]
void dr_project(const float* z, const float* w, DrShape s, float* u, cudaStream_t st);
// per (drug, dose) slot: class means of sigmoid(u + bias), clamp+logit -> m[i][k][c] (drug-major);
// if y != NULL also adds -sum log N(y | m, sigma) to out->loss and writes gfac[k][slot] = d loss / d resp / n_k.
void dr_means(const float* u, const float* bias, const int64_t* labels, const int32_t* dose_off, DrShape s,
              const float* y, float sigma, float* m, float* gfac, drv_step_out* out, uint32_t* status, cudaStream_t st);
// du[b][i] then dz[b][d] = sum_i du[b][i] w[i][d]
void dr_backward(const float* u, const float* bias, const float* w, const int64_t* labels, const int32_t* dose_off,
                 const float* gfac, DrShape s, float* du, float* dz, cudaStream_t st);
// prior log-probs of the prior draws + guide-side monotonic factor and its gradients
void dr_prior_factor(const float* w_prior, const float* b_prior, const float* eps_bias, const float* bias_loc,
                     const float* bias_scale_unc, const int32_t* dose_off, DrShape s, const drv_hparams& hp,
                     float* d_bias_loc, float* d_bias_scale_unc, drv_step_out* out, cudaStream_t st);

// counters[i] += 1 (BatchNorm num_batches_tracked; may be null), out->status = *status | NaN flag
// ... and out->loss -= scale * sum_b ll[b] (ll may be null); out->path = path (which kernels ran)
void finalize_step(drv_step_out* out, const uint32_t* status, int64_t* counters, int n, const float* ll, int B,
                   float scale, cudaStream_t st, float* loss_slot = nullptr, uint32_t path = 0);

// ---- k_pad.cu: shape padding for the tensor-core path (see the file header) ------------------------
// One tensor of a flat buffer: real layout [n_blocks x src_rows][src_cols] at src_off, padded layout
// [dst_rows][dst_cols] at dst_off; block b of the real rows starts at padded row b * blk_stride; pad[b] fills the
// padded entries of block b (and all pad columns).
struct PadSeg {
  long long src_off, dst_off;
  int src_rows, src_cols, dst_rows, dst_cols, n_blocks, blk_stride;
  float pad[2];
};
struct PadSegs {
  int n;
  PadSeg seg[80];  // segments ordered by offset in both layouts (<= 2 x 9 blocks x 4 tensors + 4)
};
void pad_params(const float* src, float* dst, const PadSegs& s, long long n_dst, cudaStream_t st);
void unpad(const float* padded, float* real, const PadSegs& s, long long n_real, int n_tail, long long tail_padded,
           long long tail_real, cudaStream_t st);
void pad_rows(const float* x, int B, int G, int Gp, float* xp, cudaStream_t st);

// ---- k_comm.cu: the calling thread's peer-mapped gradient buffer (drv_comm_bind) ----------------------------
bool comm_bound(const float* grads, long long n_floats_needed);
// all-reduce of floats [off, off + n) of the bound buffer on `st`; max_ctas > 0 caps the grid
int comm_allreduce_slice(long long off_floats, long long n_floats, int max_ctas, cudaStream_t st, int threads = 512);
int comm_world();  // ranks of the bound peer all-reduce (1 when none is bound)

// ---- k_optim.cu ------------------------------------------------------------------------------------
int aggmo_step(const long long* offs, const long long* lens, int n_tensors, const drv_aggmo& hp, float* params,
               const float* grads, float* momentum, long long mom_stride, double* norms, cudaStream_t st);

void aggmo_step_dose_response(const long long* bases, int L, int n_drugs, const int32_t* dose_off, const drv_aggmo& hp,
                              float* params, const float* grads, float* momentum, long long mom_stride, cudaStream_t st);

}  // namespace drv
