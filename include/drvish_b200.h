/*
 * drvish_b200 - C ABI of the B200-native drvish ELBO step.
 *
 * The reference (czbiohub/drvish) is pure Python and has no FFI seam of its own; the seam is
 * the Python object protocol consumed by pyro.infer.SVI.  Every entry point
 * below therefore cites the reference Python it replaces:
 *
 *   drv_elbo_step      svi.step(*xs) / svi.evaluate_loss(*xs) as driven by
 *                      src/drvish/train/__init__.py:41-46, i.e. one trace of
 *                      NBVAE.guide + NBVAE.model   (src/drvish/models/nbvae.py:57-90)   or
 *                      DRNBVAE.guide + DRNBVAE.model (src/drvish/models/drnbvae.py:87-132)
 *                      plus Pyro's Trace_ELBO loss_and_grads (loss.backward()).
 *   drv_encoder_forward   Encoder.forward   (src/drvish/models/modules.py:46-60)
 *   drv_decoder_forward   NBDecoder.forward (src/drvish/models/modules.py:77-98)
 *   drv_logit_mean_sigmoid LinearMultiBias.logit_mean_sigmoid (modules.py:159-180)
 *   drv_fill_normal / drv_fill_laplace   Normal.rsample / Laplace.rsample noise draws
 *                      (nbvae.py:89-90, modules.py:130-135,152-157) in production mode
 *
 * Conventions: all pointers are DEVICE pointers unless the name starts with h_; tensors are
 * fp32, row-major, contiguous; the caller owns every buffer (nothing is allocated or freed
 * here); all work is enqueued on `stream` with no host synchronisation; the return value is 0
 * or the cudaError_t of the failed launch / a negative DRV_E* code for invalid arguments.
 */
#ifndef DRVISH_B200_H
#define DRVISH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRV_MAX_LAYERS 8

#define DRV_EINVAL (-1)  /* bad dims / null pointer */
#define DRV_ENOSPC (-2)  /* workspace too small */

/* status bits written to drv_step_out.status by the kernels */
#define DRV_STATUS_NAN_LOSS 1u      /* mirrors pyro's warn_if_nan(loss) */
#define DRV_STATUS_MISSING_CLASS 2u /* a class 0..K-1 has no cell in the batch: the reference's
                                       scatter_add_ raises "index out of bounds" (modules.py:175-179) */
#define DRV_STATUS_BAD_LABEL 4u     /* label outside [0, K) */

/* Problem shape.  Encoder dims are [n_genes, layers..., 2*n_latent+2]; the decoder gets the
 * reversed hidden layers and ends in 2*n_genes (nbvae.py:41-42, modules.py:44,75). */
typedef struct drv_dims {
  int32_t n_cells;   /* B: cells in this minibatch */
  int32_t n_genes;   /* G */
  int32_t n_latent;  /* L */
  int32_t n_layers;  /* number of hidden layers, 1..DRV_MAX_LAYERS */
  int32_t layers[DRV_MAX_LAYERS];
  int32_t n_drugs;   /* D, 0 for NBVAE */
  int32_t n_classes; /* K */
  int32_t n_doses_total; /* sum_i C_i */
} drv_dims;

typedef struct drv_hparams {
  float scale_factor;  /* poutine.scale (nbvae.py:61) */
  float epsilon;       /* added to exp(log_r) (nbvae.py:75) */
  float lib_loc;       /* library prior (nbvae.py:46-47,66-68) */
  float lib_scale;
  float lam_scale;     /* Laplace prior scale of weight_p (modules.py:130-132) */
  float bias_scale;    /* Normal prior scale of bias_p (modules.py:133-135) */
  float sigma_scale;   /* drs_i observation noise (drnbvae.py:114) */
  float alpha;         /* monotonic-bias factor weight (modules.py:193-196) */
  float bn_eps;        /* BatchNorm1d eps (1e-5) */
  float bn_momentum;   /* BatchNorm1d momentum (0.1) */
  int32_t include_guide_factor; /* 1: add the guide-side monotonic factor (pathwise gradient) */
  int32_t flags;       /* DRV_FLAG_* */
} drv_hparams;

#define DRV_FLAG_FORCE_SIMT 1   /* never take the tcgen05/TMA kernels (debug / A-B parity) */
#define DRV_FLAG_NO_BN_UPDATE 2 /* do not touch running_mean/var (gradient checks) */
/* bits 4..64 are debug switches (force single kernels onto the CUDA-core path, no side stream) */
#define DRV_FLAG_STOP_AFTER_DECODER 128     /* part 1 of 2: return once the decoder, dose-response and latent
                                               gradients and the loss are final (data-parallel overlap); the
                                               loss is also written as a float to grads[total floats of
                                               drv_param_offsets], so grads needs that one extra slot */
#define DRV_FLAG_LOSS_IN_GRADS 2048         /* whole step in one call, loss also written to grads[total floats]
                                               (data parallel with ONE all-reduce after the step) */
/* Tail slots of the gradient buffer written with DRV_FLAG_LOSS_IN_GRADS / DRV_FLAG_STOP_AFTER_DECODER (grads needs
 * DRV_GRAD_TAIL_FLOATS extra floats): [0] loss, [1] 1.0 if DRV_STATUS_NAN_LOSS else 0, [2] MISSING_CLASS, [3] BAD_LABEL.
 * After a SUM all-reduce every rank holds the summed loss and the number of ranks that raised each status bit. */
#define DRV_GRAD_TAIL_FLOATS 4
#define DRV_FLAG_FUSED_ALLREDUCE 4096        /* the step performs the data-parallel gradient all-reduce itself (implies
                                               LOSS_IN_GRADS): `grads` must be the data pointer of the peer-mapped buffer
                                               bound with drv_comm_bind.  The decoder / dose-response slice (final before
                                               the encoder backward starts) is reduced on an internal side stream by a
                                               few CTAs NEXT TO the encoder backward; the encoder slice follows at the end
                                               of the step.  DRV_EINVAL when nothing is bound or the shape is padded */
#define DRV_FLAG_ENCODER_BACKWARD_ONLY 256  /* part 2 of 2: encoder backward of the step started by part 1
                                               (same arguments, same workspace) */

/* Results of one step: one 16-byte device record, read back with a single D2H copy. */
typedef struct drv_step_out {
  double loss;      /* -ELBO of the minibatch (what SVI.step returns) */
  uint32_t status;  /* DRV_STATUS_* */
  uint32_t path;    /* bit 0: tcgen05 decoder kernels were used; bit 1: tcgen05 encoder kernels */
} drv_step_out;

/* Number of tensors / total floats of the flat parameter buffer.  offsets[i] is the float
 * offset of the i-th tensor in state_dict order:
 *   encoder.fc.{0,2,...}: bn.weight, bn.bias, linear.weight, linear.bias per block; then
 *   decoder.fc likewise; then (D > 0) weight_loc[D][L], weight_scale_unconstrained[D][L],
 *   bias_loc[sum C], bias_scale_unconstrained[sum C].
 * Returns the tensor count (<= max_n) or a negative error. */
int drv_param_offsets(const drv_dims* dims, int64_t* offsets, int64_t* sizes, int max_n, int64_t* total_floats);

/* Same for BatchNorm buffers: running_mean, running_var per BatchNorm in module order. */
int drv_bn_offsets(const drv_dims* dims, int64_t* offsets, int64_t* sizes, int max_n, int64_t* total_floats);

size_t drv_workspace_bytes(const drv_dims* dims);

/* One SVI step.  grads == NULL -> forward only (svi.evaluate_loss; BatchNorm still uses batch
 * statistics and still updates its running buffers).
 *   x        [B][G] counts          labels [B] int64 (D>0)      y [sum_i K*C_i] drug-major (k,c)
 *   eps_z    [B][L], eps_l [B]      standard-normal noise of the two rsample sites
 *   w_prior  [D][L], b_prior [sum C] draws of i.weight_p / i.bias_p;  eps_bias [sum C]
 *   dose_offsets [D+1] int32 prefix sums of C_i
 *   params / grads  flat buffers laid out by drv_param_offsets (grads are OVERWRITTEN)
 *   bn_running      flat buffer laid out by drv_bn_offsets; bn_batches [n_bn] int64 counters */
int drv_elbo_step(const drv_dims* dims, const drv_hparams* hp,
                  const float* x, const int64_t* labels, const float* y,
                  const float* eps_z, const float* eps_l,
                  const float* w_prior, const float* b_prior, const float* eps_bias,
                  const int32_t* dose_offsets,
                  const float* params, float* grads,
                  float* bn_running, int64_t* bn_batches,
                  drv_step_out* out,
                  void* workspace, size_t workspace_bytes, void* stream);

/* The same step for the molecular-cross-validation variant MCVNBVAE (reference
 * src/drvish/models/mcv_nbvae.py:57-105): the guide encodes x0, the model scores x1,
 * and log_r is shifted by log_split_complement - log_split (per cell, [B] each) for the NB count
 * parameter only - the logits the decoder returned are NOT recomputed (:78-79).  dims->n_drugs must
 * be 0.  Same kernels: the shift rides in the FFMA that feeds exp2 in the NB epilogue. */
int drv_elbo_step_mcv(const drv_dims* dims, const drv_hparams* hp, const float* x0, const float* x1,
                      const float* log_split, const float* log_split_complement, const float* eps_z,
                      const float* eps_l, const float* params, float* grads, float* bn_running, int64_t* bn_batches,
                      drv_step_out* out, void* workspace, size_t workspace_bytes, void* stream);

/* Encoder.forward (modules.py:46-60): writes z_loc [B][L], z_scale [B][L], l_loc [B], l_scale [B]. */
int drv_encoder_forward(const drv_dims* dims, const drv_hparams* hp, const float* x,
                        const float* params, float* bn_running, int64_t* bn_batches,
                        float* z_loc, float* z_scale, float* l_loc, float* l_scale,
                        void* workspace, size_t workspace_bytes, void* stream);

/* NBDecoder.forward (modules.py:77-98): writes log_r [B][G] and logit [B][G]. */
int drv_decoder_forward(const drv_dims* dims, const drv_hparams* hp, const float* z, const float* library,
                        const float* params, float* bn_running, int64_t* bn_batches,
                        float* log_r, float* logit,
                        void* workspace, size_t workspace_bytes, void* stream);

/* LinearMultiBias.logit_mean_sigmoid for all drugs at once (modules.py:159-180):
 * out [sum_i K*C_i] drug-major (k,c).  status gets DRV_STATUS_MISSING_CLASS / BAD_LABEL. */
int drv_logit_mean_sigmoid(const drv_dims* dims, const float* z, const float* w, const float* b,
                           const int64_t* labels, const int32_t* dose_offsets,
                           float* out, uint32_t* status,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Counter-based (Philox4x32-10) noise: out[i] ~ N(0,1) / Laplace(0,1), keyed by
 * (seed, offset + step_counter[0]*stride + i).  step_counter (device, may be NULL) lets a captured
 * CUDA graph draw fresh noise on every replay; drv_counter_add advances it on the stream. */
int drv_fill_normal(float* out, int64_t n, uint64_t seed, uint64_t offset, const uint64_t* step_counter, uint64_t stride,
                    void* stream);
int drv_fill_laplace(float* out, int64_t n, uint64_t seed, uint64_t offset, const uint64_t* step_counter, uint64_t stride,
                     void* stream);
int drv_counter_add(uint64_t* counter, uint64_t v, void* stream);

/* All noise draws of one ELBO step in a single launch, in the reference's draw order (latent,
 * library - nbvae.py:89-90; then weight_p ~ Laplace(0, lam_scale), bias_p ~ Normal(0, bias_scale),
 * and the unit-normal of the guide's bias - modules.py:130-135,152-157).  Buffer k starts at Philox
 * offset `offset` + sum of the 4-rounded sizes before it, so the values equal those of
 * drv_fill_normal / drv_fill_laplace at the same offsets (times the prior scale).  state is a
 * device uint64[2] = {step counter, 0}: the step counter is advanced by one when the launch
 * completes.  The three dose-response buffers may be NULL when dims->n_drugs == 0. */
int drv_draw_step_noise(const drv_dims* dims, const drv_hparams* hp, uint64_t seed, uint64_t offset, uint64_t* state,
                        uint64_t stride, float* eps_z, float* eps_l, float* w_prior, float* b_prior, float* eps_bias,
                        void* stream);

/* Lossless 16-bit staging of count tiles for the host->device copy (the reference ships dense fp32
 * minibatches: train/__init__.py:42-43).  drv_host_pack_counts_u16 (HOST pointers, n_threads worker
 * threads) writes h_dst[i] = (uint16)h_src[i] and returns n when every entry is an integer in
 * [0, 65535], else the index of the first entry that is not (the tile must then travel as fp32);
 * -1 for invalid arguments.  drv_unpack_counts_u16 widens a device uint16 tile back to fp32
 * (bit-exact inverse); src and dst must be 16-byte aligned. */
int64_t drv_host_pack_counts_u16(const float* h_src, uint16_t* h_dst, int64_t n, int n_threads);
int drv_unpack_counts_u16(const uint16_t* src, float* dst, int64_t n, void* stream);

/* Lossless 8-bit staging (the default transport of integer count tiles): h_dst[i] = (uint8)h_src[i] where
 * h_src[i] is an integer 0..254, otherwise 255 and (i, h_src[i]) is appended to the escape list - so ANY fp32 tile
 * is reproduced bit-exactly, at one byte per entry over the bus when counts are small.  HOST pointers; h_dst
 * 64-byte aligned; n < 2^32.  Runs on a persistent pool of n_threads workers (AVX-512 / AVX2 / scalar chosen at
 * run time: drv_host_pack_isa).  Returns the number of escapes, -1 for invalid arguments, -2 when there are more
 * than max_escapes.  drv_unpack_counts_u8 widens the device copy to fp32 and patches the escapes (DEVICE pointers,
 * src and dst 16-byte aligned). */
int64_t drv_host_pack_counts_u8(const float* h_src, uint8_t* h_dst, int64_t n, uint32_t* h_esc_idx, float* h_esc_val,
                                int64_t max_escapes, int n_threads);
const char* drv_host_pack_isa(void);
int drv_unpack_counts_u8(const uint8_t* src, const uint32_t* esc_idx, const float* esc_val, int n_escapes, float* dst,
                         int64_t n, void* stream);

/* Fused optimizer step over the flat buffers: AggMo.step (reference
 * src/drvish/train/aggmo.py:42-79) applied to every parameter tensor, preceded by Pyro's per-tensor
 * gradient clipping (PyroOptim clip_args {"clip_norm": c} = torch.nn.utils.clip_grad_norm_ on each
 * parameter separately; clip_norm <= 0 disables it).  lr is the EFFECTIVE rate, i.e. already divided
 * by n_betas as the reference does at construction (aggmo.py:22).  momentum: n_betas consecutive
 * copies of the flat layout (zero-initialised by the caller before the first step); scratch: at
 * least DRV_MAX_TENSORS doubles; dose_offsets: device int32[n_drugs + 1] as for drv_elbo_step (NULL
 * without drugs) - every drug's weight_loc / weight_scale / bias_loc / bias_scale is its own tensor.
 * Two launches for the networks (one without clipping) + one for the dose-response head; gradients
 * are not modified. */
#define DRV_MAX_BETAS 4
#define DRV_MAX_TENSORS 96
typedef struct drv_aggmo {
  float lr;
  float weight_decay;
  float clip_norm;
  int32_t nesterov;
  int32_t n_betas;
  float betas[DRV_MAX_BETAS];
} drv_aggmo;
int drv_aggmo_step(const drv_dims* dims, const drv_aggmo* hp, float* params, const float* grads, float* momentum,
                   const int32_t* dose_offsets, void* scratch, void* stream);

/* Dense [n_rows][n_genes] fp32 minibatch tile from a DEVICE-resident CSR count matrix (SURVEY 8f row
 * 2; reference CupySparseDataLoader, src/drvish/data/cupy.py:36-43: `t[indices].todense()`).  indptr:
 * int32 or int64 (indptr_is_int64) [n_total_rows + 1]; indices int32; data fp32; rows: int64 [n_rows]
 * row ids of the minibatch (NULL = rows 0..n_rows-1).  Duplicate column entries keep the last one
 * written (canonical CSR has none). */
int drv_csr_gather_rows(const void* indptr, int indptr_is_int64, const int32_t* indices, const float* data,
                        const int64_t* rows, int n_rows, int n_genes, float* out, void* stream);

/* Molecular-cross-validation split of a UMI count tile (reference
 * MCVDataLoader.split_molecules, src/drvish/train/mcv_train.py:20-40), one launch per minibatch:
 *   x_data ~ Binomial(umis, data_split - overlap_factor)
 *   y_data ~ Binomial(umis - x_data, (1 - data_split) / (1 - data_split + overlap_factor))
 *   overlap = umis - x_data - y_data
 *   x0 = x_data + overlap, x1 = y_data + overlap, log_split = log data_split,
 *   log_split_complement = log data_split_complement
 * umis, x0, x1: [n_rows][n_genes] fp32 (non-negative integer counts); data_split,
 * data_split_complement, overlap_factor, log_split, log_split_complement: one float per row.  Draws
 * are Philox4x32-10 keyed by (seed, call_index, element, draw): the same arguments reproduce the same
 * split whatever the launch shape; advance call_index once per minibatch. */
int drv_split_molecules(const float* umis, const float* data_split, const float* data_split_complement,
                        const float* overlap_factor, int n_rows, int n_genes, uint64_t seed, uint64_t call_index,
                        float* x0, float* x1, float* log_split, float* log_split_complement, void* stream);

/* Gradient all-reduce over NVLink peer memory ("a single gradient all-reduce per step"; the
 * reference itself is single-device, src/drvish/train/__init__.py:42-46).  Every rank of the node allocates its flat
 * gradient buffer with drv_comm_alloc (cudaMalloc + a 64-byte CUDA IPC handle, HOST pointer h_handle64),
 * exchanges the handles out of band (torch.distributed), maps the others with drv_comm_open, and passes
 * the `world` base pointers (HOST array indexed by rank; its own entry is its allocation) to
 * drv_peer_allreduce: one kernel that sums the n_floats floats found drv_comm_header_bytes() behind every
 * base (rank order, bit-identical result on all ranks) in place.  world must be 2, 4 or 8; all ranks must
 * call it the same number of times.  Data pointer of a region = base + drv_comm_header_bytes(). */
size_t drv_comm_header_bytes(void);
int drv_comm_alloc(size_t data_bytes, void** base_out, void* h_handle64);
int drv_comm_open(const void* h_handle64, void** base_out);
int drv_comm_close(void* base);
int drv_comm_free(void* base);
int drv_peer_allreduce(void* const* h_bases, int rank, int world, int64_t n_floats, void* stream);
/* Binds the calling thread's drv_elbo_step calls to a peer-mapped gradient buffer (same arguments as
 * drv_peer_allreduce) for DRV_FLAG_FUSED_ALLREDUCE; h_bases == NULL unbinds. */
int drv_comm_bind(void* const* h_bases, int rank, int world, int64_t n_floats);
/* 1 when drv_elbo_step runs this shape as a padded (aligned) problem - see k_pad.cu -, else 0 */
int drv_shape_padded(const drv_dims* dims);

/* Section timing with CUDA events recorded on the step's stream (bench.py's live roofline
 * measurement).  drv_profile_enable(1) arms it (up to 64 steps are kept); drv_profile_read
 * synchronises on the recorded events and returns the mean milliseconds per section. */
#define DRV_N_SECTIONS 7
#define DRV_SEC_ENC_FWD 0        /* encoder forward + sampling/KL */
#define DRV_SEC_DEC_TRUNK_FWD 1  /* decoder blocks before the last Linear */
#define DRV_SEC_DEC_NB_FWD 2     /* last Linear + gene softmax + NB likelihood (forward) */
#define DRV_SEC_DOSE_RESPONSE 3  /* dose-response head forward + backward */
#define DRV_SEC_DEC_NB_BWD 4     /* fused backward of the last Linear (dlogits, dW, db, dAct) */
#define DRV_SEC_DEC_TRUNK_BWD 5  /* decoder trunk backward + sampling backward */
#define DRV_SEC_ENC_BWD 6        /* encoder backward */
int drv_profile_enable(int on);
int drv_profile_read(float* h_ms_per_section, int* h_n_steps);
/* Event pairs around single kernels of the fused decoder path (mean milliseconds per launch):
 * 0 = phase 1 (logsumexp), 1 = phase 2 (NB likelihood [+ NB part of dlogits]), 2 = phase 3 (softmax
 * coupling; hidden width <= 256: the fused coupling + weight-gradient kernel k_dec3w), 3 = dAct GEMM,
 * 4 = dW GEMM (with k_dec3w: the log_r half only).  0 when the kernel did not run. */
#define DRV_N_PROFILED_KERNELS 5
int drv_profile_read_kernels(float* h_ms_per_kernel);

/* Debug: %globaltimer stamps of k_dec3w's first CTA (and entry / exit of three CTAs), written while the environment
 * variable DRV_W3_TRACE is set; 24 x 10 values, layout in tools/w3_trace.py.  Returns the number of values copied. */
int drv_debug_w3_trace(unsigned long long* h_stamps, int n);

/* Number of kernels the last drv_* call on this thread enqueued (for bench.py's gpu_launches). */
int drv_last_launch_count(void);

/* "drvish_b200 <version> sm_100a" */
const char* drv_version(void);

#ifdef __cplusplus
}
#endif
#endif
