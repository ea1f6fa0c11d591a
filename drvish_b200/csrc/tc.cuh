// tcgen05 / TMA (sm_100a tensor-core) path: fused decoder-last-layer GEMM + gene-softmax +
// NB likelihood forward and backward.  See tc_decoder.cu.
#pragma once
#include "common.cuh"

namespace drv {

struct TcDecoderArgs {
  const float* U;     // [B][h] input of the last decoder block (output of the previous Linear)
  BnParams bn;        // statistics / affine of the last decoder BatchNorm
  const float* W;     // [2G][h]
  const float* W_hi;  // [2G][h] tf32-rounded copy (split once per step for all layers)
  const float* bias;  // [2G]
  const float* x;     // [B][G] counts scored by the likelihood
  const float* rshift;  // [B] or NULL: shift of log_r for the NB count parameter (MCVNBVAE)
  const float* lib;   // [B]
  int B, G, h;
  float eps, c;
  float *lse, *ll, *asum;  // [B] each
  drv_step_out* out;
  void* scratch;
  float* dlogits;   // [B][2G] scratch for d loss / d logits
  bool bwd;
  float *dW, *db;     // [2G][h], [2G]
  float* dY;          // [B][h]: d loss / d (BN output), already gated by the ReLU mask
  SideStream* side;   // optional: weight-gradient GEMM runs there
  const ColStatsOut* next_stats;  // optional: BatchNorm statistics of V for the next block
  bool dv_prepared;   // backward: the tf32-rounded dV and the bias gradient were already produced (bn_bwd_apply_round)
  bool w_prepared;    // tc_decoder_prepare already left the fp16 copy of W in the scratch
};
// where tc_block_backward expects the tf32-rounded dV of a block ([B][dout])
float* tc_block_rounded_dv(void* scratch, int B, int din, int dout);

size_t tc_workspace_bytes(const drv_dims& d);
// Start-of-step work that only depends on the parameters (fp16 copy of W; needs W, G, h, scratch): may run on
// another stream ahead of tc_decoder_forward.  Returns whether anything was prepared.
bool tc_decoder_prepare(const TcDecoderArgs& a, cudaStream_t st);
bool tc_decoder_supported(const drv_dims& d, const drv_hparams& hp);
void tc_decoder_forward(const TcDecoderArgs& a, cudaStream_t st);   // lse, ll, asum, loss
void tc_decoder_backward(const TcDecoderArgs& a, cudaStream_t st);  // dW, db, dY

}  // namespace drv

namespace drv {
namespace tc {
int gemm_tf32(const float* A, int lda, bool a_mn, const float* Bm, int ldb, bool b_mn, float* C, int ldc, int M, int N,
              int K, int max_ctas, bool accumulate, cudaStream_t st);
// fp16 operands (kind::f16; same mantissa as tf32, half the bytes): C (+)= alpha A B^T
// alpha_dev: optional device float multiplied into alpha (dynamic operand scaling)
int gemm_f16(const void* A, int lda, bool a_mn, const void* Bm, int ldb, bool b_mn, float* C, int ldc, int M, int N, int K,
             int max_ctas, bool accumulate, float alpha, cudaStream_t st, const float* alpha_dev = nullptr);
// C = bias + sum of the k-part slabs in index order; with `stats` also the column sums / sums of squares
// of C and (last CTA) mean / rstd / running buffers of the BatchNorm that consumes C.  Returns true
// when the statistics were produced.
bool reduce_parts(const float* parts, int n_parts, int M, int N, const float* bias, float* C, int ldc, cudaStream_t st,
                  const ColStatsOut* stats = nullptr);
// Encoder-input Linear forward with the sqrt / BatchNorm / ReLU / hi-lo split prologue fused into the
// GEMM (tc_encfwd.cu).  Writes V = relu(bn(sqrt(x))) W^T + b and the tf32 'hi' part of the activations.
// h16: fp16 operands - w_hi / w_lo are the fp16 hi / scaled-lo parts from encfwd_split_weight_h16 and a_hi
// receives the fp16 hi part of the activations.
int encfwd_fused(const float* x, int B, int G, BnParams bn, const float* w_hi, const float* w_lo, int h,
                 const float* bias, float* V, float* a_hi, float* partials, int max_parts, cudaStream_t st,
                 const ColStatsOut* stats = nullptr, bool h16 = false, bool sqrt_in = true);  // returns 1 when `stats` were produced
bool encfwd_h16_supported(int G, int h);
void encfwd_split_weight_h16(const float* W, int h, int G, void* w_hi_h, void* w_lo_h, cudaStream_t st);
// d gamma / d beta sums of the encoder's input BatchNorm from dH and W without materialising dA (tc_bn0grad.cu)
int bn0grad_fused(const float* dH_r, const float* W_hi, const float* x, int B, int G, int h, BnParams bn, double* gacc,
                  cudaStream_t st);
int gemm_tf32_multi(int n_pairs, const float* const* As, const float* const* Bs, int lda, bool a_mn, int ldb, bool b_mn,
                    float* C, int ldc, int M, int N, int K, int max_ctas, bool accumulate, const float* bias,
                    cudaStream_t st, int max_chain_kb = 0, float* partials = nullptr, int max_parts = 0,
                    const ColStatsOut* stats = nullptr);  // returns 1 when `stats` were produced
}

// One BatchNorm -> ReLU -> Linear block (reference modules.py:19-21) on the tensor cores: 3xTF32
// forward, single-pass tf32 backward.  Used for the encoder's first layer (sqrt input, B x G) and
// for the h x h trunk layers.  See tc_block.cu.
struct TcBlockArgs {
  const float* U;     // [B][din] block input (x for the first encoder block)
  bool sqrt_in;       // apply sqrt before the BatchNorm (encoder input, modules.py:56)
  bool no_fused_prologue;  // debug: materialise A_hi / A_lo instead of the fused-prologue kernel
  BnParams bn;        // statistics / affine of the block's BatchNorm
  const float* W;     // [dout][din]
  const float *W_hi, *W_lo;  // tf32 hi / lo parts of W (split once per step for all layers)
  const float* bias;  // [dout]
  int B, din, dout;
  float* V;           // [B][dout] output of the Linear
  void* scratch;
  const float* dV;    // [B][dout] backward input
  float *dW, *db;     // [dout][din], [dout]
  float* dY;          // [B][din]: d loss / d (BN output), ReLU-gated; NULL -> reduce to gacc instead
  double* gacc;       // [2][din]: sum_b dY, sum_b dY*xhat  (-> d beta, d gamma), used when dY == NULL
  SideStream* side;   // optional: weight-gradient GEMM runs there
  const ColStatsOut* next_stats;  // optional: BatchNorm statistics of V for the next block
  bool dv_prepared;   // backward: the tf32-rounded dV and the bias gradient were already produced (bn_bwd_apply_round)
  bool w_prepared;    // tc_block_prepare already left the fp16 hi / lo parts of W in the scratch
  int part;           // backward: 0 = everything, 1 = weight gradient only, 2 = everything but the weight gradient (after part 1)
};
// where tc_block_backward expects the tf32-rounded dV of a block ([B][dout])
float* tc_block_rounded_dv(void* scratch, int B, int din, int dout);
// Start-of-step work of the encoder input block that only depends on the parameters (fp16 hi / lo split of W;
// needs W, B, din, dout, sqrt_in, no_fused_prologue, scratch).  Returns whether anything was prepared.
bool tc_block_prepare(const TcBlockArgs& a, cudaStream_t st);
size_t tc_block_workspace_bytes(int B, int din, int dout);
bool tc_block_forward_supported(int B, int din, int dout, const drv_hparams& hp);
bool tc_block_backward_supported(int B, int din, int dout, const drv_hparams& hp);
bool tc_block_forward(const TcBlockArgs& a, cudaStream_t st);  // true: next_stats were produced
void tc_block_backward(const TcBlockArgs& a, cudaStream_t st);
// hi = rna_tf32(src), lo = rna_tf32(src - hi) over the whole flat parameter buffer
void split_params_tf32(const float* src, size_t n, float* hi, float* lo, cudaStream_t st);
}  // namespace drv
