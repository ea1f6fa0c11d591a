// tcgen05 / TMA (sm_100a tensor-core) path: fused decoder-last-layer GEMM + gene-softmax +
// NB likelihood forward and backward.  See tc_decoder.cu.
#pragma once
#include "common.cuh"

namespace drv {

struct TcDecoderArgs {
  const float* U;     // [B][h] input of the last decoder block (output of the previous Linear)
  BnParams bn;        // statistics / affine of the last decoder BatchNorm
  const float* W;     // [2G][h]
  const float* bias;  // [2G]
  const float* x;     // [B][G]
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
};

size_t tc_workspace_bytes(const drv_dims& d);
bool tc_decoder_supported(const drv_dims& d, const drv_hparams& hp);
void tc_decoder_forward(const TcDecoderArgs& a, cudaStream_t st);   // lse, ll, asum, loss
void tc_decoder_backward(const TcDecoderArgs& a, cudaStream_t st);  // dW, db, dY

}  // namespace drv

namespace drv {
namespace tc {
int gemm_tf32(const float* A, int lda, bool a_mn, const float* Bm, int ldb, bool b_mn, float* C, int ldc, int M, int N,
              int K, int max_ctas, bool accumulate, cudaStream_t st);
int gemm_tf32_multi(int n_pairs, const float* const* As, const float* const* Bs, int lda, bool a_mn, int ldb, bool b_mn,
                    float* C, int ldc, int M, int N, int K, int max_ctas, bool accumulate, const float* bias,
                    cudaStream_t st);
}

// Encoder first layer (BatchNorm(G) -> ReLU -> Linear(G, h), reference modules.py:19-21,56) on the
// tensor cores.  See tc_encoder.cu.
struct TcEncoderArgs {
  const float* x;     // [B][G] counts
  BnParams bn;        // statistics / affine of the input BatchNorm (over sqrt(x))
  const float* W;     // [h][G]
  const float* bias;  // [h]
  int B, G, h;
  float* H;           // [B][h] output of the Linear
  void* scratch;
  const float* dH;    // [B][h] backward input
  float *dW, *db;     // [h][G], [h]
  double* gacc;       // [2][G]: sum_b dY, sum_b dY*xhat  (-> d beta, d gamma)
};
size_t tc_encoder_workspace_bytes(const drv_dims& d);
bool tc_encoder_supported(const drv_dims& d, const drv_hparams& hp);
void tc_encoder_forward(const TcEncoderArgs& a, cudaStream_t st);
void tc_encoder_backward(const TcEncoderArgs& a, cudaStream_t st);
}  // namespace drv
