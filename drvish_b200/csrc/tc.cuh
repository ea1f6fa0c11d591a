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
              int K, int splits, bool accumulate_atomic, cudaStream_t st);
}
}  // namespace drv
