#include "tc.cuh"
namespace drv {
size_t tc_workspace_bytes(const drv_dims&) { return 0; }
bool tc_decoder_supported(const drv_dims&, const drv_hparams&) { return false; }
void tc_decoder_forward(const TcDecoderArgs&, cudaStream_t) {}
void tc_decoder_backward(const TcDecoderArgs&, cudaStream_t) {}
}  // namespace drv
