// Host-side interface of the tensor-core (tcgen05) jet kernels, used by api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/tpx.h"
#include "fcn_common.cuh"

namespace tpx {
bool tc_net_supported(const tpx_fcn_desc* net);                 // network shape has a tensor-core kernel
bool tc_net_fits(const tpx_fcn_desc* net);                      // ... its shape part (no debug switch)
bool tcw_takes_narrow(const tpx_fcn_desc* net);                 // narrow but too deep for tc: one padded 128-neuron block on tcw
bool tc_supported(const tpx_fcn_desc* net, int nd, int lap);    // ... and so has this jet
size_t tc_packed_floats(const tpx_fcn_desc* net);
int tc_pack(const tpx_fcn_desc* net, const void* const* params_host, float* img, cudaStream_t stream);
int tc_forward(const tpx_fcn_desc* net, const JetArgs& jet, int lap, const float* img, const float* x, int64_t n,
               float* u, float* du, float* lapo, float* saved, int sm_count, cudaStream_t stream);
bool tc_bwd_supported(const tpx_fcn_desc* net, int nd, int lap);
size_t tc_bwd_workspace_bytes(const tpx_fcn_desc* net, int sm_count);
int tc_backward(const tpx_fcn_desc* net, const Layout& lay32, const JetArgs& jet, int lap, const float* img, const float* x,
                int64_t n, const float* gu, const float* gdu, const float* glap, double* gacc, float* ckpt, size_t ckpt_bytes,
                int saved, int sm_count, cudaStream_t stream);
size_t tc_saved_bytes(const tpx_fcn_desc* net, int64_t n);   // activation-jet checkpoints of n points
// wide path (hidden width 65..128, tcw_kernels.cu): layer-major launches through a saved-activation buffer
bool tcw_supported(const tpx_fcn_desc* net, int nd, int lap);
size_t tcw_saved_bytes(const tpx_fcn_desc* net, int nd, int lap, int64_t n);   // checkpoints + reverse scratch
int tcw_forward(const Layout& lay, const JetArgs& jet, int lap, const float* packed, const float* x, int64_t n,
                float* u, float* du, float* lapo, float* saved, int sm_count, cudaStream_t stream);
int tcw_backward(const Layout& lay, const JetArgs& jet, int lap, const float* packed, const float* x, int64_t n,
                 const float* gu, const float* gdu, const float* glap, double* gacc, float* saved, int sm_count,
                 cudaStream_t stream);
// the same for the sin activation (tcw_sin.cu)
int tcws_forward(const Layout& lay, const JetArgs& jet, int lap, const float* packed, const float* x, int64_t n,
                 float* u, float* du, float* lapo, float* saved, int sm_count, cudaStream_t stream);
int tcws_backward(const Layout& lay, const JetArgs& jet, int lap, const float* packed, const float* x, int64_t n,
                  const float* gu, const float* gdu, const float* glap, double* gacc, float* saved, int sm_count,
                  cudaStream_t stream);
// points-as-lanes fp16-split path (hidden width <= 64; tcp_fwd.cu, tcp_bwd.cu): supersedes tc_* where it applies
bool tcp_net_supported(const tpx_fcn_desc* net);
bool tcp_supported(const tpx_fcn_desc* net, int nd, int lap);
size_t tcp_packed_floats(const tpx_fcn_desc* net);
int tcp_pack(const tpx_fcn_desc* net, const void* const* params_host, float* img, cudaStream_t stream);
size_t tcp_ckpt_floats_per_tile(int L, int S);
// reverse kernel of that path (tce_bwd.cu: rows = (stream, point), split-fp16 operand images)
bool tce_bwd_supported(const tpx_fcn_desc* net, int nd, int lap);
size_t tce_saved_bytes(const tpx_fcn_desc* net, int nd, int lap, int64_t n);   // checkpoints + tile statistics + dW partial sums
int64_t tce_tiles_in(const tpx_fcn_desc* net, int nd, int lap, size_t bytes);
void tce_saved_split(const tpx_fcn_desc* net, int nd, int lap, int64_t ntiles, float* base, float** ckpt, float** stats, float** dwpart);
int tcp_forward(const tpx_fcn_desc* net, const JetArgs& jet, int lap, const float* img, const float* x, int64_t n,
                float* u, float* du, float* lapo, float* ckpt, float* stats, int sm_count, cudaStream_t stream);
int tce_backward(const tpx_fcn_desc* net, const Layout& lay32, const JetArgs& jet, int lap, const float* img, const float* x,
                 int64_t n, const float* gu, const float* gdu, const float* glap, double* gacc, float* buf, size_t buf_bytes,
                 int saved, int sm_count, cudaStream_t stream);
extern long long* g_tc_debug;   // development: device buffer of >= 512 int64 timestamps, or NULL
}  // namespace tpx
