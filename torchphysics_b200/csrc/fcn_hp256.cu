// Instantiates the FP32 jet kernels for padded hidden width 256.
#include "fcn_kernels.cuh"
namespace tpx {
TPX_INSTANTIATE_HP(256)
}
