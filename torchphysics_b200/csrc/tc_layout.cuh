// Packed parameter image and tile constants shared by the tensor-core forward and reverse kernels.
#pragma once
#include <stddef.h>

namespace tpx {
namespace tc {

constexpr int HP = 64;             // padded hidden width
constexpr int TILE = 32;           // points per tile (x 4 stream quadrants = 128 UMMA rows)

// ------------------------------------------------------------------------------------------
// packed parameter image of the tensor-core path (floats)
//   [W0: d_in x 64][bias: L x 64][WL: d_out x 64][bL: 8]  pad to 256 floats
//   per hidden layer l = 1..L-1, 4 blocks of 4096 floats in swz128(row, k, 64) order:
//     W_hi, W_lo   rows = output neuron j, k = input neuron i     (forward  B operand)
//     WT_hi, WT_lo rows = input neuron i,  k = output neuron j    (reverse  B operand)
// ------------------------------------------------------------------------------------------
__host__ __device__ inline size_t TcLayout_small(int L, int d_in, int d_out) {
  size_t n = (size_t)d_in * HP + (size_t)L * HP + (size_t)d_out * HP + 8;
  return (n + 255) / 256 * 256;
}
struct TcLayout {
  int L, d_in, d_out;
  __host__ __device__ size_t w0() const { return 0; }
  __host__ __device__ size_t bias(int l) const { return (size_t)d_in * HP + (size_t)l * HP; }
  __host__ __device__ size_t wl() const { return (size_t)d_in * HP + (size_t)L * HP; }
  __host__ __device__ size_t bl() const { return wl() + (size_t)d_out * HP; }
  __host__ __device__ size_t small() const { return TcLayout_small(L, d_in, d_out); }
  __host__ __device__ size_t mat(int l, int part) const { return small() + ((size_t)(l - 1) * 4 + part) * 4096; }
  __host__ __device__ size_t total() const { return small() + (size_t)(L - 1) * 4 * 4096; }
};


}  // namespace tc
}  // namespace tpx
