// Lean epilogue variant 1 of the halo-resident convolution kernel (see conv_halo_kernel.cuh): mode 0 (+residual, statistics), single-pass bf16, no dropout scale / bias / zero boundary.
#include "conv_halo_kernel.cuh"

namespace b200 {

int launch_halo_ev1(int KC, int BN, int TD, int kws, const ConvMaps& maps, const ConvArgs& a, const HaloArgs& h, int grid, cudaStream_t st) {
  return launch_halo_table<1>(KC, BN, TD, kws, maps, a, h, grid, st);
}

}  // namespace b200
