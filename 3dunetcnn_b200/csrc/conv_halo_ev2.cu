// Lean epilogue variant 2 of the halo-resident convolution kernel (see conv_halo_kernel.cuh): mode 1 (GroupNorm/activation backward + statistics), single-pass bf16.
#include "conv_halo_kernel.cuh"

namespace b200 {

int launch_halo_ev2(int KC, int BN, int TD, int kws, const ConvMaps& maps, const ConvArgs& a, const HaloArgs& h, int grid, cudaStream_t st) {
  return launch_halo_table<2>(KC, BN, TD, kws, maps, a, h, grid, st);
}

}  // namespace b200
